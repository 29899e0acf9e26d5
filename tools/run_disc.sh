#!/bin/bash
# GPU check of the Oobleck discriminator: its parity tests (all of them, no -x) and the stand-alone bench leg
# (KVAE_DISC_GENERIC=1 python bench.py --workload discriminator: the same on the generic layer kernels)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_discriminator.py -q -m gpu 2>&1 | tail -120 > gpurun_out/r02_disc_gputest.log
timeout 600 python bench.py --workload discriminator --steps 12 > gpurun_out/r02_bench_discriminator.json 2> gpurun_out/r02_bench_discriminator.err
tail -5 gpurun_out/r02_disc_gputest.log
cat gpurun_out/r02_bench_discriminator.json
tail -5 gpurun_out/r02_bench_discriminator.err
