#!/bin/bash
# GPU check of the Oobleck discriminator: its parity tests (all of them, no -x) and the stand-alone bench leg, on the
# kernels written for the nets' conv geometry and (KVAE_DISC_GENERIC=1) on the generic layer kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_disc_gpu.csv
timeout 900 python -m pytest tests/test_gpu_discriminator.py -q -m gpu 2>&1 | tail -120 > gpurun_out/r02_disc_gputest.log
timeout 600 python bench.py --workload discriminator --steps 12 > gpurun_out/r02_bench_discriminator.json 2> gpurun_out/r02_bench_discriminator.err
KVAE_DISC_GENERIC=1 timeout 600 python bench.py --workload discriminator --steps 12 > gpurun_out/r02_bench_discriminator_generic.json 2>> gpurun_out/r02_bench_discriminator.err
tail -5 gpurun_out/r02_disc_gputest.log
cat gpurun_out/r02_bench_discriminator.json gpurun_out/r02_bench_discriminator_generic.json
tail -5 gpurun_out/r02_bench_discriminator.err
