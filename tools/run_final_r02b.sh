#!/bin/bash
# last GPU call of the round: the whole GPU suite, the stand-alone discriminator bench, the default bench line
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -80 > gpurun_out/r02_gputest_final.log
tail -3 gpurun_out/r02_gputest_final.log
timeout 120 python bench.py --workload discriminator --steps 12 > gpurun_out/r02_bench_discriminator.json 2> gpurun_out/r02_bench_discriminator.err
cut -c1-600 gpurun_out/r02_bench_discriminator.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_final_tree.json 2> gpurun_out/r02_bench_final_tree.err
cut -c1-400 gpurun_out/r02_bench_final_tree.json
