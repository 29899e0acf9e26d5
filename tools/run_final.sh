#!/bin/bash
# Round-end validation on one B200: GPU tests, smoke, bench line (+ optional ncu passes with NCU=1).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench.log; echo "bench exit $?"
cut -c1-300 gpurun_out/bench_final.json
if [ -n "$NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
ncu --set full --import-source on --clock-control none -k regex:conv_ru2 -s 40 -c 1 -f -o gpurun_out/prof_ru2_bench python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ru2_bench.log 2>&1; echo "ncu full exit $?"
fi
