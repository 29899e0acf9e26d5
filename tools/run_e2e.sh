#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -k "host_pipeline" > gpurun_out/pytest_hp.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_hp.log
for rep in 1 2; do
python bench.py --no-cpu-baseline 2>gpurun_out/bench_e2e.err | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value', round(j['value']), 'e2e', round(j['e2e']['value']), 'ms', j['ms_per_step'])"
done
tail -3 gpurun_out/bench_e2e.err
