#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_training.py tests/test_gpu_ddp.py -x -q > gpurun_out/train_tests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/train_tests.log
for rep in 1 2; do
  KVAE_LOAD_PARAMS_PER_LAYER=1 python bench.py --workload train --no-cpu-baseline --steps 20 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('per-layer ', j['ms_per_step'], j.get('gpu_launches'))"
  python bench.py --workload train --no-cpu-baseline --steps 20 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('batched   ', j['ms_per_step'], j.get('gpu_launches'))"
done
