#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
for rep in 1 2; do
  for w in 2 0; do
    KVAE_WIDE_TILES=$w python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('wide=$w roundtrip value', round(j['value']), 'e2e', round(j['e2e']['value']), 'ms', round(j['ms_per_step'],2), 'decode', round(j['decode_only']['ms_per_step'],2), j['clocks']['sm_mhz'])"
  done
done
for w in 2 0; do
  KVAE_WIDE_TILES=$w python bench.py --workload train --no-cpu-baseline --steps 20 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('wide=$w train ms', round(j['ms_per_step'],2))"
  KVAE_WIDE_TILES=$w python bench.py --workload o12_decode --no-cpu-baseline --steps 5 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('wide=$w o12 value', round(j['value']), 'ms', round(j['ms_per_step'],2))"
  KVAE_WIDE_TILES=$w python bench.py --workload stream --no-cpu-baseline --steps 20 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('wide=$w stream per_chunk_ms', round(j['per_chunk_ms'],4), 'hop', round(j['exact_context_stream']['ms_per_hop'],4))"
done
