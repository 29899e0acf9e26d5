#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/ab_bench.log
: > $LOG
for rep in 1 2; do
  echo "--- old (non-uniform lib, KVAE_RU_EPI=0, one producer)" >> $LOG
  KVAE_LIB=$PWD/build/libkvae_nu.so KVAE_RU_EPI=0 KVAE_SPLIT_PRODUCER=0 python bench.py --no-cpu-baseline >> $LOG 2>/dev/null
  echo "--- new" >> $LOG
  python bench.py --no-cpu-baseline >> $LOG 2>/dev/null
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_bench.log'):
    if l.startswith('---'): print(l.strip()); continue
    j=json.loads(l)
    d=j['roofline_dominant_kernel']
    print(f"  step {j['ms_per_step']:.2f} ms value {j['value']:.0f} e2e {j['e2e']['value']:.0f} decode {j['decode_only']['ms_per_step']:.2f} ms {j['decode_only']['tflops']:.0f} TF  RU {d['kernel_ms_per_step']:.2f} ms  clocks {j['clocks']['sm_mhz']}")
PY
