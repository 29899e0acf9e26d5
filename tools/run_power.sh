#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/power.log
: > $LOG
( nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 200 > gpurun_out/power_smi.log ) &
SMI=$!
sleep 1
echo "--- new, 60 steps" >> $LOG
python bench.py --steps 60 --no-cpu-baseline >> $LOG 2>/dev/null
sleep 2
echo "--- old, 60 steps" >> $LOG
KVAE_LIB=$PWD/build/libkvae_nu.so KVAE_RU_EPI=0 KVAE_SPLIT_PRODUCER=0 python bench.py --steps 60 --no-cpu-baseline >> $LOG 2>/dev/null
kill $SMI
python - <<'PY'
import json
for l in open('gpurun_out/power.log'):
    if l.startswith('---'): print(l.strip()); continue
    j=json.loads(l); d=j['roofline_dominant_kernel']
    print(f"  step {j['ms_per_step']:.2f} ms value {j['value']:.0f} decode {j['decode_only']['ms_per_step']:.2f} ms RU {d['kernel_ms_per_step']:.2f} clocks {j['clocks']}")
PY
sort gpurun_out/power_smi.log | uniq -c | sort -rn | head -25
