#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
LOG=gpurun_out/frag.log
: > $LOG
run() { AB_STEPS=1 env "$@" timeout 300 python tools/ab_decode.py "$*" >> $LOG 2>&1; }
run KVAE_FRAG_EPI=0
run KVAE_FRAG_EPI=1
run KVAE_FRAG_EPI=0
run KVAE_FRAG_EPI=1
grep -E "^AB|Traceback|rror" $LOG | cut -c1-200
python - <<'PY'
import re
runs=[];cur=None
for l in open('gpurun_out/frag.log'):
    if l.startswith('AB '): cur=[l.split(':')[0]]; runs.append(cur)
    m=re.match(r'\s+dec step\s+(\d+)\s+([\d.]+) ms',l)
    if m and cur is not None: cur.append(float(m.group(2)))
print('step  '+'  '.join(r[0][3:] for r in runs))
for i in list(range(0,24))+[29]:
    print(f'{i:4d}  '+'  '.join(f'{r[1+i]:15.3f}' for r in runs if len(r)>1+i))
PY
for rep in 1 2; do
  for w in 0 1; do
    KVAE_FRAG_EPI=$w python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('frag=$w roundtrip value', round(j['value']), 'e2e', round(j['e2e']['value']), 'ms', round(j['ms_per_step'],2), 'decode', round(j['decode_only']['ms_per_step'],2), j['clocks']['sm_mhz'])"
  done
done
