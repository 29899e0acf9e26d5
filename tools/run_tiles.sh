#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/tiles.log
: > $LOG
for w in 2 0 1 2; do
  AB_STEPS=1 KVAE_WIDE_TILES=$w timeout 300 python tools/ab_decode.py "wide_tiles=$w" >> $LOG 2>&1
done
grep -E "^AB|Traceback|rror" $LOG | cut -c1-200
python - <<'PY'
import re
runs=[];cur=None
for l in open('gpurun_out/tiles.log'):
    if l.startswith('AB '): cur=[l.split(':')[0]]; runs.append(cur)
    m=re.match(r'\s+dec step\s+(\d+)\s+([\d.]+) ms',l)
    if m and cur is not None: cur.append(float(m.group(2)))
print('step  '+'  '.join(r[0][3:] for r in runs))
for i in range(0,23):
    print(f'{i:4d}  '+'  '.join(f'{r[1+i]:13.3f}' for r in runs if len(r)>1+i))
PY
