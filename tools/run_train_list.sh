#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches3.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train3.log 2>&1; echo "exit $?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/train_launches3.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
recs=[]
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    try: recs.append((int(r[ii]), r[ki].split('(')[0], float(r[vi].replace(',',''))))
    except: pass
# last step: find last adamw launches -> the step is between the previous adamw group and the last
ad=[i for i,(id_,k,v) in enumerate(recs) if 'adamw' in k]
# steps end with adamw launches (3 per step?) ; take the records after the 4th-from-last group
groups=[]
for i in ad:
    if not groups or i-groups[-1][-1]>5: groups.append([i])
    else: groups[-1].append(i)
start=groups[-2][-1]+1 if len(groups)>=2 else 0
last=recs[start:groups[-1][-1]+1]
agg=collections.defaultdict(lambda:[0,0.0])
for id_,k,v in last: agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
print(f"last step: {len(last)} launches, {tot/1e6:.3f} ms of kernel time")
for k,(c,t) in sorted(agg.items(),key=lambda kv:-kv[1][1])[:16]:
    print(f"{k[:62]:62s} {c:4d} launches {t/1e6:9.3f} ms {100*t/tot:5.1f}%")
PY
