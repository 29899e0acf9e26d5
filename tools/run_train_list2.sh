#!/bin/bash
# ncu launch list of the training step (last of 3 steps), default and KVAE_BWD_DA_F32=1; prints per-kernel totals
mkdir -p gpurun_out
for mode in bf16; do
  if [ $mode = f32 ]; then export KVAE_BWD_DA_F32=1; else unset KVAE_BWD_DA_F32; fi
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches_$mode.csv python tools/prof_train.py 4 3 > gpurun_out/ncu_train_$mode.log 2>&1; echo "exit $?"
  python - $mode <<'PY'
import csv,collections,sys
mode=sys.argv[1]
rows=list(csv.reader(open(f'gpurun_out/train_launches_{mode}.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
recs=[]
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    try: recs.append((int(r[ii]), r[ki].split('(')[0], float(r[vi].replace(',',''))))
    except: pass
ad=[i for i,(id_,k,v) in enumerate(recs) if 'adamw' in k]
groups=[]
for i in ad:
    if not groups or i-groups[-1][-1]>5: groups.append([i])
    else: groups[-1].append(i)
start=groups[-2][-1]+1 if len(groups)>=2 else 0
last=recs[start:groups[-1][-1]+1]
agg=collections.defaultdict(lambda:[0,0.0])
for id_,k,v in last: agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
print(f"[{mode}] last step: {len(last)} launches, {tot/1e6:.3f} ms of kernel time")
for k,(c,t) in sorted(agg.items(),key=lambda kv:-kv[1][1])[:14]:
    print(f"{k[:62]:62s} {c:4d} launches {t/1e6:9.3f} ms {100*t/tot:5.1f}%")
PY
done
