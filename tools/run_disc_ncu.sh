#!/bin/bash
# launch list (ncu, gpu__time_duration only) of the discriminator bench leg, our module only, one warm-up + one timed step
mkdir -p gpurun_out
KVAE_DISC_NO_REF=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_disc_launches_ncu.csv \
  python bench.py --workload discriminator --steps 1 > gpurun_out/ncu_disc.log 2>&1
tail -2 gpurun_out/ncu_disc.log | cut -c1-300
