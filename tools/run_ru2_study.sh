#!/bin/bash
# conv_ru2_kernel study: L2 prefetch A/B, ablations, one ncu --set full capture.
mkdir -p gpurun_out
LOG=gpurun_out/ru2_study.log
: > $LOG
run() { echo "--- $*" >> $LOG; env "$@" KVAE_RU_EPI=2 timeout 120 ./build/umma_probe ru 1 4 442368 1 >> $LOG 2>&1; echo "exit $?" >> $LOG; }
for rep in 1 2; do
  run KVAE_RU_PF=0
  run KVAE_RU_PF=1
  run KVAE_RU_PF=2
  run KVAE_RU_PF=3
done
for dbg in 1 2 3 32 64 96 99; do run KVAE_RU_PF=0 KVAE_RU_DBG=$dbg; done
run KVAE_RU_PF=3 KVAE_RU_DBG=32
run KVAE_RU_PF=3 KVAE_RU_DBG=2
KVAE_RU_EPI=2 timeout 280 ncu --set full --import-source on --clock-control none -k regex:conv_ru2 -s 3 -c 1 -f -o gpurun_out/prof_ru2 ./build/umma_probe ru 1 4 442368 1 > gpurun_out/ncu_ru2.log 2>&1
echo "ncu exit $?" >> $LOG
grep -E "PERF|exit [1-9]|^--- |failed|timeout|rror" $LOG | cut -c1-170
