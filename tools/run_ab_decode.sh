#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/ab_decode.log
: > $LOG
for pf in 0 1 3; do
  echo "--- probe epi 2 pf $pf" >> $LOG
  KVAE_RU_PF=$pf KVAE_RU_EPI=2 timeout 120 ./build/umma_probe ru 1 4 442368 1 >> $LOG 2>&1
done
for spec in "1 2 8192" "9 3 5000"; do
  KVAE_RU_GRID=7 KVAE_RU_EPI=2 timeout 100 ./build/umma_probe ru $spec >> $LOG 2>&1
done
KVAE_SPLIT_PRODUCER=0 KVAE_RU_EPI=1 timeout 300 python tools/ab_decode.py "one producer,epi1" >> $LOG 2>&1
KVAE_SPLIT_PRODUCER=0 KVAE_RU_EPI=2 timeout 300 python tools/ab_decode.py "one producer,epi2(split A)" >> $LOG 2>&1
AB_STEPS=1 KVAE_RU_EPI=2 timeout 300 python tools/ab_decode.py "split producers,epi2" >> $LOG 2>&1
KVAE_RU_PF=1 KVAE_RU_EPI=2 timeout 300 python tools/ab_decode.py "split producers,epi2,pf1" >> $LOG 2>&1
grep -E "^AB|PERF|RESULT|--- probe|rror|Traceback|   dec step" $LOG | cut -c1-220
