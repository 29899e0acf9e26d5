#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/ru_sb.log
: > $LOG
for cfg in "0 2" "0 3" "0 4" "131 2" "131 4" "131 7"; do
  set -- $cfg
  echo "--- dbg $1 SB $2" >> $LOG
  KVAE_RU_DBG=$1 KVAE_RU_SB=$2 KVAE_RU_EPI=2 timeout 120 ./build/umma_probe ru 1 4 442368 1 >> $LOG 2>&1
  echo "exit $?" >> $LOG
done
grep -E "PERF|RU d=|exit [1-9]|--- dbg|failed|timeout|rror" $LOG | cut -c1-170
