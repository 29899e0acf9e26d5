#!/bin/bash
# Ablation of the fused ResidualUnit kernel through the KVAE_RU_DBG bits (conv_ru.cuh / conv_ru2.cuh): 1 no skip loads,
# 2 no output stores, 16 one of seven taps multiplied, 32 no weight loads, 64 no activation loads, 128 epilogue warps
# run the barrier protocol only.  Results are garbage by design; only the time per launch is read.
mkdir -p gpurun_out
LOG=gpurun_out/ru_ablate3.log
: > $LOG
for dbg in ${DBGS:-0 1 2 3 32 64 96 99 128 131 227 243 16}; do
  echo "--- dbg $dbg" >> $LOG
  KVAE_RU_DBG=$dbg KVAE_RU_EPI=${EPI:-2} timeout 120 ./build/umma_probe ru 1 4 442368 1 >> $LOG 2>&1
  echo "exit $? (dbg $dbg)" >> $LOG
done
grep -E "PERF|exit [1-9]|--- dbg|failed|timeout|rror" $LOG | cut -c1-170
