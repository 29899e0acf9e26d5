#!/bin/bash
# Ablation of the fused ResidualUnit kernel: which stage bounds a tile?  (KVAE_RU_DBG bits: 1 no skip loads,
# 2 no TMA stores, 4 no SnakeBeta in EPI1, 8 none in EPI2, 16 only 1 of 7 taps multiplied)
mkdir -p gpurun_out
for d in 0 1 2 3 4 8 12 16 28 31; do
  echo -n "dbg=$d  " >> gpurun_out/ru_ablate.log
  KVAE_RU_DBG=$d timeout 60 ./build/umma_probe ru 1 4 442368 1 2>&1 | grep PERF >> gpurun_out/ru_ablate.log
done
cat gpurun_out/ru_ablate.log
