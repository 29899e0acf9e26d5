#!/bin/bash
# A/B of the fused ResidualUnit kernels (KVAE_RU_EPI=0 first generation, 1 fragment-mapped epilogue, 2 conv_ru2_kernel):
# correctness against the CPU loops of umma_probe (incl. many tiles per CTA via KVAE_RU_GRID), then per-launch time
# at the bench's row counts.
mkdir -p gpurun_out
LOG=gpurun_out/ru_ab.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
for epi in ${EPIS:-2}; do
  for spec in "1 2 1500" "3 1 2048" "9 3 777"; do
    KVAE_RU_EPI=$epi timeout 100 ./build/umma_probe ru $spec >> $LOG 2>&1
    echo "exit $? (epi $epi ru $spec)" >> $LOG
  done
  for spec in "1 2 8192" "9 3 5000"; do
    KVAE_RU_GRID=7 KVAE_RU_EPI=$epi timeout 100 ./build/umma_probe ru $spec >> $LOG 2>&1
    echo "exit $? (epi $epi grid 7 ru $spec)" >> $LOG
  done
done
if grep -q "FAIL\|exit [1-9]" $LOG; then grep -E "RESULT|exit [1-9]|first bad|timeout|rror" $LOG | cut -c1-170; exit 1; fi
for cfg in ${CFGS:-"0 3 6" "1 3 6" "2 3 6" "2 2 5" "2 4 8" "2 5 9" "2 1 3" "2 0 1"}; do
  set -- $cfg
  for spec in "1 4 442368" "9 4 442368"; do
    echo "--- epi $1 k0 $2 k1 $3" >> $LOG
    KVAE_RU_EPI=$1 KVAE_RU_K0=$2 KVAE_RU_K1=$3 timeout 120 ./build/umma_probe ru $spec 1 >> $LOG 2>&1
    echo "exit $? (epi $1 perf $spec)" >> $LOG
  done
done
grep -E "RESULT|PERF|exit [1-9]|--- epi|failed|timeout|rror" $LOG | cut -c1-170
