#!/bin/bash
# Runs every umma_probe case/variant in its own process (a trap in one cannot poison the rest).
mkdir -p gpurun_out
LOG=gpurun_out/probe.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
read NC NP < <(./build/umma_probe count)
for c in $(seq 0 $((NC-1))); do
  for v in 0 1 2; do
    timeout 60 ./build/umma_probe $c $v >> $LOG 2>&1
    echo "exit $? (case $c variant $v)" >> $LOG
  done
done
for p in $(seq 0 $((NP-1))); do
  timeout 120 ./build/umma_probe perf $p >> $LOG 2>&1
  echo "exit $? (perf $p)" >> $LOG
done
grep -E "RESULT|PERF|exit [1-9]" $LOG | tail -80
