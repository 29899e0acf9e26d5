#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/tma_copy.log
: > $LOG
for cfg in "32 16 2 16" "32 16 4 16" "32 16 2 32" "32 32 2 16" "32 64 2 16" "64 16 2 16" "64 16 4 16" "64 32 2 16" "128 16 2 16" "128 32 2 16" "128 64 2 8" "64 16 2 8" "32 16 2 8" "128 16 4 16"; do
  timeout 60 ./build/tma_copy_probe $cfg >> $LOG 2>&1 || echo "exit $? ($cfg)" >> $LOG
done
cat $LOG
