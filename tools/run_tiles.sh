#!/bin/bash
# tuning sweeps of conv_umma2's plan-level knobs on the 8-clip decode / encode (tools/ab_decode.py)
mkdir -p gpurun_out
LOG=gpurun_out/tiles.log
: > $LOG
run() { AB_STEPS=1 env "$@" timeout 300 python tools/ab_decode.py "$*" >> $LOG 2>&1; }
run KVAE_X=0
run KVAE_SA=2
run KVAE_SA=3
run KVAE_SB=3
run KVAE_SB=4
run KVAE_X=0
grep -E "^AB|Traceback|rror" $LOG | cut -c1-200
