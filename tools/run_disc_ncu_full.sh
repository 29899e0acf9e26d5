#!/bin/bash
# ncu --set full of the forward conv of the first multi-scale net's layers 2-4 (32->64, 64->128, 128->256 at 8 x 5 s stereo)
mkdir -p gpurun_out
KVAE_DISC_NO_REF=1 timeout 140 ncu --set full --import-source on --clock-control none \
  --kernel-name regex:disc_conv15_fwd_kernel --launch-skip 1 --launch-count 3 -f -o gpurun_out/r02_disc_conv15_fwd_msd0_l2_l4 \
  python bench.py --workload discriminator --steps 1 > gpurun_out/ncu_full_disc.log 2>&1
tail -3 gpurun_out/ncu_full_disc.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
