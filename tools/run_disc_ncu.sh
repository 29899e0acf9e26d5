#!/bin/bash
# launch list (ncu, gpu__time_duration only) of the discriminator bench leg + the trainer test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -q -m gpu -k discriminator 2>&1 | tail -30 > gpurun_out/r02_disc_trainer_test.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_disc_launches_ncu.csv \
  python bench.py --workload discriminator --steps 4 > gpurun_out/ncu_disc.log 2>&1
tail -5 gpurun_out/r02_disc_trainer_test.log
