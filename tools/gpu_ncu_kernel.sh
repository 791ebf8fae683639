#!/bin/bash
# one ncu --set full capture of a few launches of one kernel inside a short training step: gpu_ncu_kernel.sh tag regex skip count
TAG=${1:-k}; PAT=${2:-dgrad_kernel}; SKIP=${3:-100}; CNT=${4:-3}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$PAT -s $SKIP -c $CNT -o gpurun_out/${TAG}_prof \
    python tools/train_step_time.py 2 32 5 4 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/${TAG}_prof.ncu-rep
