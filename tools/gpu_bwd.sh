#!/bin/bash
# Backward-pass GPU pass: parity tests, training-step timing, ncu launch list of one training step.
TAG=${1:-bw}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -30 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python tools/train_step_time.py > gpurun_out/${TAG}_train_time.txt 2>&1; cat gpurun_out/${TAG}_train_time.txt
timeout 300 python tools/train_step_time.py 10 4 5 4 >> gpurun_out/${TAG}_train_time.txt 2>&1; tail -4 gpurun_out/${TAG}_train_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_train_launches.csv \
    python tools/train_step_time.py 2 32 5 4 > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
