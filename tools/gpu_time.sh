#!/bin/bash
TAG=${1:-tm}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -3
timeout 300 python tools/train_step_time.py > gpurun_out/${TAG}_train_time.txt 2>&1; cat gpurun_out/${TAG}_train_time.txt
timeout 300 python tools/train_step_time.py 10 4 5 4 >> gpurun_out/${TAG}_train_time.txt 2>&1; tail -6 gpurun_out/${TAG}_train_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_train_launches.csv \
    python tools/train_step_time.py 2 32 5 4 > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/${TAG}_train_launches.csv 8
