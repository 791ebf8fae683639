#!/bin/bash
# Reverse-program kernel: backward parity tests, then the training-step timing with both backward paths + per-op profile.
TAG=${1:-prog}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q > gpurun_out/${TAG}_pytest_bwd.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_bwd.log
tail -15 gpurun_out/${TAG}_pytest_bwd.log
timeout 300 python tools/train_step_time.py > gpurun_out/${TAG}_train_time.txt 2>&1; cat gpurun_out/${TAG}_train_time.txt
SQAIR_BWD_LAUNCHES=1 timeout 300 python tools/train_step_time.py > gpurun_out/${TAG}_train_time_launches.txt 2>&1; cat gpurun_out/${TAG}_train_time_launches.txt
if [ -f sqair_b200/csrc/exp_progprof.so ]; then
  SQAIR_LIB=$PWD/sqair_b200/csrc/exp_progprof.so SQAIR_PROG_PRINT=1 timeout 300 python tools/train_step_time.py > gpurun_out/${TAG}_prof.txt 2>&1
  L=$(grep -n "reverse program" gpurun_out/${TAG}_prof.txt | tail -2 | head -1 | cut -d: -f1); tail -n +$L gpurun_out/${TAG}_prof.txt | head -40
fi
