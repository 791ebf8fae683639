#!/bin/bash
# parity tests + instrumented per-phase cycle counters of the sequence kernel for one launch shape
TAG=${1:-ph}; SHAPE=${2:-5:4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -6 gpurun_out/${TAG}_pytest_gpu.log
SQAIR_LIB=$PWD/sqair_b200/csrc/libsqair_b200_prof.so SWEEP=$SHAPE timeout 300 python tools/sweep_rows.py > gpurun_out/${TAG}_phase.txt 2>&1
grep "profile tid\|trace\] total\|ms/step" gpurun_out/${TAG}_phase.txt | tail -5
grep "trace\] layer" gpurun_out/${TAG}_phase.txt | tail -37
SWEEP=$SHAPE,4:3,6:4 timeout 300 python tools/sweep_rows.py 2>&1 | grep ms/step
