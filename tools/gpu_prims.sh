#!/bin/bash
# Primitive micro-benchmarks + the instrumented (SQAIR_PROFILE) sequence kernel; GPU box only.
TAG=${1:-prims}
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sync_bench tools/sync_bench.cu && /tmp/sync_bench > gpurun_out/${TAG}_sync.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ring_bw tools/ring_bw.cu && /tmp/ring_bw > gpurun_out/${TAG}_ring.txt 2>&1
SQAIR_LIB=$PWD/sqair_b200/csrc/libsqair_b200_prof.so SWEEP=4:3:32:3,3:2:32:3,2:1:32:4 timeout 300 python tools/sweep_rows.py > gpurun_out/${TAG}_phase.txt 2>&1
cat gpurun_out/${TAG}_sync.txt; tail -20 gpurun_out/${TAG}_ring.txt; tail -5 gpurun_out/${TAG}_phase.txt
