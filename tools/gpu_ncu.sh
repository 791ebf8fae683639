#!/bin/bash
# one full ncu capture (with source counters) of the sequence kernel at the bench workload
TAG=${1:-ncu}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sqair_sequence -s 4 -c 1 -o gpurun_out/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/${TAG}_prof.ncu-rep
