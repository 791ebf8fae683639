#!/bin/bash
# Round-2 evidence pass on one GPU: full GPU suite, bench (both arms), launch lists, ncu captures, sanitizer.
TAG=${1:-r02z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/train_step_time.py > gpurun_out/${TAG}_train_time.txt 2>&1; head -8 gpurun_out/${TAG}_train_time.txt
# launch list of the inference bench (as in round 1) and of one full-size training step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4500 -c 3200 --csv --log-file gpurun_out/${TAG}_train_launches.csv \
    python tools/train_step_time.py > gpurun_out/${TAG}_ncu_train_launch.log 2>&1; echo "ncu train launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sqair_sequence -s 4 -c 1 -o gpurun_out/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgrad_kernel -s 300 -c 3 -o gpurun_out/${TAG}_dgrad_prof \
    python tools/train_step_time.py 2 32 5 4 > gpurun_out/${TAG}_ncu_dgrad.log 2>&1; echo "ncu dgrad rc=$?"
ls -la gpurun_out | tail -15
