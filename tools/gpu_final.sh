#!/bin/bash
# Round-2 evidence pass on one GPU: full GPU suite, bench (both arms), launch lists, ncu capture of the reverse-program kernel,
# per-operation profile of the instrumented build.
TAG=${1:-r02b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/train_step_time.py > gpurun_out/${TAG}_train_time.txt 2>&1; head -8 gpurun_out/${TAG}_train_time.txt
SQAIR_BWD_LAUNCHES=1 timeout 300 python tools/train_step_time.py > gpurun_out/${TAG}_train_time_launches.txt 2>&1; head -8 gpurun_out/${TAG}_train_time_launches.txt
timeout 300 python tools/train_step_time.py 10 4 5 4 > gpurun_out/${TAG}_train_time_c3shard.txt 2>&1; head -3 gpurun_out/${TAG}_train_time_c3shard.txt
# launch list of the inference bench (as in round 1) and of full-size training steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 420 --csv --log-file gpurun_out/${TAG}_train_launches.csv \
    python tools/train_step_time.py > gpurun_out/${TAG}_ncu_train_launch.log 2>&1; echo "ncu train launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bwd_program -s 2 -c 1 -o gpurun_out/${TAG}_program_prof \
    python tools/train_step_time.py > gpurun_out/${TAG}_ncu_program.log 2>&1; echo "ncu program rc=$?"
if [ -f sqair_b200/csrc/exp_progprof.so ]; then
  SQAIR_LIB=$PWD/sqair_b200/csrc/exp_progprof.so SQAIR_PROG_PRINT=1 timeout 300 python tools/train_step_time.py > gpurun_out/${TAG}_program_ops_raw.txt 2>&1
  L=$(grep -n "reverse program" gpurun_out/${TAG}_program_ops_raw.txt | tail -2 | head -1 | cut -d: -f1); tail -n +$L gpurun_out/${TAG}_program_ops_raw.txt | head -36 > gpurun_out/${TAG}_program_ops.txt
fi
ls -la gpurun_out | tail -15
