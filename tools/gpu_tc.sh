#!/bin/bash
# tcgen05 weight-gradient GEMM: unit parity, throughput, launch timing, then the whole GPU suite and the step timing
TAG=${1:-tc}
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -k "wgrad_op" 2>&1 | tail -3
timeout 120 python tools/wgrad_bench.py 2>&1 | grep wgrad | tee gpurun_out/${TAG}_wgrad_bench.txt
SQAIR_NO_TC=1 timeout 120 python tools/wgrad_bench.py 2>&1 | grep wgrad | sed 's/^/[mma.sync kernel] /' | tee -a gpurun_out/${TAG}_wgrad_bench.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wgrad_tc -c 40 --csv --log-file gpurun_out/${TAG}_wgrad_launches.csv python tools/wgrad_bench.py > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_wgrad_launches.csv 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 3 -c 1 -o gpurun_out/${TAG}_wgrad_prof python tools/wgrad_bench.py > gpurun_out/${TAG}_wgrad_ncu.log 2>&1; echo "ncu full rc=$?"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -4 gpurun_out/${TAG}_pytest_gpu.log
SQAIR_VERBOSE=1 timeout 300 python tools/train_step_time.py 2>&1 | tee gpurun_out/${TAG}_train_time.txt | grep -v "^sqair_backward" | head -8; grep "^sqair_backward" gpurun_out/${TAG}_train_time.txt | tail -1
