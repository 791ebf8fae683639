#!/bin/bash
# full GPU suite + bench line + training-step timing
TAG=${1:-full}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -8 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python tools/train_step_time.py > gpurun_out/${TAG}_train_time.txt 2>&1; cat gpurun_out/${TAG}_train_time.txt
