#!/bin/bash
# Training-surface GPU pass: new tests first, then whole suite, bench line (both arms).
TAG=${1:-tr}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/${TAG}_pytest_train.log 2>&1; echo "train tests rc=$?"
tail -30 gpurun_out/${TAG}_pytest_train.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
