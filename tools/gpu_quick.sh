#!/bin/bash
# Quick GPU pass: parity tests, shape sweep, short bench.
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -25 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python tools/sweep_rows.py > gpurun_out/${TAG}_sweep.log 2>&1; tail -14 gpurun_out/${TAG}_sweep.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
