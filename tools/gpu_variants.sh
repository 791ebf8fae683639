#!/bin/bash
# GPU box: time every tools/build_variants.sh variant at one launch shape (with output checksums)
TAG=${1:-var}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_variants.txt
bash tools/run_variants.sh >> gpurun_out/${TAG}_variants.txt 2>&1
cat gpurun_out/${TAG}_variants.txt
