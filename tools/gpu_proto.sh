#!/bin/bash
# tcgen05 forward-layer prototype against the mma.sync stream of the product kernel (same box, same session)
TAG=${1:-proto}
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/fwd_dense_tc tools/fwd_dense_tc.cu 2>/dev/null && timeout 120 /tmp/fwd_dense_tc 2>&1 | tee gpurun_out/${TAG}_fwd_dense_tc.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_stream tools/mma_stream.cu 2>/dev/null && timeout 120 /tmp/mma_stream 2>&1 | grep "grid 148" | grep "threads 384" | tee gpurun_out/${TAG}_mma_stream.txt
