#!/bin/bash
# reverse-program kernel launch shapes (rows per cluster : cluster size)
TAG=${1:-bshape}
mkdir -p gpurun_out
for rc in "0 0" "10 8" "8 4" "6 4" "5 2" "12 8"; do set -- $rc
  echo "SQAIR_BWD_ROWS=$1 SQAIR_BWD_CLUSTER=$2"
  SQAIR_BWD_ROWS=$1 SQAIR_BWD_CLUSTER=$2 timeout 200 python tools/train_step_time.py 2>&1 | grep "backward as a CUDA graph\|grad norm\|rror"
done 2>&1 | tee gpurun_out/${TAG}.txt
