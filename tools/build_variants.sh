#!/bin/bash
# builds tuning variants of the library (R=5 only) in parallel: name=flags ...
cd sqair_b200/csrc
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr \
      -DSQAIR_ONLY_R=${EXP_R:-5} $flags -shared -o exp_${name}.so sqair_api.cu sqair_train.cu sqair_wgrad_tc.cu 2> exp_${name}.log
    python -c "import ctypes; ctypes.CDLL('./exp_${name}.so')" || echo "$name: DOES NOT LOAD"
    echo "$name: $(grep -A1 'sqair_sequence_kernel' exp_${name}.log | grep -E 'spill' | head -1) $(grep -A2 'sqair_sequence_kernel' exp_${name}.log | grep 'Used' | head -1)" ) &
done
wait
