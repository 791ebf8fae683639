#!/bin/bash
# times every tuning variant built by tools/build_variants.sh at one launch shape (GPU box)
SHAPE=${SHAPE:-5:4}
for so in sqair_b200/csrc/exp_*.so; do
  n=$(basename $so .so)
  echo -n "$n: "; SQAIR_LIB=$PWD/$so SWEEP=$SHAPE python tools/sweep_rows.py 2>&1 | grep "ms/step\|profile tid 0\|Error\|error" | tail -3
done
