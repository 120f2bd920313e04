#!/bin/bash
# A/B on one box: training-step kernel times as built vs rebuilt with an extra nvcc flag, e.g.
#   bash tools/ab_build_flag.sh -DNERF_INTERLEAVE_STORES=0        bash tools/ab_build_flag.sh -DNERF_STREAMING_STORES=0
cd "$(dirname "$0")/.."
run() { python tools/profile_train_step.py 2>&1 | grep -v Warning | grep "total\|wgrad\|mlp_tc" | head -4; }
echo "== as built"; run; run
touch cse-573-minimal-nerf_b200/csrc/mlp_tc3.cu cse-573-minimal-nerf_b200/csrc/mlp_tc_bwd3.cu
make -C cse-573-minimal-nerf_b200/csrc EXTRA="$1" > /dev/null 2>&1
echo "== rebuilt with $1"; run; run
