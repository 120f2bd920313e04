#!/bin/bash
# A/B on one box: training-step kernel times as built vs rebuilt with an extra nvcc flag, e.g.
#   bash tools/ab_build_flag.sh -DNERF_INTERLEAVE_STORES=0        bash tools/ab_build_flag.sh -DNERF_STREAMING_STORES=0
cd "$(dirname "$0")/.."
run() { python tools/profile_train_step.py 2>&1 | grep -v Warning | grep "total\|wgrad\|mlp_tc" | head -4; }
echo "== as built"; run; run
for flag in "$@"; do
    touch cse-573-minimal-nerf_b200/csrc/*.cu
    make -C cse-573-minimal-nerf_b200/csrc EXTRA="$flag" > /dev/null 2>&1
    echo "== rebuilt with $flag"; run; run
done
