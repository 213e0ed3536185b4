#!/bin/bash
# Dev helper: builds tuning variants of liboarfish_em.so into oarfish_b200/lib/variants/ (git-ignored, travels with gpurun).
# usage: tools/dev/build_variants.sh name "-DOAR_LANE_THREADS=128 -DOAR_LANE_MIN_CTAS=6" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/../.."
mkdir -p oarfish_b200/lib/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v $flags \
    -shared -o oarfish_b200/lib/variants/liboarfish_em_$name.so oarfish_b200/csrc/*.cu 2> oarfish_b200/lib/variants/$name.ptxas.log &
done
wait
for f in oarfish_b200/lib/variants/*.ptxas.log; do echo "== $f"; grep -A2 "em_sweep_laneILb0ELb0" $f | grep "spill\|Used"; done
