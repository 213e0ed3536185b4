#!/bin/bash
# Dev helper: builds tuning variants of liboarfish_em.so into oarfish_b200/lib/variants/ (git-ignored, travels with gpurun).
# usage: tools/dev/build_variants.sh name "-DOAR_GREEDY_SCARCE=0" [name2 "flags2" ...]
#        a name of the form rev:<git rev> builds that revision's sources (flags still apply)
set -e
cd "$(dirname "$0")/../.."
mkdir -p oarfish_b200/lib/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  src=oarfish_b200/csrc
  case $name in rev:*)
    rev=${name#rev:}; name=rev_$rev; src=/tmp/oar_variant_$rev/oarfish_b200/csrc
    rm -rf /tmp/oar_variant_$rev; mkdir -p /tmp/oar_variant_$rev
    git archive $rev oarfish_b200/csrc include | tar -x -C /tmp/oar_variant_$rev ;;
  esac
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v $flags \
    -shared -o oarfish_b200/lib/variants/liboarfish_em_$name.so $src/*.cu 2> oarfish_b200/lib/variants/$name.ptxas.log &
done
wait
for f in oarfish_b200/lib/variants/*.ptxas.log; do echo "== $f"; grep -A2 "em_sweep_tiledILb0ELb" $f | grep "spill\|Used"; done
