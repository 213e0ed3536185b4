#!/bin/bash
# Dev helper (GPU box): A/B timing of sweep variants on C3 with tools/dev/micro2.py.
# usage: tools/dev/ab.sh "<lib|-> <OAR_SWEEP> <ctas/SM> [OAR_TILE_SPAN]" ...     (lib = name under oarfish_b200/lib/variants, - = the product library)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  lib=$1; sw=$2; cps=$3; span=$4
  if [ "$lib" = "-" ]; then unset OAR_EM_LIB; else export OAR_EM_LIB=$PWD/oarfish_b200/lib/variants/liboarfish_em_$lib.so; fi
  echo "== lib=$lib sweep=$sw span=${span:-default}" | tee -a gpurun_out/ab.log
  OAR_TILE_SPAN=$span OAR_SWEEP=$sw timeout 240 python tools/dev/micro2.py C3 $cps 2>&1 | tail -4 | tee -a gpurun_out/ab.log
done
