#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu --format=csv,noheader
  for lib in rev_48a20b9 rev_daf4a21 noprologue; do
    OAR_EM_LIB=$V/liboarfish_em_$lib.so OAR_FUSED_UPDATE=0 timeout 150 python tools/dev/ab.py C3 $lib:5 2>&1 | tail -1
  done
  timeout 200 python tools/dev/ab.py C3 product:5 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_rev_48a20b9.so timeout 150 python tools/dev/ab.py C3 rev_48a20b9:5 2>&1 | tail -1
} 2>&1 | tee gpurun_out/call10.log
