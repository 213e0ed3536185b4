#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv,noheader
  for i in 1 2; do
  timeout 150 python tools/dev/ab.py C3 head:0 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_rev_HEAD~2.so timeout 150 python tools/dev/ab.py C3 before:0 2>&1 | tail -1
  done
} 2>&1 | tee gpurun_out/call35.log
