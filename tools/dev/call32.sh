#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "A lean weighted, 5 CTAs"; timeout 150 python tools/dev/ab.py C3 A:0 2>&1 | tail -1
  echo "B fused weighted, 5 CTAs"; OAR_FUSED_WTS_MAX_TILES=10000000 timeout 150 python tools/dev/ab.py C3 B:0 2>&1 | tail -1
  echo "C lean weighted, 6 CTAs"; OAR_EM_LIB=$V/liboarfish_em_w6.so timeout 150 python tools/dev/ab.py C3 C:0 2>&1 | tail -1
  echo "D fused weighted, 6 CTAs"; OAR_FUSED_WTS_MAX_TILES=10000000 OAR_EM_LIB=$V/liboarfish_em_w6.so timeout 150 python tools/dev/ab.py C3 D:0 2>&1 | tail -1
  echo "A lean weighted, 5 CTAs"; timeout 150 python tools/dev/ab.py C3 A:0 2>&1 | tail -1
} 2>&1 | tee gpurun_out/call32.log
