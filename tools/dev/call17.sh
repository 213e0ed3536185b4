#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== cells: sparse (5k expressed) and dense"
  OAR_TRACE=1 timeout 200 python tools/bench_cells.py 256 50000 200000 5000 2>&1 | grep -v "layout:" | tail -4
  OAR_TRACE=1 timeout 200 python tools/bench_cells.py 256 50000 200000 0 2>&1 | grep -v "layout:" | tail -4
  echo "== weighted variants"
  OAR_EM_LIB=$V/liboarfish_em_wcommon.so timeout 150 python tools/dev/ab.py C3 wcommon:5 2>&1 | tail -1
  timeout 150 python tools/dev/ab.py C3 product:5 product:5:992 product:5:1000 product:5:976 2>&1 | tail -4
} 2>&1 | tee gpurun_out/call17.log
