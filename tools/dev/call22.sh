#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== contiguous vs strided walk"
  OAR_EM_LIB=$V/liboarfish_em_strided.so timeout 150 python tools/dev/ab.py C3 strided:5 2>&1 | tail -1
  timeout 150 python tools/dev/ab.py C3 contiguous:5 contiguous:5 2>&1 | tail -2
  timeout 100 python tools/dev/ab.py C2 contiguous:5 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_strided.so timeout 100 python tools/dev/ab.py C2 strided:5 2>&1 | tail -1
  echo "== parity"
  timeout 600 python -m pytest tests -m gpu -x -q -k "not c3 and not c2" 2>&1 | tail -2
} 2>&1 | tee gpurun_out/call22.log
