#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  timeout 150 python tools/dev/ab.py C3 align128:5 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_align16.so timeout 150 python tools/dev/ab.py C3 align16:5 2>&1 | tail -1
  timeout 150 python tools/dev/ab.py C3 align128:5 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_align16.so timeout 150 python tools/dev/ab.py C3 align16:5 2>&1 | tail -1
} 2>&1 | tee gpurun_out/call29.log
