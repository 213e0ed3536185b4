#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  timeout 100 python -m pytest tests -m gpu -x -q -k layout_positions 2>&1 | tail -1
  for i in 1 2; do
  timeout 200 python tools/dev/sustained.py "6 CTAs, side" 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_cta5.so timeout 200 python tools/dev/sustained.py "5 CTAs (48 regs), side" 2>&1 | tail -1
  OAR_UPDATE_MODE=serial timeout 200 python tools/dev/sustained.py "6 CTAs, serial" 2>&1 | tail -1
  OAR_UPDATE_MODE=serial OAR_EM_LIB=$V/liboarfish_em_cta5.so timeout 200 python tools/dev/sustained.py "5 CTAs (48 regs), serial" 2>&1 | tail -1
  done
} 2>&1 | tee gpurun_out/call40.log
