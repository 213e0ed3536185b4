#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  for i in 1 2; do
  echo "new, policy as is"; timeout 150 python tools/dev/ab.py C3 new:0 2>&1 | tail -1
  echo "new, fused weighted everywhere"; OAR_FUSED_WTS_MAX_TILES=10000000 timeout 150 python tools/dev/ab.py C3 newfw:0 2>&1 | tail -1
  echo "HEAD"; OAR_EM_LIB=$V/liboarfish_em_rev_HEAD.so timeout 150 python tools/dev/ab.py C3 head:0 2>&1 | tail -1
  done
  timeout 100 python tools/dev/ab.py C2 new:0 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_rev_HEAD.so timeout 100 python tools/dev/ab.py C2 head:0 2>&1 | tail -1
  echo "== parity (small)"
  timeout 600 python -m pytest tests -m gpu -x -q -k "not c3 and not c2" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/call37.log
