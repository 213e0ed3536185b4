#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== parity (small)"
  timeout 600 python -m pytest tests -m gpu -x -q -k "not c3 and not c2" 2>&1 | tail -3
  echo "== A/B"
  OAR_EM_LIB=$V/liboarfish_em_rev_a1dc74b.so timeout 150 python tools/dev/ab.py C3 base:5 2>&1 | tail -1
  timeout 150 python tools/dev/ab.py C3 new:5 2>&1 | tail -1
  for n in wcommon wparam wboth nocommon; do
    OAR_EM_LIB=$V/liboarfish_em_$n.so timeout 150 python tools/dev/ab.py C3 $n:5 2>&1 | tail -1
  done
  OAR_EM_LIB=$V/liboarfish_em_rev_a1dc74b.so timeout 150 python tools/dev/ab.py C3 base:5 2>&1 | tail -1
  timeout 150 python tools/dev/ab.py C3 new:5 2>&1 | tail -1
} 2>&1 | tee gpurun_out/call24.log
