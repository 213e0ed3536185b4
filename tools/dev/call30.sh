#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  OAR_TRACE=1 timeout 150 python tools/dev/ab.py C3 new:5 new:5 2>&1 | grep -v "^\[oar\] cells" | tail -4
  OAR_EM_LIB=$V/liboarfish_em_cta6.so timeout 150 python tools/dev/ab.py C3 cta6:6 cta6:5 2>&1 | tail -2
  timeout 100 python tools/dev/dump_lpos.py C3 2>&1 | tail -1
  echo "== parity (small)"
  timeout 600 python -m pytest tests -m gpu -x -q -k "not c3 and not c2" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/call30.log
