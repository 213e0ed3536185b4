#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== A/B C3"
  OAR_TRACE=1 timeout 150 python tools/dev/ab.py C3 new:5 2>&1 | grep -v "^\[oar\] cells" | tail -3
  OAR_EM_LIB=$V/liboarfish_em_rev_HEAD~1.so OAR_TRACE=1 timeout 150 python tools/dev/ab.py C3 prev:5 2>&1 | tail -3
  timeout 150 python tools/dev/ab.py C3 new:5 2>&1 | tail -1
  timeout 100 python tools/dev/ab.py C2 new:5 2>&1 | tail -1
  timeout 100 python tools/dev/dump_lpos.py C3 2>&1 | tail -1
  echo "== parity (small)"
  timeout 600 python -m pytest tests -m gpu -x -q -k "not c3 and not c2" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/call26.log
