#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  OAR_TRACE=1 timeout 150 python tools/dev/ab.py C3 new:5 new:5 2>&1 | grep -v "^\[oar\] cells" | tail -4
  timeout 100 python tools/dev/dump_lpos.py C3 2>&1 | tail -1
} 2>&1 | tee gpurun_out/call27.log
