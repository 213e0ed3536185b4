#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== parity (small)"
  timeout 600 python -m pytest tests -m gpu -x -q -k "not c3 and not c2" 2>&1 | tail -3
  timeout 200 python tools/dev/em_time.py fused-bookctas | tail -2
  OAR_FUSED_UPDATE=0 timeout 200 python tools/dev/em_time.py serial | tail -2
  timeout 200 python tools/dev/sustained.py "weighted: lean + em_update" 6 400 2>&1 | tail -1
  OAR_FUSED_WTS_MAX_TILES=10000000 timeout 200 python tools/dev/sustained.py "weighted: fused book CTAs" 6 400 2>&1 | tail -1
  timeout 200 python tools/dev/sustained.py "weighted: lean + em_update" 6 400 2>&1 | tail -1
  OAR_FUSED_WTS_MAX_TILES=10000000 timeout 200 python tools/dev/sustained.py "weighted: fused book CTAs" 6 400 2>&1 | tail -1
  timeout 100 python tools/dev/ab.py C2 book:0 2>&1 | tail -1
  timeout 150 python tools/dev/ab.py C3 book:0 2>&1 | tail -1
} 2>&1 | tee gpurun_out/call42.log
