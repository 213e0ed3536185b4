#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  OAR_UPDATE_MODE=side timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  echo "== parity (small), side mode forced"
  OAR_UPDATE_MODE=side timeout 600 python -m pytest tests -m gpu -x -q -k "not c3 and not c2" 2>&1 | tail -3
  for m in side fused serial side; do
    echo "C3 mode $m"; OAR_UPDATE_MODE=$m timeout 150 python tools/dev/ab.py C3 $m:0 2>&1 | tail -1
  done
  for m in side fused serial; do
    echo "C2 mode $m"; OAR_UPDATE_MODE=$m timeout 150 python tools/dev/ab.py C2 $m:0 2>&1 | tail -1
  done
} 2>&1 | tee gpurun_out/call38.log
