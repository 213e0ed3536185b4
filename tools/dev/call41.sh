#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  for i in 1 2; do
  timeout 200 python tools/dev/sustained.py "side hi-prio 128" 6 400 2>&1 | tail -1
  OAR_UPDATE_THREADS=64 timeout 200 python tools/dev/sustained.py "side hi-prio 64" 6 400 2>&1 | tail -1
  OAR_UPDATE_THREADS=32 timeout 200 python tools/dev/sustained.py "side hi-prio 32" 6 400 2>&1 | tail -1
  OAR_UPDATE_MODE=serial timeout 200 python tools/dev/sustained.py "serial" 6 400 2>&1 | tail -1
  done
} 2>&1 | tee gpurun_out/call41.log
