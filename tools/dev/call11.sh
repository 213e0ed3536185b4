#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== A/B"
  timeout 200 python tools/dev/ab.py C3 product:5 product:5 2>&1 | tail -2
  timeout 100 python tools/dev/ab.py C2 product:5 2>&1 | tail -1
  echo "== parity"
  timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
  bash tools/measure_round.sh
} 2>&1 | tee gpurun_out/call11.log
