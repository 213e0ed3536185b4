#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  timeout 150 python tools/dev/ab.py C3 default:0 five:5 default:0 2>&1 | tail -3
  timeout 150 python tools/dev/ab.py C2 default:0 five:5 default:0 2>&1 | tail -3
  echo "== full GPU suite"
  timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
} 2>&1 | tee gpurun_out/call31.log
