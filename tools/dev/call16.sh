#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== A/B C3"
  timeout 200 python tools/dev/ab.py C3 fused:5 2>&1 | tail -1
  OAR_FUSED_UPDATE=0 timeout 200 python tools/dev/ab.py C3 unfused:5 2>&1 | tail -1
  echo "== A/B C2"
  timeout 100 python tools/dev/ab.py C2 fused:5 2>&1 | tail -1
  OAR_FUSED_UPDATE=0 timeout 100 python tools/dev/ab.py C2 unfused:5 2>&1 | tail -1
  echo "== parity fused / unfused"
  timeout 400 python -m pytest tests -m gpu -x -q -k "not c3 and not c2" 2>&1 | tail -2
  OAR_FUSED_UPDATE=0 timeout 400 python -m pytest tests -m gpu -x -q -k "not c3 and not c2" 2>&1 | tail -2
} 2>&1 | tee gpurun_out/call16.log
