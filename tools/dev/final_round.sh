#!/bin/bash
# Dev helper (GPU box): the whole GPU suite, then tools/measure_round.sh
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== full GPU suite"
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  bash tools/measure_round.sh
} 2>&1 | tee gpurun_out/final_round.log
