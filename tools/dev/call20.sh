#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  timeout 600 python -m pytest tests -m gpu -x -q -k "not c3 and not c2" 2>&1 | tail -5
  timeout 100 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "filtered" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" | tail -3
} 2>&1 | tee gpurun_out/call20.log
