#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  for i in 1 2; do
  timeout 150 python tools/dev/ab.py C3 head:0 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_rev_HEAD~2.so timeout 150 python tools/dev/ab.py C3 before:0 2>&1 | tail -1
  done
  echo "== --tool racecheck"
  timeout 250 compute-sanitizer --tool racecheck python tools/dev/sanit.py 2>&1 | grep -E "SUMMARY|^ok|Error|Warning|hazard" | sort | uniq -c | sort -rn | head -12
  echo "== parity (small)"
  timeout 600 python -m pytest tests -m gpu -x -q -k "not c3 and not c2" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/call36.log
