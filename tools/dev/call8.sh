#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  echo "== smoke (hang guard)"
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -3
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== create trace"
  OAR_TRACE=1 timeout 120 python tools/dev/build_prof.py C3 3 2>&1 | tail -4
  echo "== A/B C3: fused update on / off"
  timeout 200 python tools/dev/ab.py C3 fused:5 fused:5 2>&1 | tail -2
  OAR_FUSED_UPDATE=0 timeout 200 python tools/dev/ab.py C3 unfused:5 unfused:5 2>&1 | tail -2
  echo "== A/B C2"
  timeout 100 python tools/dev/ab.py C2 fused:5 fused:5 2>&1 | tail -1
  OAR_FUSED_UPDATE=0 timeout 100 python tools/dev/ab.py C2 unfused:5 unfused:5 2>&1 | tail -1
  echo "== parity suite (product lib)"
  timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
  echo "== parity, fused update off"
  OAR_FUSED_UPDATE=0 timeout 400 python -m pytest tests -m gpu -x -q -k "not c3 and not c2" 2>&1 | tail -3
  echo "== racecheck"
  timeout 200 compute-sanitizer --tool racecheck python tools/dev/sanit.py 2>&1 | grep -E "SUMMARY|^ok|hazard|Error" | head -4
  echo "== cells trace"
  OAR_TRACE=1 timeout 200 python tools/bench_cells.py 256 50000 200000 2>&1 | tail -4
} 2>&1 | tee gpurun_out/call8.log
