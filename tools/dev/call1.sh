#!/bin/bash
# Round 2, GPU call 1: A/B of every sweep variant on C3/C2, the full parity suite (incl. the new C2/C3 oracle tests and
# the variants round 1 never ran), racecheck, and one ncu --set full capture of the streaming sweep.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  echo "== A/B C3 (product lib)"
  timeout 400 python tools/dev/ab.py C3 2b:5 3:5 3:6 3:4 1c:4 1b:4 2>&1 | tail -8
  echo "== A/B C3 (scarce-first greedy lib)"
  OAR_EM_LIB=$PWD/oarfish_b200/lib/variants/liboarfish_em_scarce.so timeout 300 python tools/dev/ab.py C3 2b:5 3:5 2>&1 | tail -3
  echo "== A/B C2"
  timeout 200 python tools/dev/ab.py C2 2b:5 3:5 3:6 2>&1 | tail -4
  echo "== store creation with and without the overlapped upload"
  timeout 120 python tools/dev/build_prof.py C3 3 | tail -2
  OAR_UPLOAD_OVERLAP=1 timeout 120 python tools/dev/build_prof.py C3 3 | tail -2
  echo "== parity suite"
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
  echo "== racecheck"
  for sw in 3 1c; do
    OAR_SWEEP=$sw timeout 120 compute-sanitizer --tool racecheck python tools/dev/sanit.py 2>&1 | grep -E "SUMMARY|^ok|hazard|Error" | head -4
  done
  OAR_SWEEP=3 timeout 120 compute-sanitizer --tool memcheck python tools/dev/sanit.py 2>&1 | grep -E "SUMMARY|^ok|Error" | head -4
  echo "== ncu --set full, streaming sweep"
  OAR_SWEEP=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:em_sweep_tiled -s 50 -c 1 -f \
      -o gpurun_out/r2_sweep3 python tools/dev/prof.py C3 > gpurun_out/ncu_sweep3.log 2>&1
  tail -2 gpurun_out/ncu_sweep3.log
} 2>&1 | tee gpurun_out/call1.log
