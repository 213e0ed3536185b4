#!/bin/bash
# Round 2, GPU call 2: lane-parallel phase 2 in the streaming sweep (A/B against the thread-per-item version), prefetch depths,
# parity of the new code, destroy trace, ncu of the new kernel.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  echo "== A/B C3 product lib (P2Q=1)"
  timeout 300 python tools/dev/ab.py C3 3:5 3:4 2b:5 2>&1 | tail -4
  echo "== P2Q=0"
  OAR_EM_LIB=$V/liboarfish_em_p2q0.so timeout 200 python tools/dev/ab.py C3 3:5 2>&1 | tail -2
  echo "== L2_AHEAD=2"
  OAR_EM_LIB=$V/liboarfish_em_l2a2.so timeout 200 python tools/dev/ab.py C3 3:5 2>&1 | tail -2
  echo "== L2_AHEAD=6 REC_AHEAD=6"
  OAR_EM_LIB=$V/liboarfish_em_l2a6.so timeout 200 python tools/dev/ab.py C3 3:5 2>&1 | tail -2
  echo "== parity (kernel 6 = OAR_SWEEP=3 and the full-size tests)"
  OAR_SWEEP=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "6 or c2 or c3 or cells or coverage" 2>&1 | tail -4
  echo "== destroy trace"
  OAR_TRACE=1 timeout 200 python tools/dev/e2e_prof.py C3 2>&1 | tail -12
  echo "== racecheck"
  OAR_SWEEP=3 timeout 120 compute-sanitizer --tool racecheck python tools/dev/sanit.py 2>&1 | grep -E "SUMMARY|^ok|hazard|Error" | head -4
  echo "== ncu"
  OAR_SWEEP=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:em_sweep_tiled -s 50 -c 1 -f \
      -o gpurun_out/r2_sweep3q python tools/dev/prof.py C3 > gpurun_out/ncu_sweep3q.log 2>&1
  tail -2 gpurun_out/ncu_sweep3q.log
} 2>&1 | tee gpurun_out/call2.log
