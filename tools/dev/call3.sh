#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  echo "== A/B C3: early-gather two-barrier sweep (2e) vs 2b"
  timeout 300 python tools/dev/ab.py C3 2e:5 2b:5 2e:4 2>&1 | tail -4
  echo "== C2"
  timeout 200 python tools/dev/ab.py C2 2b:5 2e:5 2b:5 2e:5 2>&1 | tail -4
  echo "== parity kernel 7"
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "7" 2>&1 | tail -4
  OAR_SWEEP=2e timeout 120 compute-sanitizer --tool racecheck python tools/dev/sanit.py 2>&1 | grep -E "SUMMARY|^ok|hazard|Error" | head -4
  echo "== ncu"
  OAR_SWEEP=2e timeout 300 ncu --set full --clock-control none --import-source on -k regex:em_sweep_tiled -s 50 -c 1 -f \
      -o gpurun_out/r2_sweep2e python tools/dev/prof.py C3 > gpurun_out/ncu_sweep2e.log 2>&1
  tail -2 gpurun_out/ncu_sweep2e.log
} 2>&1 | tee gpurun_out/call3.log
