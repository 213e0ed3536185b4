#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== A/B C3 / C2 (fused weighted variants matter on C2 only)"
  timeout 150 python tools/dev/ab.py C3 new:5 2>&1 | tail -1
  timeout 100 python tools/dev/ab.py C2 new:5 2>&1 | tail -1
  for n in fw_both fw_common; do
    OAR_EM_LIB=$V/liboarfish_em_$n.so timeout 100 python tools/dev/ab.py C2 $n:5 2>&1 | tail -1
  done
  timeout 100 python tools/dev/ab.py C2 new:5 2>&1 | tail -1
  echo "== full GPU suite"
  timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  echo "== ncu --set full, one steady-state sweep, plain and weighted"
  timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:em_sweep_tiled -s 5 -c 1 -f \
      -o gpurun_out/sweep_plain python tools/dev/prof.py C3 > gpurun_out/ncu_plain.log 2>&1; tail -1 gpurun_out/ncu_plain.log
  timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:em_sweep_tiled -s 5 -c 1 -f \
      -o gpurun_out/sweep_weighted python tools/dev/prof_w.py C3 > gpurun_out/ncu_weighted.log 2>&1; tail -1 gpurun_out/ncu_weighted.log
} 2>&1 | tee gpurun_out/call25.log
