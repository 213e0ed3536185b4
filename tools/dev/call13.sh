#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== A/B"
  OAR_EM_LIB=$V/liboarfish_em_rev_daf4a21.so timeout 150 python tools/dev/ab.py C3 rev_daf4a21:5 2>&1 | tail -1
  timeout 200 python tools/dev/ab.py C3 product:5 product:5 2>&1 | tail -2
  OAR_EM_LIB=$V/liboarfish_em_nocommon.so timeout 150 python tools/dev/ab.py C3 nocommon:5 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_noprologue.so OAR_FUSED_UPDATE=0 timeout 150 python tools/dev/ab.py C3 noprologue:5 2>&1 | tail -1
  echo "== ncu plain / weighted (product)"
  timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:em_sweep_tiled -s 5 -c 1 -f \
      -o gpurun_out/sweep_plain python tools/dev/prof.py C3 > gpurun_out/ncu_plain.log 2>&1; tail -1 gpurun_out/ncu_plain.log
  timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:em_sweep_tiled -s 5 -c 1 -f \
      -o gpurun_out/sweep_weighted python tools/dev/prof_w.py C3 > gpurun_out/ncu_weighted.log 2>&1; tail -1 gpurun_out/ncu_weighted.log
  OAR_EM_LIB=$V/liboarfish_em_rev_daf4a21.so timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:em_sweep_tiled -s 5 -c 1 -f \
      -o gpurun_out/sweep_weighted_daf4a21 python tools/dev/prof_w.py C3 > gpurun_out/ncu_weighted_d.log 2>&1; tail -1 gpurun_out/ncu_weighted_d.log
} 2>&1 | tee gpurun_out/call13.log
