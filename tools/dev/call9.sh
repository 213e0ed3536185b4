#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  echo "== smoke (hang guard)"
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv,noheader
  echo "== A/B C3"
  timeout 200 python tools/dev/ab.py C3 product:5 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_noscarce.so OAR_FUSED_UPDATE=0 timeout 150 python tools/dev/ab.py C3 noscarce:5 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_noprologue.so OAR_FUSED_UPDATE=0 timeout 150 python tools/dev/ab.py C3 noprologue:5 2>&1 | tail -1
  timeout 200 python tools/dev/ab.py C3 product:5 2>&1 | tail -1
  nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv,noheader
  echo "== ncu plain"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:em_sweep_tiled -s 50 -c 1 -f \
      -o gpurun_out/r2_sweep_plain3 python tools/dev/prof.py C3 > gpurun_out/ncu_plain3.log 2>&1; tail -2 gpurun_out/ncu_plain3.log
} 2>&1 | tee gpurun_out/call9.log
