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
  timeout 100 python tools/dev/ab.py C2 product:5 2>&1 | tail -1
  echo "== parity"
  timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
  echo "== ncu"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:em_sweep_tiled -s 40 -c 1 -f \
      -o gpurun_out/r2_sweep_plain4 python tools/dev/prof.py C3 > gpurun_out/ncu_plain4.log 2>&1; tail -2 gpurun_out/ncu_plain4.log
} 2>&1 | tee gpurun_out/call12.log
