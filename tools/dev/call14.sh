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
  OAR_EM_LIB=$V/liboarfish_em_noxsparam.so timeout 150 python tools/dev/ab.py C3 noxsparam:5 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_noxsparam_nocommon.so timeout 150 python tools/dev/ab.py C3 noxsparam_nocommon:5 2>&1 | tail -1
} 2>&1 | tee gpurun_out/call14.log
