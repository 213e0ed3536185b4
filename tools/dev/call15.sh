#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== A/B C3"
  OAR_EM_LIB=$V/liboarfish_em_rev_daf4a21.so timeout 150 python tools/dev/ab.py C3 rev_daf4a21:5 2>&1 | tail -1
  timeout 200 python tools/dev/ab.py C3 fused:5 2>&1 | tail -1
  OAR_FUSED_UPDATE=0 timeout 200 python tools/dev/ab.py C3 unfused:5 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_nocommonw.so timeout 150 python tools/dev/ab.py C3 nocommonw_fused:5 2>&1 | tail -1
  OAR_FUSED_UPDATE=0 OAR_EM_LIB=$V/liboarfish_em_nocommonw.so timeout 150 python tools/dev/ab.py C3 nocommonw_unfused:5 2>&1 | tail -1
  echo "== A/B C2"
  timeout 100 python tools/dev/ab.py C2 fused:5 2>&1 | tail -1
  OAR_FUSED_UPDATE=0 timeout 100 python tools/dev/ab.py C2 unfused:5 2>&1 | tail -1
} 2>&1 | tee gpurun_out/call15.log
