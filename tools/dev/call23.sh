#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== fused weighted instantiation: base / common path"
  timeout 150 python tools/dev/ab.py C3 product:5 2>&1 | tail -1
  for v in fw_param fw_common fw_both; do
    OAR_EM_LIB=$V/liboarfish_em_$v.so timeout 150 python tools/dev/ab.py C3 $v:5 2>&1 | tail -1
  done
  OAR_FUSED_UPDATE=0 timeout 150 python tools/dev/ab.py C3 unfused:5 2>&1 | tail -1
  timeout 150 python tools/dev/ab.py C3 product:5 2>&1 | tail -1
} 2>&1 | tee gpurun_out/call23.log
