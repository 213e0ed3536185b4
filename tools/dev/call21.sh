#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== 6 CTAs/SM (40 registers)"
  OAR_CTAS_PER_SM=6 OAR_EM_LIB=$V/liboarfish_em_cta6.so timeout 150 python tools/dev/ab.py C3 cta6:6 2>&1 | tail -1
  OAR_EM_LIB=$V/liboarfish_em_cta6.so timeout 150 python tools/dev/ab.py C3 cta6_5:5 2>&1 | tail -1
  echo "== bench C5 only"
  timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['c5'])"
} 2>&1 | tee gpurun_out/call21.log
