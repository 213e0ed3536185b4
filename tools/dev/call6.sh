#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  echo "== smoke (hang guard)"
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -3
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== A/B C3: split prev[] table"
  OAR_EM_LIB=$V/liboarfish_em_nosplit.so timeout 150 python tools/dev/ab.py C3 nosplit:5 2>&1 | tail -1
  timeout 200 python tools/dev/ab.py C3 new:5 new:5 2>&1 | tail -2
  echo "== A/B C2"
  OAR_EM_LIB=$V/liboarfish_em_nosplit.so timeout 100 python tools/dev/ab.py C2 nosplit:5 2>&1 | tail -1
  timeout 100 python tools/dev/ab.py C2 new:5 2>&1 | tail -1
  echo "== create time"
  timeout 120 python tools/dev/build_prof.py C3 4 2>&1 | tail -4
  echo "== robustness"
  timeout 400 python tools/bench_robust.py C3 2>&1 | tail -3 | tee gpurun_out/robust.jsonl
  echo "== robustness with clustering off (permuted only matters)"
  OAR_CLUSTER_IDS=0 timeout 400 python tools/bench_robust.py C3 2>&1 | tail -3 | tee gpurun_out/robust_nocluster.jsonl
  echo "== parity suite (product lib)"
  timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
  echo "== parity with clustering forced on"
  OAR_CLUSTER_IDS=1 timeout 300 python -m pytest tests -m gpu -x -q -k "matches_oracle or golden or edge or cells" 2>&1 | tail -3
  echo "== racecheck"
  timeout 200 compute-sanitizer --tool racecheck python tools/dev/sanit.py 2>&1 | grep -E "SUMMARY|^ok|hazard|Error" | head -4
  echo "== ncu plain"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:em_sweep_tiled -s 50 -c 1 -f \
      -o gpurun_out/r2_sweep_plain2 python tools/dev/prof.py C3 > gpurun_out/ncu_plain2.log 2>&1; tail -2 gpurun_out/ncu_plain2.log
} 2>&1 | tee gpurun_out/call6.log
