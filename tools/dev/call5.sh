#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  echo "== smoke (hang guard)"
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -3
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== A/B C3"
  for lib in rev_HEAD noscarce noscan; do
    OAR_EM_LIB=$V/liboarfish_em_$lib.so timeout 150 python tools/dev/ab.py C3 $lib:5 2>&1 | tail -1
  done
  timeout 200 python tools/dev/ab.py C3 new:5 new:5 2>&1 | tail -2
  echo "== A/B C2"
  OAR_EM_LIB=$V/liboarfish_em_rev_HEAD.so timeout 100 python tools/dev/ab.py C2 rev_HEAD:5 2>&1 | tail -1
  timeout 100 python tools/dev/ab.py C2 new:5 2>&1 | tail -1
  echo "== parity suite (product lib)"
  timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
  echo "== racecheck"
  timeout 200 compute-sanitizer --tool racecheck python tools/dev/sanit.py 2>&1 | grep -E "SUMMARY|^ok|hazard|Error" | head -4
  echo "== ncu plain + weighted"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:em_sweep_tiled -s 50 -c 1 -f \
      -o gpurun_out/r2_sweep_plain python tools/dev/prof.py C3 > gpurun_out/ncu_plain.log 2>&1; tail -2 gpurun_out/ncu_plain.log
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:em_sweep_tiled -s 50 -c 1 -f \
      -o gpurun_out/r2_sweep_weighted python tools/dev/prof_w.py C3 > gpurun_out/ncu_weighted.log 2>&1; tail -2 gpurun_out/ncu_weighted.log
  echo "== f-rows"
  timeout 200 python tools/bench_frows.py C3 2>&1 | tail -1 | tee gpurun_out/frows.json
  echo "== bench (short)"
  timeout 500 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; tail -c 4000 gpurun_out/bench_short.json; tail -3 gpurun_out/bench_short.err
} 2>&1 | tee gpurun_out/call5.log
