#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== full GPU suite"
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  echo "== bench N=1"
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
  echo "== bench defaults (no flags)"
  timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; head -c 200 gpurun_out/bench_default.json; echo
  echo "== f-rows"
  timeout 400 python tools/bench_frows.py C3 2>/dev/null | tail -1 > gpurun_out/frows.json; tail -c 700 gpurun_out/frows.json; echo
} 2>&1 | tee gpurun_out/final_n1.log
