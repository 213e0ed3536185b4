#!/bin/bash
# N-GPU measurement: multi-GPU ABI tests, then the bench under torchrun at N = $1 (driver settings: --steps 20 --warmup 3).
#   gpurun --gpus N -- bash tools/bench_multi.sh N      -> gpurun_out/bench_nN.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
{
  nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -1
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== multi-GPU tests"
  timeout 400 python -m pytest tests -m gpu -x -q -k "multi or mirror or concurrent" 2>&1 | tail -3
  echo "== bench N=$N"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print({k:d[k] for k in ['value','replicates_per_sec','n_gpus','ms_per_step']}, d['job']['replicates_by_rank'], 'bcast', d['bcast_ms'], 'c5', d['c5']['cells_per_sec'] if d.get('c5') else None, d['clocks'])
PY
  tail -2 gpurun_out/bench_n$N.err
} 2>&1 | tee gpurun_out/bench_multi_n$N.log
