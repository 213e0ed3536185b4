#!/bin/bash
# 2-GPU call: the multi-GPU ABI, the torchrun bench of both arms
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
V=$PWD/oarfish_b200/lib/variants
{
  nvidia-smi --query-gpu=index,name --format=csv,noheader
  echo "== smoke (hang guard)"
  timeout 150 python __graft_entry__.py smoke 2>&1 | tail -3
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
  echo "== create trace: product / noscarce"
  OAR_TRACE=1 timeout 120 python tools/dev/build_prof.py C3 3 2>&1 | tail -6
  OAR_TRACE=1 OAR_EM_LIB=$V/liboarfish_em_noscarce.so timeout 120 python tools/dev/build_prof.py C3 3 2>&1 | tail -6
  echo "== multi-GPU tests"
  timeout 400 python -m pytest tests -m gpu -x -q -k "multi or mirror or concurrent or progress or reference_dumps or batched_cells or binomial" 2>&1 | tail -6
  echo "== bench N=2 (20 replicates)"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
  tail -c 2500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
  echo "== reference arm under torchrun N=2"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
  tail -c 1200 gpurun_out/bench_ref_n2.json; tail -3 gpurun_out/bench_ref_n2.err
} 2>&1 | tee gpurun_out/call7.log
