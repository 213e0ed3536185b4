mkdir -p gpurun_out
{
echo "== pytest gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "== bench n1"
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json
echo "== bench lane layout"
OAR_LAYOUT=lane timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1_lane.json 2> gpurun_out/bench_n1_lane.err; tail -c 1500 gpurun_out/bench_n1_lane.json
echo "== bench reference"
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
} > gpurun_out/final1.log 2>&1
tail -n 40 gpurun_out/final1.log
