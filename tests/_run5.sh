mkdir -p gpurun_out
{
echo "== checks (chunk layout)"
timeout 300 python tests/_lane_check.py check | tail -n 8
echo "== timing chunk"
timeout 300 python tests/_lane_check.py time C3 5
timeout 300 python tests/_lane_check.py time C2 5
echo "== pytest"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6
} > gpurun_out/chunk_items.log 2>&1
cat gpurun_out/chunk_items.log
