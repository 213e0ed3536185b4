mkdir -p gpurun_out
{
timeout 300 python tests/_lane_check.py check | tail -n 2
timeout 300 python tests/_lane_check.py time C3 5
timeout 300 python tests/_lane_check.py time C2 5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
} > gpurun_out/chunk_items3.log 2>&1
cat gpurun_out/chunk_items3.log
