mkdir -p gpurun_out
{
timeout 300 python tests/_lane_check.py check | tail -n 4
echo "== timing default"
timeout 300 python tests/_lane_check.py time C3 10
V=$PWD/oarfish_b200/lib/variants
for v in w1:20 w4:5 r12:10 c12:12 w1c24:24; do
  n=${v%%:*}; c=${v##*:}
  echo "== variant $n"
  OAR_EM_LIB=$V/liboarfish_em_$n.so timeout 300 python tests/_lane_check.py time C3 $c
done
echo "== C2"
timeout 300 python tests/_lane_check.py time C2 10
echo "== pytest"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:em_sweep_lane -s 3 -c 1 -o gpurun_out/prof_lane3 -f python tests/_prof2.py C3 > gpurun_out/prof_lane3.log 2>&1
} > gpurun_out/lane4.log 2>&1
tail -n 30 gpurun_out/lane4.log
