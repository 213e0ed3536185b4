mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:em_sweep_lane -s 3 -c 1 -o gpurun_out/prof_lane2 -f python tests/_prof2.py C3 > gpurun_out/prof_lane2.log 2>&1
tail -n 3 gpurun_out/prof_lane2.log
