#!/bin/bash
# Round measurements on the GPU box (everything lands in gpurun_out/):
#   bench lines (both arms), ncu launch list of the bench command, ncu --set full of one steady-state sweep
#   (both sweep variants), compute-sanitizer runs of tools/dev/sanit.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for sw in 2b 1b; do
  OAR_SWEEP=$sw timeout 200 ncu --set full --clock-control none --import-source on -k regex:em_sweep_tiled -s 50 -c 1 -f \
      -o gpurun_out/sweep_$sw python tools/dev/prof.py C3 > gpurun_out/ncu_$sw.log 2>&1
done
if [ "$1" != "nosan" ]; then
for cfg in "2b racecheck" "1b racecheck" "2b memcheck" "1b memcheck" "2b synccheck" "1b synccheck"; do
  set -- $cfg
  echo "== OAR_SWEEP=$1 --tool $2" >> gpurun_out/sanitizer.log
  OAR_SWEEP=$1 timeout 170 compute-sanitizer --tool $2 python tools/dev/sanit.py 2>&1 | grep -E "SUMMARY|^ok|Error|hazard" | head -8 >> gpurun_out/sanitizer.log
done
fi
cat gpurun_out/sanitizer.log 2>/dev/null
