#!/bin/bash
# Round measurements on the GPU box (everything lands in gpurun_out/; tools/summarize_profiles.py turns it into profiles/):
#   bench lines of both arms, ncu launch list of the bench command, ncu --set full of one steady-state sweep (plain and
#   bootstrap-weighted), compute-sanitizer runs of tools/dev/sanit.py, the f-row and robustness tools.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== smoke (hang guard)"
timeout 150 python __graft_entry__.py smoke 2>&1 | tail -2
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
echo "== bench, N = 1 (driver defaults: --steps 20 is BASELINE config 4, 100 replicates)"
timeout 900 python bench.py --steps ${STEPS:-20} --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c2 --no-c5 > gpurun_out/bench_under_ncu.log 2>&1
tail -c 200 gpurun_out/bench_under_ncu.log
echo "== ncu --set full, one steady-state sweep, plain and weighted"
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:em_sweep_tiled -s 5 -c 1 -f \
    -o gpurun_out/sweep_plain python tools/dev/prof.py C3 > gpurun_out/ncu_plain.log 2>&1; tail -1 gpurun_out/ncu_plain.log
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:em_sweep_tiled -s 5 -c 1 -f \
    -o gpurun_out/sweep_weighted python tools/dev/prof_w.py C3 > gpurun_out/ncu_weighted.log 2>&1; tail -1 gpurun_out/ncu_weighted.log
echo "== f-rows, robustness"
timeout 300 python tools/bench_frows.py C3 2>/dev/null | tail -1 > gpurun_out/frows.json; head -c 300 gpurun_out/frows.json; echo
timeout 400 python tools/bench_robust.py C3 2>/dev/null | tail -3 > gpurun_out/robust.jsonl; cut -c1-160 gpurun_out/robust.jsonl
if [ "$1" != "nosan" ]; then
  echo "== compute-sanitizer"
  : > gpurun_out/sanitizer.log
  for tool in racecheck memcheck synccheck; do
    echo "== --tool $tool" >> gpurun_out/sanitizer.log
    timeout 250 compute-sanitizer --tool $tool python tools/dev/sanit.py 2>&1 | grep -E "SUMMARY|^ok|Error|hazard" | head -8 >> gpurun_out/sanitizer.log
  done
  cat gpurun_out/sanitizer.log
fi
