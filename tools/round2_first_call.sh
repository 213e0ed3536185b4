#!/bin/bash
# First GPU call of round 2: parity + timing of everything that round 1 prepared but could not run
# (its GPU budget was spent).  Build the variant library first, on the CPU box:
#     tests/_build_variants.sh scarce "-DOAR_GREEDY_SCARCE=1"
# then:  gpurun --timeout 600 -- tools/round2_first_call.sh        (results: gpurun_out/ab.log, gpurun_out/round2_first.log)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== A/B on C3: default, single-barrier variants, scarce-first greedy"
  tests/_ab.sh "- 2b 5" "- 1b 4" "- 1c 4" "scarce 2b 5"
  echo "== store creation with and without the overlapped upload (create wall, 3 repeats each)"
  python tests/_build_prof.py C3 3 | tail -2
  OAR_UPLOAD_OVERLAP=1 python tests/_build_prof.py C3 3 | tail -2
  echo "== parity suite incl. the experimental kernel (OAR_SWEEP=1c), then with the overlapped upload"
  OAR_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests -m gpu -x -q | tail -3
  OAR_UPLOAD_OVERLAP=1 timeout 300 python -m pytest tests -m gpu -x -q | tail -3
  echo "== racecheck of the experimental kernel"
  OAR_SWEEP=1c timeout 60 compute-sanitizer --tool racecheck python tests/_sanit.py 2>&1 | grep -E "SUMMARY|^ok|hazard" | head -4
} 2>&1 | tee gpurun_out/round2_first.log
