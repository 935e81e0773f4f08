#!/usr/bin/env bash
# gpurun -- 'bash tools/gpu_session_sanitizer.sh [tag]'   (one GPU, ~10 minutes)
# compute-sanitizer passes over the small GPU tests (memcheck / racecheck evidence).  Written when no
# GPU time was left in round 2: until it has run, profiles/r02f_emulated_cuda_source_suite.txt (the same tests executed
# from the CUDA source on the host with guard pages around every device block) is the memory-checking evidence.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r03}
OUT=gpurun_out; mkdir -p $OUT
SMALL="tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_sell_sym.py tests/test_zz_gpu_r02_symmetric_tangent.py"
SKIP='not cook and not golden'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 0 \
    python -m pytest $SMALL -q -m gpu -x -k "$SKIP" -p no:cacheprovider > $OUT/${TAG}_sanitizer_$tool.txt 2>&1
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|passed|failed|Error|========= (Invalid|Race|Hazard)" $OUT/${TAG}_sanitizer_$tool.txt | head -20
done
