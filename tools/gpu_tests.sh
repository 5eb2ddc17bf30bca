#!/bin/bash
# GPU visit, tests only: the -m gpu suite (prints of the tie tallies kept), then the randomised parity sweep.
#   gpurun --timeout 1500 -- bash tools/gpu_tests.sh <tag> [fuzz seeds]
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q -rP --durations=8 > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
grep -E "structural|configs\[|passed|failed|error" $out/${tag}_pytest.log | tail -40
if [ -n "$2" ]; then
  timeout 900 python tests/fuzz_parity.py $2 > $out/${tag}_fuzz.log 2>&1
  tail -5 $out/${tag}_fuzz.log
fi
