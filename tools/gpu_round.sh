#!/bin/bash
# One GPU visit: parity tests, bench lines of both arms, ncu launch list and one full capture of the DP kernels.
#   gpurun --timeout 1800 -- bash tools/gpu_round.sh <tag> [skip-tests]
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
if [ "$2" != "skip-tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q -rP > $out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?" >> $out/${tag}_pytest.log
  grep -E "configs\[[03]\]|passed|failed|error" $out/${tag}_pytest.log | tail -6
fi
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; python tools/bench_show.py $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
python tools/bench_show.py $out/${tag}_bench_ref.json
# launch list of the same command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-api --no-overlap > $out/${tag}_ncu_launches.log 2>&1
# full capture of the DP kernels at the bench size: sweep5 plain, score, path2, sweep5 wobble, no_snp, snp3
ncu --set full --clock-control none --import-source on -k regex:'sweep|snp3|path2|score' -c 10 -f -o $out/${tag}_full \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-api --no-overlap --no-consensus > $out/${tag}_ncu_full.log 2>&1
echo "ncu rc=$?"
