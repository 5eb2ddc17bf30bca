#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list and one full capture of the DP kernels.
#   gpurun --timeout 1500 -- bash tools/gpu_round.sh <tag> [skip-tests]
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
if [ "$2" != "skip-tests" ]; then
  python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?" >> $out/${tag}_pytest.log
  tail -5 $out/${tag}_pytest.log
fi
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; cat $out/${tag}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
cat $out/${tag}_bench_ref.json
# launch list of the same command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
# full capture of the DP kernels on a smaller batch (ncu replays each kernel ~40 times)
ncu --set full --clock-control none --import-source on -k regex:'sweep|snp|path' -c 9 -f -o $out/${tag}_full \
  python bench.py --reads 128 --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
echo "ncu rc=$?"
