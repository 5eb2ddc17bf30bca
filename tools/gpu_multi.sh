#!/bin/bash
# Multi-GPU visit: consensus mode over N GPUs against one GPU through the public API, then the bench line at N GPUs.
#   gpurun --gpus N --timeout 1200 -- bash tools/gpu_multi.sh <tag> <N> [bench steps]
tag=${1:-rXX}; n=${2:-2}; steps=${3:-5}
out=gpurun_out
mkdir -p $out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
  tools/consensus_nccl_check.py 96 60000 1000 > $out/${tag}_consensus_${n}gpu.txt 2>&1
echo "consensus check rc=$?"; tail -3 $out/${tag}_consensus_${n}gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $n --steps $steps --warmup 3 > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err
echo "bench rc=$?"; tail -c 3000 $out/${tag}_bench_${n}gpu.json; tail -5 $out/${tag}_bench_${n}gpu.err
