#!/bin/bash
# Multi-GPU measurement batch (BASELINE configs[1..4]); each line is one torchrun job, results under gpurun_out/.
# usage: tools/run_multi.sh "<space separated job names>"   jobs: s8 s4 s2 b2 b4 b8 m8 c1 c2 c4 c8
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29500
EXTRA=${EXTRA:-}
for job in $1; do
  port=$((port+1))
  n=${job:1}
  case ${job:0:1} in
    s) cmd="$TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 5 --warmup 3 $EXTRA";;
    b) cmd="$TR --nproc-per-node $n --master-port $port bench.py --gpus $n --model 14B --steps 2 --warmup 3";;
    m) cmd="$TR --nproc-per-node $n --master-port $port bench.py --gpus $n --model 14B --attn int8 --ffn-bits 4 --steps 2 --warmup 3";;
    c) if [ "$n" = "1" ]; then cmd="python tools/bench_calibration.py"; else cmd="$TR --nproc-per-node $n --master-port $port tools/bench_calibration.py"; fi;;
  esac
  echo "== $job: $cmd"
  t0=$(date +%s)
  timeout 600 $cmd > gpurun_out/r2_multi_$job${TAG:-}.json 2> gpurun_out/r2_multi_$job${TAG:-}.err
  echo "   rc=$? $(( $(date +%s) - t0 )) s: $(grep -o "\"value\": [0-9.]*" gpurun_out/r2_multi_$job${TAG:-}.json | head -1)"
  tail -c 200 gpurun_out/r2_multi_$job${TAG:-}.err | tr '\n' ' '; echo
done
