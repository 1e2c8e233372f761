#!/bin/bash
# Scaling runs on one 8-GPU box (as the driver launches them): N = 8, 4, 2. Usage (under gpurun --gpus 8): bash tools/gpu_scale.sh <tag>
tag=${1:-scale}; out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 500 --warmup 50 \
    > $out/bench_n${n}_$tag.json 2> $out/bench_n${n}_$tag.err
  python -c "
import json, sys
d = json.load(open('$out/bench_n${n}_$tag.json')); print('N =', d['n_gpus'], round(d['ms_per_step'] * 1e3, 1), 'us/step', round(d['value'] / 1e6, 2), 'M cells/s', 'e2e', round(d['e2e']['value'] / 1e6, 2), 'predict', round(d['modal_predict']['value'] / 1e6, 1))" || tail -5 $out/bench_n${n}_$tag.err
done
