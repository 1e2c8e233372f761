#!/bin/bash
# A/B of step-graph variants on one box: short bench runs (device-resident value only) under different env switches.
# Usage (under gpurun): bash tools/gpu_ab.sh <tag> "ENV1=a ENV2=b" "ENV1=c" ...
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
tag=${1:-ab}; shift
run() { echo "== $*"; env $* timeout 300 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline --no-predict 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print(round(d['ms_per_step'] * 1e3, 1), 'us/step', d['launches_per_step'], 'launches', [round(x, 1) for x in d['gemm_stages_us']])
    elif 'rror' in ln: print(ln.strip())
"; }
for v in "$@"; do run $v; done
