#!/bin/bash
# A/B of step-graph variants on one box: short bench runs (device-resident value only) under different env switches.
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
tag=${1:-ab}
timeout 300 ./tools/gemm_lab check > $out/lab_check_$tag.txt 2>&1; tail -8 $out/lab_check_$tag.txt
timeout 300 ./tools/gemm_lab time > $out/lab_time_$tag.txt 2>&1; grep "bn128 split0\|bn256" $out/lab_time_$tag.txt
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline --no-predict 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print(round(d['ms_per_step'] * 1e3, 1), 'us/step', d['launches_per_step'], 'launches', [round(x, 1) for x in d['gemm_stages_us']])
    elif 'rror' in ln: print(ln.strip())
"; }
run JB_FUSE=0
run JB_FUSE=1
run JB_FUSE=1 JB_FUSE_SPLITK=1
run JB_FUSE=1 JB_WGRAD_BN=256
