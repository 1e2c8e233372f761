#!/bin/bash
# One GPU call: smoke, GPU tests, bench (ours + reference arm), in-graph trace, ncu launch list and full captures.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [quick]
tag=${1:-s3}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1
tail -2 $out/smoke_$tag.log
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -40 $out/pytest_$tag.log
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"; tail -3 $out/bench_$tag.err; cat $out/bench_$tag.json
timeout 300 python tools/trace_step.py --out $out/trace_$tag.json > $out/trace_$tag.txt 2>&1; tail -40 $out/trace_$tag.txt
if [ "$2" != "quick" ]; then
  timeout 300 python bench.py --impl reference --steps 40 --warmup 3 > $out/bench_ref_$tag.json 2>&1; cat $out/bench_ref_$tag.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
    python bench.py --profile --steps 6 --warmup 8 > $out/ncu_launch_$tag.log 2>&1; tail -2 $out/ncu_launch_$tag.log
  # full captures: the batched wgrad launch (12th GEMM of a step), the first two encoder GEMMs (3xTF32; the second with split-K), Adam
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 35 -c 3 -f -o $out/gemm_$tag \
    python bench.py --profile --steps 3 --warmup 3 > $out/ncu_gemm_$tag.log 2>&1; tail -2 $out/ncu_gemm_$tag.log
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:persistent -s 12 -c 2 -f -o $out/predict_$tag \
    python bench.py --predict-only > $out/ncu_predict_$tag.log 2>&1; tail -2 $out/ncu_predict_$tag.log
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adam -s 3 -c 1 -f -o $out/adam_$tag \
    python bench.py --profile --steps 3 --warmup 3 > $out/ncu_adam_$tag.log 2>&1; tail -2 $out/ncu_adam_$tag.log
fi
ls -la $out | tail -20
