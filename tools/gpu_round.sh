#!/bin/bash
# One GPU call: smoke, GPU tests, bench (ours + reference arm). Usage (under gpurun): bash tools/gpu_round.sh <tag> [quick]
tag=${1:-s0}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1
tail -2 $out/smoke_$tag.log
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -30 $out/pytest_$tag.log
if [ "$2" == "quick" ]; then
  timeout 600 python bench.py --no-cpu-baseline --no-fit > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"; tail -3 $out/bench_$tag.err; cat $out/bench_$tag.json
else
  timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"; tail -3 $out/bench_$tag.err; cat $out/bench_$tag.json
  timeout 600 python bench.py --impl reference --steps 40 --warmup 5 > $out/bench_ref_$tag.json 2>&1; cat $out/bench_ref_$tag.json
fi
