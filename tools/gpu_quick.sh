#!/bin/bash
# Quick GPU iteration: smoke, parity tests (fail fast), phase profile. Usage (under gpurun): bash tools/gpu_quick.sh <tag> [pytest -k expr]
tag=${1:-q0}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke exit $?"; tail -3 $out/smoke_$tag.log
if [ -n "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x -k "$2" > $out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -25 $out/pytest_$tag.log
else
  timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -25 $out/pytest_$tag.log
fi
timeout 300 python tools/profile_step.py > $out/prof_$tag.txt 2>&1; echo "prof exit $?"; cat $out/prof_$tag.txt | head -50
