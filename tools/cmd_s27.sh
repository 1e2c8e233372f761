export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_s27.log 2>&1; echo "memcheck exit $?"; grep -c "Invalid\|Error" gpurun_out/memcheck_s27.log; tail -4 gpurun_out/memcheck_s27.log
timeout 300 python bench.py --predict-only 2>/dev/null | tail -1 | cut -c1-200
