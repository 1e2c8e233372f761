export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_eval.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -8
run() { echo "== $*"; env $* timeout 300 python bench.py --predict-only 2>/dev/null | tail -1 | cut -c1-160; }
run JB_EVAL_PERSIST=1
run JB_EVAL_PERSIST=1
