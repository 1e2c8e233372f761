bash tools/gpu_ab.sh s18 "JB_L2_PERSIST=0" "JB_L2_PERSIST=1 JB_DEBUG_STAGES=1" "JB_L2_PERSIST=0" "JB_L2_PERSIST=1"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
