bash tools/gpu_ab.sh s10 "JB_SLAB_CW=16" "JB_SLAB_CW=8" "JB_SLAB_CW=16 JB_PDL=0" "JB_SLAB_CW=8 JB_SIDE=0"
JB_SLAB_CW=8 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
JB_SLAB_CW=8 timeout 300 python tools/trace_step.py --out gpurun_out/trace_s10.json > gpurun_out/trace_s10.txt 2>&1; tail -34 gpurun_out/trace_s10.txt
