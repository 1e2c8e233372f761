bash tools/gpu_ab.sh s7 "JB_PIPE=0" "JB_PIPE=1" "JB_PIPE=1 JB_PIPE_BLOCKS=148" "JB_PIPE=1 JB_PIPE_BLOCKS=296" "JB_PIPE=1 JB_PIPE_PER_THREAD=1" "JB_PIPE=1 JB_PIPE_PRIO=0 JB_PIPE_BLOCKS=148" "JB_PIPE=1 JB_PIPE_BLOCKS=148 JB_PIPE_PDL=0"
JB_PIPE=1 timeout 300 python tools/trace_step.py --out gpurun_out/trace_s7.json > gpurun_out/trace_s7.txt 2>&1; tail -42 gpurun_out/trace_s7.txt
JB_PIPE=1 JB_PIPE_BLOCKS=148 timeout 300 python tools/trace_step.py --out gpurun_out/trace_s7b.json > gpurun_out/trace_s7b.txt 2>&1; tail -42 gpurun_out/trace_s7b.txt
