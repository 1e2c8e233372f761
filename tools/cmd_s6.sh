bash tools/gpu_round.sh s6 quick
bash tools/gpu_ab.sh s6 "JB_PIPE=0 JB_WGRAD_BN=128" "JB_PIPE=0" "JB_PIPE=1 JB_PIPE_PDL=0" "JB_PIPE=1"
