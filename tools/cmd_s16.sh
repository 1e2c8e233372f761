export PYTHONUNBUFFERED=1
run() { echo "== $*"; env $* python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 400 --warmup 40 --no-predict 2>&1 | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(round(d['ms_per_step'] * 1e3, 1), 'us/step', round(d['value'] / 1e6, 2), 'M cells/s')"; }
run JB_DP_MODE=single
run JB_DP_MODE=overlap
run JB_DP_MODE=single NCCL_ALGO=NVLS
run JB_DP_MODE=single NCCL_NVLS_ENABLE=0
