export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 500 --warmup 50 --no-predict > gpurun_out/bench_n2_s15.json 2> gpurun_out/bench_n2_s15.err; tail -5 gpurun_out/bench_n2_s15.err; cut -c1-330 gpurun_out/bench_n2_s15.json
