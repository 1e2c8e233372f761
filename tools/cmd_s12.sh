export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 600 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline > gpurun_out/bench_s12.json 2> gpurun_out/bench_s12.err; tail -3 gpurun_out/bench_s12.err
python -c "
import json; d=json.load(open('gpurun_out/bench_s12.json')); print(d['ms_per_step']*1e3, d['value'], 'e2e', d['e2e']['value'], 'pred', d['modal_predict']['value'], d['modal_predict']['e2e']['value'], d['roofline'])"
