"""The benchmark's JSON contract on a real GPU (short run): one line on stdout, the keys the driver reads, sane values."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '60', '--warmup', '5', '--no-cpu-baseline', '--no-fit'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['metric'].startswith('train cells/sec') and d['unit'] == 'cells/s' and d['n_gpus'] == 1
    assert d['steps'] == 60 and d['warmup'] == 5 and d['higher_is_better'] is True and d['scaling'] == 'weak'
    assert d['value'] > 1e5 and abs(d['value'] - 512 / (d['ms_per_step'] * 1e-3)) < 1e-6 * d['value']
    assert d['vs_baseline'] is None and d['data'] == 'synthetic' and 'workload' in d['config']
    e = d['e2e']
    assert 0 < e['value'] <= 1.05 * d['value'] and e['h2d_bytes_per_step'] > 2 * 512 * 512 * 4 and e['d2h_bytes_per_step'] == 32
    # the 60 timed steps are ONE launch of the persistent step kernel (+ the barrier-counter reset node is a memset, and the
    # one-thread control-block update kernel): 2 kernel launches
    assert d['gpu_launches'] >= 1 and abs(d['gpu_launches'] - d['launches_per_step'] * 60) < 1e-6
    r = d['roofline']
    assert r['bound'] == 'hbm' and 0 < r['frac'] < 1 and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert r['traffic'] is None or r['traffic'] > 0
    g = d['gemm_roofline']
    assert g['bound'] == 'tensor' and 0 < g['frac'] < 1
    names = [nm for nm, _ in d['step_profile']['phases']]
    assert 'adam' in names and 'wgrad x12' in names and abs(d['step_profile']['sum_us'] - d['ms_per_step'] * 1e3) < 0.25 * d['ms_per_step'] * 1e3
    assert set(d['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
    p = d['modal_predict']
    assert p['value'] > 1e6 and p['e2e']['value'] > 1e5 and 0 < p['roofline']['frac'] < 1
    assert all(k in d['final_losses'] for k in ('KL', 'Rec', 'CosSim', 'F', 'total', 'grad_norm'))
