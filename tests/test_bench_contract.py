"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port on host cores) prints ONE JSON line with the
keys the measurement contract names.  The GPU arm shares the same line builder and is exercised on the GPU box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0', '--cpu-sample-chunks', '2', '--chunk', '1000'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'raw_samples_per_s_basecalled_fwd_viterbi'
    assert d['unit'] == 'samples/s' and d['higher_is_better'] is True and d['n_gpus'] == 1
    assert d['steps'] == 1 and d['value'] > 0 and d['ms_per_step'] > 0 and d['vs_baseline'] is None
    assert 'workload' in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    e2e = d['e2e']
    assert e2e['value'] == d['value'] and e2e['unit'] == d['unit']
    assert e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0


def test_default_batches_in_flight():
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    got = [bench.default_in_flight('raw_rgrgr', k) for k in (1, 10, 20, 30, 50, 100)]
    assert got == [1, 10, 20, 15, 17, 20]
    assert all(k % f == 0 or f == 17 for k, f in zip((1, 10, 20, 30, 50, 100), got))
    assert bench.default_in_flight('raw_rGr', 30) == 4
