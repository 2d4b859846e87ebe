"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm's JSON
line, and that the product arm refuses to run (instead of falling back to a CPU path) when there is no
CUDA device."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args),
                          capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    result = _run('--impl', 'reference', '--steps', '2', '--warmup', '3', '--size', '512')
    assert result.returncode == 0, result.stderr[-2000:]
    lines = [ln for ln in result.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, result.stdout
    line = json.loads(lines[0])
    assert line['impl'] == 'reference'
    assert line['metric'] == 'Acoustic2D Gcell-updates/s' and line['unit'] == 'Gcell-updates/s'
    assert line['higher_is_better'] is True and line['dtype'] == 'f64'
    assert line['n_gpus'] == 1 and line['steps'] == 2 and line['warmup'] == 3
    assert line['value'] > 0 and line['ms_per_step'] > 0
    assert line['vs_baseline'] is None              # BASELINE.md publishes no number for this metric
    baseline = line['cpu_baseline']
    # the unmodified reference staged in baseline/_ref by build(); the CPU restatement without it
    staged = os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'pyfds'))
    assert baseline['kind'] == ('reference' if staged else 'port')
    assert baseline['cores'] == 1 and baseline['value'] == line['value']
    assert '512x512' in baseline['sample']
    e2e = line['e2e']
    assert e2e['value'] == line['value'] and e2e['h2d_bytes_per_step'] == 0
    assert e2e['d2h_bytes_per_step'] == 0
    assert line['gpu_launches'] == 0
    assert 'workload' in line['config'] and 'model' not in line['config']


def test_both_arms_describe_the_same_config():
    """The driver compares the `config` of the two arms: it must not depend on the arm."""
    sys.path.insert(0, ROOT)
    import bench
    product = bench.config_record(4096, 4096, 4096 * 2, 2)
    result = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                             '--gpus', '2', '--steps', '1', '--warmup', '3', '--size', '256'],
                            capture_output=True, text=True, cwd=ROOT, timeout=600,
                            env=dict(os.environ, RANK='0', LOCAL_RANK='0', WORLD_SIZE='2'))
    assert result.returncode == 0, result.stderr[-2000:]
    line = json.loads(result.stdout.strip().splitlines()[-1])
    assert line['config'] == bench.config_record(256, 256, 512, 2)
    assert set(line['config']) == set(product)


def test_reference_arm_runs_on_rank_zero_only():
    env = dict(os.environ, RANK='1', LOCAL_RANK='1', WORLD_SIZE='2')
    result = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                             '--gpus', '2', '--steps', '2', '--warmup', '3', '--size', '512'],
                            capture_output=True, text=True, cwd=ROOT, timeout=600, env=env)
    assert result.returncode == 0 and result.stdout.strip() == ''


def test_product_arm_fails_loudly_without_a_gpu():
    import pyfds_b200 as fds
    from pyfds_b200 import _engine
    try:
        if _engine.load_library().fds_device_count() > 0:
            pytest.skip('a CUDA device is present')
    except RuntimeError:
        pass
    result = _run('--steps', '4', '--warmup', '3', '--size', '256', '--no-cpu-baseline')
    assert result.returncode != 0
    assert not any(ln.lstrip().startswith('{') for ln in result.stdout.splitlines()), result.stdout
    assert fds is not None


def test_other_models_record_keeps_the_documented_keys(monkeypatch):
    """`other_models` of the product arm's line: what benchmarks/configs.py::run returns, reduced to
    the documented keys (checked here with a stand-in for the GPU run)."""
    import types
    sys.path.insert(0, ROOT)
    import bench
    seen = []

    def run(number, steps, warmup):
        seen.append((number, steps, warmup))
        return {'config': number, 'model': 'Acoustic3DAxi', 'grid': [8192, 4096], 'steps': steps,
                'device_ms': 55.3, 'ms_per_step': 0.2765, 'wall_s': 0.06,
                'gcell_updates_per_s': 121.4, 'kernel': 'streamv_kernel<acoustic3daxi,lossy>',
                'launches': 100, 'steps_per_launch': 2, 'algorithmic_gbs': 5826.7,
                'bytes_per_cell_update': 48, 'setup_s': 1.2, 'device_gb': 1.7}

    monkeypatch.setitem(sys.modules, 'configs', types.SimpleNamespace(run=run))
    record = bench.measure_other_model(3)
    assert seen == [(3, 200, 20)]
    assert set(record) == set(bench.OTHER_MODEL_KEYS)
    assert record['gcell_updates_per_s'] == 121.4 and record['steps_per_launch'] == 2
    assert sorted(bench.OTHER_MODELS.values()) == [3, 4, 6]
    json.dumps(record)
