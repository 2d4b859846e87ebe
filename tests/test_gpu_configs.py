"""The BASELINE.json configurations at their FULL sizes against the CPU restatement, bit for bit
(parity variant = the config's geometry with a random initial state so that every cell
is exercised), plus every configuration on a reduced grid for more steps. The restatement steps the
same scipy DIA operators as the reference (oracle/restate.py, backend 'scipy'), which it was pinned to
by tests/test_oracle.py; it takes a few seconds per configuration on the GPU box's host.

Config 5 (32768 x 32768) is beyond any CPU oracle: its parity is multi-GPU == single-GPU
(tests/test_gpu_multi.py, bench.py's `parity` record) on top of config 2's kernel."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'benchmarks'))

import scenarios  # noqa: E402
from conftest import bits  # noqa: E402
from oracle import restate  # noqa: E402

pytestmark = pytest.mark.gpu


def _randomise(field, seed):
    rng = np.random.default_rng(seed)
    for name in field._device_components:
        scale = 20.0 if name == 'temperature' else 1e-3
        getattr(field, name).values = scale * rng.standard_normal(field.num_points)


def _compare(field, steps, backend):
    stepper = restate.stepper_for(field, backend=backend).run(steps)
    first = steps // 2
    field.simulate(first)                    # segmented: the second call resumes at field.step
    field.simulate(steps - first)
    got, expected = scenarios.collect(field), scenarios.collect_stepper(stepper)
    assert sorted(got) == sorted(expected)
    for key in expected:
        a, b = np.asarray(got[key]), np.asarray(expected[key])
        assert a.shape == b.shape and np.array_equal(bits(a), bits(b)), key
    engine = field.__dict__['_engine_state'].engine
    kernel = engine.last_launch_info()[2]
    engine.close()
    field.__dict__['_engine_state'].engine = None
    return kernel


# (config, full-size keyword arguments, steps, kernel expected)
FULL = [
    (2, dict(nx=4096, ny=4096), 9, 'stream2d_kernel<acoustic2d,lossless>'),
    (6, dict(nx=4096, ny=4096), 7, 'streamv_kernel<acoustic2d,lossy>'),
    (3, dict(nx=8192, ny=4096), 7, 'streamv_kernel<acoustic3daxi,lossy>'),
    (7, dict(nx=8192, ny=4096), 9, 'stream2d_kernel<acoustic3daxi,lossless>'),
    (4, dict(nx=8192, ny=8192), 6, 'stream2d_kernel<thermal2d>'),
]


@pytest.mark.parametrize('number,size,steps,kernel', FULL)
def test_config_at_full_size_equals_cpu_restatement_bitwise(library, number, size, steps, kernel):
    import configs
    field, _ = configs.CONFIGS[number](t_samples=steps + 1, **size)
    _randomise(field, 100 + number)
    assert _compare(field, steps, 'scipy') == kernel


def test_config1_full_run_equals_cpu_restatement_bitwise(library):
    """Acoustic1D, 10 000 cells, all 20 000 steps, the Gauss pulse and the probe of the config."""
    import configs
    field, _ = configs.CONFIGS[1]()
    kernel = _compare(field, 20000, 'restated')
    assert kernel.startswith('line1d_kernel')


@pytest.mark.parametrize('number', sorted(set(range(1, 9))))
def test_config_on_reduced_grid_equals_cpu_restatement_bitwise(library, number):
    import configs
    assert configs.check(number)
