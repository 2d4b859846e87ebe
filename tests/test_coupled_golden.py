"""Coupled fields against goldens of the REAL reference (``tests/golden/coupled_*.npz``, produced by
``oracle/gen_golden.py`` from the unmodified ``pyfds.SynchronizedFields`` / ``BoundaryCoupling`` /
``MaterialCoupling*`` / ``ThermoAcoustic1D``):

* on the CPU, the host classes of ``pyfds_b200/coupling.py`` and ``coupled_fields.py`` driving the CPU
  restatement of the member fields -- bitwise;
* on the GPU (``-m gpu``), ``group.simulate()``: bitwise where the device evaluates the reference's
  arithmetic operation for operation (linear couplings, the viscous-heating term), and within the
  north star's 1e-12 relative L2 where a transfer function goes through the device's ``exp``
  (one unit in the last place from NumPy's)."""

import os

import numpy as np
import pytest

import pyfds_b200 as fds
import scenarios
from conftest import bits
from coupled_emulation import run_group_on_cpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _golden(name):
    with np.load(os.path.join(GOLDEN, 'coupled_' + name + '.npz')) as data:
        return {k: data[k] for k in data.files if k != 'versions'}


def _assert_bitwise(got, expected, name):
    assert sorted(got) == sorted(expected)
    for key in expected:
        a, b = np.asarray(got[key]), np.asarray(expected[key])
        assert a.shape == b.shape, (name, key, a.shape, b.shape)
        assert np.array_equal(bits(a), bits(b)), (name, key)


@pytest.mark.parametrize('name', sorted(scenarios.COUPLED_SCENARIOS))
def test_host_classes_on_cpu_restatement_equal_reference_golden_bitwise(name):
    group, steps = scenarios.COUPLED_SCENARIOS[name](fds)
    _assert_bitwise(run_group_on_cpu(group, steps), _golden(name), name)


#: scenarios whose device path repeats the reference's operations one for one
BITWISE_ON_DEVICE = ('thermoacoustic1d', 'thermoacoustic1d_stepping', 'boundary_coupling_linear',
                     'boundary_coupling_stepping', 'material_coupling_powerlaw')


def _simulate_segmented(group, steps):
    first = steps // 3
    group.simulate(first)
    group.simulate(steps - first)
    return scenarios.collect_group(group)


@pytest.mark.gpu
@pytest.mark.parametrize('name', BITWISE_ON_DEVICE)
def test_device_session_equals_reference_golden_bitwise(library, name):
    group, steps = scenarios.COUPLED_SCENARIOS[name](fds)
    got = _simulate_segmented(group, steps)
    assert getattr(group, '_last_session', None) == 'device', 'the group did not run on the device'
    _assert_bitwise(got, _golden(name), name)


@pytest.mark.gpu
def test_two_dimensional_material_coupling_equals_reference_golden_bitwise(library):
    """A 2-D target with a different material in every cell: the transfer function stays on the host
    (NumPy, as in the reference), every re-assembly uploads per-cell coefficient arrays and the field
    steps on the per-cell instantiation of the one-step kernel -- bitwise."""
    name = 'material_coupling_2d'
    group, steps = scenarios.COUPLED_SCENARIOS[name](fds)
    got = _simulate_segmented(group, steps)
    assert getattr(group, '_last_session', None) == 'per step'
    _assert_bitwise(got, _golden(name), name)
    engine = group.fields[0].__dict__['_engine_state'].engine
    assert engine.last_launch_info()[2] == 'step2d_kernel<acoustic2d,lossy>'


@pytest.mark.gpu
def test_device_exponential_law_within_north_star_tolerance(library):
    """exp() on the device and in NumPy may differ in the last place: rel. L2 <= 1e-12
    (BASELINE.json north_star), not bitwise."""
    name = 'material_coupling_exponential'
    group, steps = scenarios.COUPLED_SCENARIOS[name](fds)
    got = _simulate_segmented(group, steps)
    assert getattr(group, '_last_session', None) == 'device'
    expected = _golden(name)
    assert sorted(got) == sorted(expected)
    for key in expected:
        a = np.asarray(got[key], dtype=np.float64).reshape(-1)
        b = np.asarray(expected[key], dtype=np.float64).reshape(-1)
        assert a.shape == b.shape
        scale = np.linalg.norm(b)
        assert np.linalg.norm(a - b) <= 1e-12 * scale, (key, np.linalg.norm(a - b) / max(scale, 1e-300))
