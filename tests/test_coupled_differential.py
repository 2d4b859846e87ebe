"""Coupled fields against the UNMODIFIED reference on seeded RANDOM groups (where the reference is
staged in ``baseline/_ref``; skipped elsewhere): the classes of ``pyfds_b200/coupling.py`` /
``coupled_fields.py`` drive the CPU restatement of the member fields (``tests/coupled_emulation.py``),
the reference runs its own ``SynchronizedFields.simulate`` on the same group -- fields and probe
signals bit for bit. Covers what the seven goldens fix to single values: every additive / accumulate /
stepping combination of ``BoundaryCoupling`` (pyfds/coupling.py:90-140), the general
``MaterialCoupling`` with a Python transfer function, the exponential and power laws with and without
a change threshold (:143-300), ``ThermoAcoustic1D`` at random steppings (pyfds/coupled_fields.py:10-65)."""

import os
import sys
import types
import warnings

import numpy as np
import pytest

import pyfds_b200 as fds
import scenarios
from conftest import bits
from coupled_emulation import run_group_on_cpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, 'baseline', '_ref')


@pytest.fixture(scope='module')
def pyfds():
    if not os.path.isdir(os.path.join(STAGED, 'pyfds')):
        pytest.skip('reference not staged (baseline/_ref)')
    for mod in ('matplotlib', 'matplotlib.patches', 'matplotlib.pyplot', 'matplotlib.animation'):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    if STAGED not in sys.path:
        sys.path.insert(0, STAGED)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import pyfds as reference
    return reference


def two_fields(package, rng_seed, steps, nx):
    """An Acoustic1D and a Thermal1D of the same length with random materials, state, boundaries and
    probes -- the same for either package (everything is drawn from a generator seeded here)."""
    rng = np.random.default_rng(rng_seed)
    sound = package.Acoustic1D(t_delta=1e-7, t_samples=steps, x_delta=1e-3, x_samples=nx,
                               material=package.AcousticMaterial(
                                   float(rng.uniform(400, 900)), float(rng.uniform(0.01, 2)),
                                   shear_viscosity=float(rng.uniform(0, 2e-3))))
    heat = package.Thermal1D(t_delta=1e-3, t_samples=steps, x_delta=1e-3, x_samples=nx,
                             material=package.ThermalMaterial(900, 2700, float(rng.uniform(50, 300))))
    a, b = sorted(int(k) for k in rng.choice(nx, 2, replace=False))
    heat.add_material_region(heat.get_line_region((a * 1e-3, b * 1e-3)),
                             package.ThermalMaterial(450, 7800, 50))
    sound.pressure.values = 1e-2 * rng.standard_normal(nx)
    sound.velocity.values = 1e-2 * rng.standard_normal(nx)
    heat.temperature.values = 20 + 3 * np.sin(np.arange(nx) / 17.0) + 0.1 * rng.standard_normal(nx)
    sound.velocity.add_boundary(sound.get_point_region(0))
    k = int(rng.integers(1, nx - 1))
    sound.pressure.add_boundary(sound.get_point_region(k * 1e-3), value=rng.standard_normal(steps),
                                additive=True)
    heat.temperature.add_boundary(heat.get_point_region(0), value=float(rng.uniform(10, 50)))
    sound.pressure.add_output(sound.get_point_region(int(rng.integers(0, nx)) * 1e-3))
    sound.velocity.add_output(sound.get_point_region(int(rng.integers(0, nx)) * 1e-3))
    heat.temperature.add_output(heat.get_point_region(int(rng.integers(0, nx)) * 1e-3))
    heat.heat_flux.add_output(heat.get_line_region((2e-3, 6e-3)))
    return sound, heat


def random_group(package, seed):
    rng = np.random.default_rng(seed)
    steps, nx = int(rng.integers(20, 60)), int(rng.integers(40, 140))
    kind = ('boundary', 'boundary', 'exponential', 'power', 'general', 'thermoacoustic')[seed % 6]
    if kind == 'thermoacoustic':
        group = package.ThermoAcoustic1D(
            x_samples=nx, x_delta=1e-3, t_samples=steps, t_delta=1e-7,
            thermal_material=package.ThermalMaterial(900, 2700, 200),
            acoustic_material=package.AcousticMaterial(700, 0.01, shear_viscosity=1e-3),
            stepping=int(rng.integers(1, 6)))
        sound, heat = group.fields
        sound.pressure.values = rng.standard_normal(nx)
        sound.velocity.values = rng.standard_normal(nx)
        sound.velocity.add_boundary(sound.get_point_region(0))
        heat.temperature.add_output(heat.get_line_region((3e-3, 9e-3)))
        sound.pressure.add_output(sound.get_point_region(5e-3))
        return group, steps
    sound, heat = two_fields(package, seed + 50000, steps, nx)
    components = {'p': sound.pressure, 'v': sound.velocity, 't': heat.temperature,
                  'q': heat.heat_flux}
    interactions = []
    if kind == 'boundary':
        for _ in range(int(rng.integers(1, 4))):
            source, target = rng.choice(['p', 'v', 't', 'q'], 2, replace=False)
            scale = float(rng.choice([1e-4, -2.5, 3e-6, 0.5]))
            interactions.append(package.BoundaryCoupling(
                components[source], components[target], scenarios.linear_transfer(package, scale),
                additive=bool(rng.integers(0, 2)), accumulate=bool(rng.integers(0, 2)),
                stepping=int(rng.integers(1, 6))))
    else:
        threshold = None if rng.integers(0, 2) else float(rng.uniform(0.001, 0.05))
        stepping = int(rng.integers(1, 7))
        if kind == 'exponential':
            interactions.append(package.MaterialCouplingExponential(
                heat.temperature, sound, 'sound_velocity', a=float(rng.uniform(0.5, 1.2)),
                b=float(rng.uniform(-0.02, 0.02)), rel_change_threshold=threshold,
                stepping=stepping))
        elif kind == 'power':
            interactions.append(package.MaterialCouplingPowerLaw(
                sound.pressure, heat, 'density', power=int(rng.integers(1, 4)),
                factor=float(rng.uniform(1, 60)), rel_change_threshold=threshold,
                stepping=stepping))
        else:
            slope = float(rng.uniform(1e-3, 1e-2))
            interactions.append(package.MaterialCoupling(
                heat.temperature, sound, 'density', lambda values: 1.0 + slope * np.tanh(values / 30),
                rel_change_threshold=threshold, stepping=stepping))
    return package.SynchronizedFields([sound, heat], interactions), steps


@pytest.mark.parametrize('seed', range(18))
def test_random_group_equals_the_reference_bitwise(pyfds, seed):
    ours, steps = random_group(fds, seed)
    theirs, _ = random_group(pyfds, seed)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        theirs.simulate(steps)
    expected = scenarios.collect_group(theirs)
    got = run_group_on_cpu(ours, steps)
    assert sorted(got) == sorted(expected)
    for key in expected:
        a, b = np.asarray(got[key], dtype=np.float64), np.asarray(expected[key], dtype=np.float64)
        assert a.shape == b.shape, (seed, key, a.shape, b.shape)
        assert np.array_equal(bits(a), bits(b)), (seed, key)
