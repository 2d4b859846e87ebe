"""The CPU restatement (``oracle/restate.py``) against the UNMODIFIED reference on seeded RANDOM
scenarios, where the reference is staged (``baseline/_ref``; skipped elsewhere). The committed goldens
pin the restatement on 27 fixed scenarios; here every model steps random grids, random material
regions (lossless and lossy), overlapping boundaries of all three value kinds on every component,
probes and a random initial state -- fields and probe signals bit for bit, for all three backends of
the restatement's mat-vec. The restatement reads the reference's own field objects, so both sides
see the very same scenario."""

import os
import sys
import types
import warnings

import numpy as np
import pytest

import scenarios
from conftest import bits
from oracle import restate

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


MODELS = {
    # class, dimensions, material factory (package, rng, lossy), components
    'Acoustic1D': (1, 'acoustic', ('pressure', 'velocity')),
    'Acoustic2D': (2, 'acoustic', ('pressure', 'velocity_x', 'velocity_y')),
    'Acoustic3DAxi': (2, 'acoustic', ('pressure', 'velocity_x', 'velocity_y')),
    'Thermal1D': (1, 'thermal', ('temperature', 'heat_flux')),
    'Thermal2D': (2, 'thermal', ('temperature', 'heat_flux_x', 'heat_flux_y')),
    'Thermal3DAxi': (2, 'thermal', ('temperature', 'heat_flux_x', 'heat_flux_y')),
}


def material(package, kind, rng, lossy):
    if kind == 'acoustic':
        c, rho = float(rng.uniform(300, 1600)), float(rng.uniform(1, 1200))
        if not lossy:
            return package.AcousticMaterial(c, rho)
        if rng.integers(0, 2):
            return package.AcousticMaterial(c, rho, absorption_coef=float(rng.uniform(0, 50)))
        return package.AcousticMaterial(c, rho, shear_viscosity=float(rng.uniform(0, 1e-2)),
                                        bulk_viscosity=float(rng.uniform(0, 1e-2)))
    conductivity = float(rng.uniform(1, 300))
    if rng.integers(0, 2):
        conductivity = (conductivity, float(rng.uniform(1, 300)))
    return package.ThermalMaterial(float(rng.uniform(300, 1000)), float(rng.uniform(500, 8000)),
                                   conductivity)


def random_scenario(package, klass, seed, lossy):
    dims, kind, components = MODELS[klass]
    rng = np.random.default_rng(seed)
    steps = int(rng.integers(6, 16))
    nx, ny = int(rng.integers(12, 48)), int(rng.integers(9, 30))
    dx, dy = float(rng.choice([1e-3, 5e-4, 2e-3])), float(rng.choice([1e-3, 7e-4]))
    dt = 1e-7 if kind == 'acoustic' else 1e-4
    if dims == 1:
        field = getattr(package, klass)(x_samples=nx, x_delta=dx, t_samples=steps, t_delta=dt,
                                        material=material(package, kind, rng, lossy))
    else:
        field = getattr(package, klass)(x_samples=nx, x_delta=dx, y_samples=ny, y_delta=dy,
                                        t_samples=steps, t_delta=dt,
                                        material=material(package, kind, rng, lossy))

    def region():
        """A random line (1-D, or 2-D in any direction), rectangle or point on grid points."""
        if dims == 1:
            a, b = sorted(int(k) for k in rng.choice(nx, 2, replace=False))
            if rng.integers(0, 3) == 0:
                return field.get_point_region(a * dx)
            return field.get_line_region((a * dx, b * dx))
        ax, bx = (int(k) for k in rng.integers(0, nx, 2))
        ay, by = (int(k) for k in rng.integers(0, ny, 2))
        what = int(rng.integers(0, 4))
        if what == 0 or (ax == bx and ay == by):
            return field.get_point_region((ax * dx, ay * dy))
        if what == 1:
            return field.get_line_region((ax * dx, ay * dy, bx * dx, by * dy))
        x0, y0 = min(ax, bx), min(ay, by)
        return field.get_rect_region((x0 * dx, y0 * dy, abs(ax - bx) * dx, abs(ay - by) * dy))

    for _ in range(int(rng.integers(0, 4))):
        field.add_material_region(region(), material(package, kind, rng, lossy))
    scale = 1e-3 if kind == 'acoustic' else 20.0
    for name in components:
        component = getattr(field, name)
        component.values = scale * rng.standard_normal(field.num_points)
        for _ in range(int(rng.integers(0, 4))):
            where = region()
            n = len(where.indices)
            choice = int(rng.integers(0, 3))
            value = (float(scale * rng.standard_normal()), scale * rng.standard_normal(steps),
                     [scale * rng.standard_normal(steps) for _ in range(n)])[choice]
            component.add_boundary(where, value=value, additive=bool(rng.integers(0, 2)))
        for _ in range(int(rng.integers(0, 3))):
            component.add_output(region())
    return field, steps


def assert_same(got, expected, context):
    assert sorted(got) == sorted(expected), context
    for key in expected:
        a, b = np.asarray(got[key], dtype=np.float64), np.asarray(expected[key], dtype=np.float64)
        assert a.shape == b.shape, (context, key, a.shape, b.shape)
        assert np.array_equal(bits(a), bits(b)), (context, key)


CASES = [(klass, lossy) for klass in sorted(MODELS)
         for lossy in ((False, True) if MODELS[klass][1] == 'acoustic' else (False,))]


@pytest.mark.parametrize('klass,lossy', CASES)
@pytest.mark.parametrize('seed', range(4))
def test_restatement_equals_the_reference_on_random_scenarios(pyfds, klass, lossy, seed):
    field, steps = random_scenario(pyfds, klass, 7000 + 10 * seed + lossy, lossy)
    steppers = [restate.stepper_for(field, backend=backend)
                for backend in ('restated', 'scipy', 'c')]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        first = steps // 2
        field.simulate(first)                 # segmented, as the device path is driven
        field.simulate(steps - first)
    expected = scenarios.collect(field)
    for stepper in steppers:
        stepper.run(steps)
        assert_same(scenarios.collect_stepper(stepper), expected,
                    '{} lossy={} seed={} backend={}'.format(klass, lossy, seed, stepper.backend))


@pytest.mark.parametrize('seed', range(3))
def test_restatement_equals_the_reference_with_flow(pyfds, seed):
    """AcousticFlow2D (pyfds/acoustic_flow.py:44-57): the row shift after every step, periods of
    either sign and rows that never move."""
    rng = np.random.default_rng(7100 + seed)
    nx, ny, steps = int(rng.integers(12, 40)), int(rng.integers(8, 24)), 14
    flow = rng.uniform(-4000, 4000, ny)
    flow[rng.integers(0, ny)] = 1e-9          # a row that practically never moves
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        field = pyfds.AcousticFlow2D(flow, x_samples=nx, x_delta=1e-3, y_samples=ny, y_delta=1e-3,
                                     t_samples=steps, t_delta=1e-7,
                                     material=pyfds.AcousticMaterial(1500, 1000, 1e-3))
    for name in ('pressure', 'velocity_x', 'velocity_y'):
        getattr(field, name).values = 1e-3 * rng.standard_normal(field.num_points)
    field.pressure.add_output(field.get_line_region((0, 0, (nx - 1) * 1e-3, 0)))
    stepper = restate.stepper_for(field)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        field.simulate(steps)
    stepper.run(steps)
    assert_same(scenarios.collect_stepper(stepper), scenarios.collect(field), 'flow seed {}'.format(seed))


# ---- the reference's documented usage (doc/ex_acoustics.rst, doc/usage.rst), scaled down --------------

def documented_waveguide(package, steps=90):
    """doc/ex_acoustics.rst:52-93 in the style the documentation uses -- ``boundaries.append(
    Boundary(...))``, ``outputs.append(Output(...))``, ``max(fld.y.vector)``, a Gauss pulse built
    from ``fld.t.vector`` -- on a grid a tenth of the documented size."""
    fld = package.Acoustic3DAxi(t_delta=1e-6, t_samples=steps, x_delta=1e-3, x_samples=31,
                                y_delta=1e-3, y_samples=91,
                                material=package.AcousticMaterial(sound_velocity=700, density=1000,
                                                                  shear_viscosity=1e-2))
    t = fld.t.vector - 1.2e-5
    ex_signal = np.exp(-(t / 4e-6) ** 2) * np.cos(2 * np.pi * 5e4 * t)
    fld.velocity_x.add_boundary(fld.get_line_region((0, 0, 0, max(fld.y.vector))))
    fld.velocity_x.boundaries.append(package.Boundary(
        fld.get_line_region((20e-3, 0, 20e-3, max(fld.y.vector)))))
    fld.pressure.boundaries.append(package.Boundary(
        fld.get_line_region((0, 20e-3, 19e-3, 20e-3)), value=ex_signal, additive=True))
    fld.pressure.outputs.append(package.Output(fld.get_line_region((0, 70e-3, 20e-3, 70e-3))))
    return fld, steps


def documented_line(package, steps=400):
    """doc/ex_acoustics.rst:9-45: the 1-D pulse between a rigid and a pressure-release end."""
    fld = package.Acoustic1D(t_delta=1e-7, t_samples=steps, x_delta=1e-3, x_samples=301,
                             material=package.AcousticMaterial(sound_velocity=700, density=0.01,
                                                               shear_viscosity=1e-3))
    t = fld.t.vector - 0.1e-4
    ex_signal = np.exp(-(t / 3e-6) ** 2) * np.cos(2 * np.pi * 1e5 * t)
    fld.velocity.boundaries.append(package.Boundary(fld.get_point_region(0)))
    fld.pressure.boundaries.append(package.Boundary(fld.get_point_region(max(fld.x.vector))))
    fld.pressure.boundaries.append(package.Boundary(fld.get_point_region(0.1), value=ex_signal,
                                                    additive=True))
    fld.pressure.outputs.append(package.Output(fld.get_point_region(0.2)))
    return fld, steps


@pytest.mark.parametrize('builder', [documented_waveguide, documented_line])
def test_documented_usage_builds_the_same_scenario_under_this_package(pyfds, builder):
    """The script of the documentation, written once and run with either package: the reference
    steps its own field; this package's field (same calls, same attribute names) is stepped by the
    restatement. Equal bits mean the two packages built the same scenario out of the script."""
    import pyfds_b200 as fds
    theirs, steps = builder(pyfds)
    ours, _ = builder(fds)
    assert bool(ours.is_stable()) == bool(theirs.is_stable())
    stepper = restate.stepper_for(ours).run(steps)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        theirs.simulate()                       # no argument: all t_samples steps (fields.py:74-75)
    assert_same(scenarios.collect_stepper(stepper), scenarios.collect(theirs), builder.__name__)
    mean = np.mean(stepper.signals('pressure')[0], axis=0)
    assert np.array_equal(bits(mean), bits(theirs.pressure.outputs[0].mean_signal))
