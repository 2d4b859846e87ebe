"""Randomised comparison of the host side with the UNMODIFIED reference where it is staged
(``baseline/_ref``, put there by ``__graft_entry__.build()``; the tests are skipped elsewhere): what the
goldens pin for a fixed set of cases -- index maps of every region constructor (values AND order),
``get_index`` / ``get_position``, painted material vectors, the DIA operator factories, assembled
``a_*`` operators, boundary application and probe writing on the host -- here for seeded random
geometries. Bit equality throughout (pyfds/fields.py:23-57, 158-226, 273-535, 557-611;
pyfds/regions.py:125-145)."""

import os
import sys
import types
import warnings

import numpy as np
import pytest

import pyfds_b200 as fds
from conftest import bits

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


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype.kind == b.dtype.kind and np.array_equal(bits(a), bits(b))


def outcome(call):
    """('ok', result) or ('error', exception type name, message): positions off the grid points make
    several constructors fail with an assertion (pyfds/fields.py:568-569), which must match too."""
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return ('ok', call())
    except (AssertionError, ValueError, IndexError, ZeroDivisionError, TypeError) as error:
        return ('error', type(error).__name__, str(error))


def grid_kwargs(rng):
    return dict(x_samples=int(rng.integers(9, 70)), x_delta=float(rng.choice([1e-3, 2.5e-4, 0.1, 3e-2])),
                y_samples=int(rng.integers(7, 50)), y_delta=float(rng.choice([1e-3, 5e-4, 0.2, 7e-3])),
                t_samples=int(rng.integers(5, 40)), t_delta=float(rng.choice([1e-7, 1e-3, 2e-8])))


def both(pyfds, klass, kwargs, material):
    ours = getattr(fds, klass)(material=getattr(fds, material[0])(*material[1]), **kwargs)
    theirs = getattr(pyfds, klass)(material=getattr(pyfds, material[0])(*material[1]), **kwargs)
    return ours, theirs


def random_point(rng, field, on_grid):
    """A position inside the grid: on a grid point, or anywhere (the constructors round)."""
    kx, ky = rng.integers(0, field.x.samples), rng.integers(0, field.y.samples)
    if on_grid:
        return float(kx * field.x.increment), float(ky * field.y.increment)
    return (float(rng.uniform(0, (field.x.samples - 1) * field.x.increment)),
            float(rng.uniform(0, (field.y.samples - 1) * field.y.increment)))


@pytest.mark.parametrize('seed', range(12))
def test_region_constructors_match_the_reference(pyfds, seed):
    rng = np.random.default_rng(1000 + seed)
    kwargs = grid_kwargs(rng)
    ours, theirs = both(pyfds, 'Acoustic2D', kwargs, ('AcousticMaterial', (1500, 1000)))
    X = (ours.x.samples - 1) * ours.x.increment
    Y = (ours.y.samples - 1) * ours.y.increment
    checked = 0
    for case in range(40):
        kind = ('point', 'line', 'rect', 'tri', 'ellipse')[case % 5]
        on_grid = bool(rng.integers(0, 5))      # one case in five off the grid points
        if kind == 'point':
            args = (random_point(rng, ours, on_grid),)
            calls = [lambda f=f: f.get_point_region(*args) for f in (ours, theirs)]
        elif kind == 'line':
            p, q = random_point(rng, ours, on_grid), random_point(rng, ours, on_grid)
            if p == q:
                continue
            calls = [lambda f=f: f.get_line_region(p + q) for f in (ours, theirs)]
        elif kind == 'rect':
            p = random_point(rng, ours, on_grid)
            size = (float(rng.uniform(0, X - p[0])), float(rng.uniform(0, Y - p[1])))
            calls = [lambda f=f: f.get_rect_region(p + size) for f in (ours, theirs)]
        elif kind == 'tri':
            corners = sum((random_point(rng, ours, on_grid) for _ in range(3)), ())
            calls = [lambda f=f: f.get_tri_region(corners) for f in (ours, theirs)]
        else:
            centre = random_point(rng, ours, on_grid)
            radii = (float(rng.uniform(0.5 * ours.x.increment, 0.4 * X)),
                     float(rng.uniform(0.5 * ours.y.increment, 0.4 * Y)))
            calls = [lambda f=f: f.get_ellipse_region(centre, radii) for f in (ours, theirs)]
        a, b = outcome(calls[0]), outcome(calls[1])
        assert a[0] == b[0], (seed, case, kind, a, b)
        if a[0] == 'error':
            assert a[1:] == b[1:], (seed, case, kind, a, b)
            continue
        checked += 1
        a, b = a[1], b[1]
        ia, ib = np.asarray(a.indices), np.asarray(b.indices)
        assert ia.shape == ib.shape and np.array_equal(ia, ib), (seed, case, kind)
    assert checked >= 12, checked


@pytest.mark.parametrize('seed', range(6))
def test_index_and_position_lookups_match_the_reference(pyfds, seed):
    rng = np.random.default_rng(2000 + seed)
    kwargs = grid_kwargs(rng)
    ours, theirs = both(pyfds, 'Acoustic2D', kwargs, ('AcousticMaterial', (1500, 1000)))
    for _ in range(200):
        p = random_point(rng, ours, bool(rng.integers(0, 4)))
        assert outcome(lambda: int(ours.get_index(p))) == outcome(lambda: int(theirs.get_index(p))), p
        k = int(rng.integers(0, ours.num_points))
        assert same(ours.get_position(k), theirs.get_position(k)), k
    line = dict(x_samples=kwargs['x_samples'], x_delta=kwargs['x_delta'],
                t_samples=kwargs['t_samples'], t_delta=kwargs['t_delta'])
    ours1, theirs1 = both(pyfds, 'Acoustic1D', line, ('AcousticMaterial', (700, 0.01)))
    for _ in range(100):
        x = float(rng.integers(0, ours1.x.samples) * ours1.x.increment) if rng.integers(0, 4) \
            else float(rng.uniform(0, (ours1.x.samples - 1) * ours1.x.increment))
        assert outcome(lambda: int(ours1.get_index(x))) == outcome(lambda: int(theirs1.get_index(x)))
        k = int(rng.integers(0, ours1.num_points))
        assert same(ours1.get_position(k), theirs1.get_position(k))
    assert same(ours.x.vector, theirs.x.vector) and same(ours.y.vector, theirs.y.vector)


def paint(field, package, rng_seed, materials):
    """The same random material regions on either package's field."""
    rng = np.random.default_rng(rng_seed)
    X = (field.x.samples - 1) * field.x.increment
    Y = (field.y.samples - 1) * field.y.increment
    for m in materials:
        kx, ky = int(rng.integers(0, field.x.samples - 1)), int(rng.integers(0, field.y.samples - 1))
        wx = int(rng.integers(1, field.x.samples - kx))
        wy = int(rng.integers(1, field.y.samples - ky))
        rect = (kx * field.x.increment, ky * field.y.increment, wx * field.x.increment,
                wy * field.y.increment)
        field.add_material_region(field.get_rect_region(rect), getattr(package, m[0])(*m[1]))
    # an ellipse: the constructor looks up centre +- the y radius (pyfds/fields.py:524-525)
    ry = int(rng.integers(1, max(2, field.y.samples // 3)))
    cy = int(rng.integers(ry, field.y.samples - ry))
    cx = int(rng.integers(0, field.x.samples))
    centre = (cx * field.x.increment, cy * field.y.increment)
    field.add_material_region(
        field.get_ellipse_region(centre, (0.23 * X, ry * field.y.increment)),
        getattr(package, materials[0][0])(*materials[0][1]))


ACOUSTIC = [('AcousticMaterial', (1200, 900, 1e-3)), ('AcousticMaterial', (1350, 950, 0, 2e-3)),
            ('AcousticMaterial', (1480, 998, 1e-3, 3e-3, 0.6, 4180, 4150))]
THERMAL = [('ThermalMaterial', (450, 7800, (50, 30))), ('ThermalMaterial', (900, 2700, 200))]


@pytest.mark.parametrize('klass,base,materials,parameters,operators', [
    ('Acoustic2D', ('AcousticMaterial', (1500, 1000, 5e-4)), ACOUSTIC,
     ('sound_velocity', 'density', 'absorption_coef'),
     ('a_p_vx', 'a_p_vy', 'a_vx_p', 'a_vy_p', 'a_vx_vx', 'a_vy_vy')),
    ('Acoustic3DAxi', ('AcousticMaterial', (1500, 1000, 5e-4)), ACOUSTIC,
     ('sound_velocity', 'density', 'absorption_coef'),
     ('a_p_vx', 'a_p_vy', 'a_vx_p', 'a_vy_p', 'a_vx_vx', 'a_vy_vy')),
    ('Thermal2D', ('ThermalMaterial', (900, 2700, 200)), THERMAL,
     ('heat_capacity', 'density'), None),
])
@pytest.mark.parametrize('seed', range(3))
def test_painted_materials_and_assembled_operators_match_the_reference(pyfds, klass, base, materials,
                                                                       parameters, operators, seed):
    rng = np.random.default_rng(3000 + seed)
    kwargs = grid_kwargs(rng)
    ours, theirs = both(pyfds, klass, kwargs, base)
    paint(ours, fds, 3100 + seed, materials)
    paint(theirs, pyfds, 3100 + seed, materials)
    for name in parameters:
        assert same(ours.material_vector(name), theirs.material_vector(name)), name
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        theirs.assemble_matrices()
    ours.assemble_matrices()        # host only: bakes the snapshot the lazy a_* operators come from
    if operators is None:
        operators = [name for name in vars(theirs) if name.startswith('a_')]
        assert operators
    for name in operators:
        a, b = getattr(ours, name).todia(), getattr(theirs, name).todia()
        assert np.array_equal(a.offsets, b.offsets), name
        assert same(a.data, b.data), name


@pytest.mark.parametrize('seed', range(4))
def test_operator_factories_match_the_reference(pyfds, seed):
    rng = np.random.default_rng(4000 + seed)
    kwargs = grid_kwargs(rng)
    ours, theirs = both(pyfds, 'Acoustic2D', kwargs, ('AcousticMaterial', (1500, 1000)))
    factors = rng.standard_normal(ours.num_points)
    for name in ('d_x', 'd_y'):
        for variant in ('forward', 'backward', 'central'):
            for f in (None, factors):
                a = getattr(ours, name)(factors=f, variant=variant).todia()
                b = getattr(theirs, name)(factors=f, variant=variant).todia()
                assert np.array_equal(a.offsets, b.offsets) and same(a.data, b.data), (name, variant)
    for name in ('d_x2', 'd_y2'):
        for f in (None, factors):
            a, b = getattr(ours, name)(factors=f).todia(), getattr(theirs, name)(factors=f).todia()
            assert np.array_equal(a.offsets, b.offsets) and same(a.data, b.data), name


@pytest.mark.parametrize('seed', range(4))
def test_host_boundaries_and_probes_match_the_reference(pyfds, seed):
    """FieldComponent.apply_bounds / write_outputs on the host (the statement the device tables are
    baked from): overlapping regions, additive and overwriting, scalar / signal / per-point signal values."""
    rng = np.random.default_rng(5000 + seed)
    kwargs = grid_kwargs(rng)
    ours, theirs = both(pyfds, 'Acoustic2D', kwargs, ('AcousticMaterial', (1500, 1000)))
    steps = kwargs['t_samples']
    state = rng.standard_normal(ours.num_points)
    for field in (ours, theirs):
        r = np.random.default_rng(5100 + seed)
        field.pressure.values = state.copy()
        for k in range(6):
            p, q = random_point(r, field, True), random_point(r, field, True)
            if p == q:
                q = (p[0], p[1] + field.y.increment) if p[1] == 0 else (p[0], 0.0)
            region = field.get_line_region(p + q) if k % 2 else field.get_rect_region(
                (min(p[0], q[0]), min(p[1], q[1]), abs(p[0] - q[0]), abs(p[1] - q[1])))
            n = len(region.indices)
            # a scalar, one signal for the region, a list with one signal per point
            value = (float(r.standard_normal()), r.standard_normal(steps),
                     [r.standard_normal(steps) for _ in range(n)])[k % 3]
            field.pressure.add_boundary(region, value=value, additive=bool(k & 2))
        field.pressure.add_output(field.get_line_region(
            (0, 0, (field.x.samples - 1) * field.x.increment, 0)))
        field.pressure.add_output(field.get_point_region(random_point(r, field, True)))
    for step in range(steps):
        for field in (ours, theirs):
            field.pressure.apply_bounds(step)
            field.pressure.write_outputs()
            field.pressure.values = field.pressure.values * 0.5 + 0.25
        assert same(ours.pressure.values, theirs.pressure.values), step
    for a, b in zip(ours.pressure.outputs, theirs.pressure.outputs):
        assert same(np.asarray(a.signals), np.asarray(b.signals))
