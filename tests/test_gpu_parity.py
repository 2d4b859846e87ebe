"""Parity of the CUDA path against the reference (-m gpu): every scenario of tests/scenarios.py is run
through the drop-in's public API (`simulate()` -> ctypes -> C ABI -> sm_100a kernels) and compared

* bit for bit (fp64 viewed as int64, so signed zeros count) with the golden fields and probe signals
  that the REAL reference produced in the build container (tests/golden/, oracle/gen_golden.py);
* bit for bit with the CPU restatement (oracle/restate.py) on larger grids and on cases that have no
  golden file.

The north-star tolerance is a relative L2 error <= 1e-12; bit equality is stricter and is what we
assert. Nothing here reads /root/reference.
"""

import os

import numpy as np
import pytest

import pyfds_b200 as fds
import scenarios
from conftest import bits
from oracle import restate

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def assert_same(got, expected, context=''):
    assert sorted(got) == sorted(expected), context
    for key in expected:
        g, e = np.asarray(got[key]), np.asarray(expected[key])
        assert g.shape == e.shape, (context, key, g.shape, e.shape)
        if not np.array_equal(bits(g), bits(e)):
            denom = np.linalg.norm(e.ravel()) or 1.0
            rel = np.linalg.norm((g - e).ravel()) / denom
            bad = int(np.sum(bits(g) != bits(e)))
            raise AssertionError('{} {}: {} of {} values differ, rel L2 {:.3e}'.format(
                context, key, bad, e.size, rel))


@pytest.mark.parametrize('name', sorted(scenarios.SCENARIOS))
def test_scenario_matches_reference_golden_bitwise(library, name):
    field, steps = scenarios.SCENARIOS[name](fds)
    first = steps // 3
    field.simulate(first)                 # segmented exactly like the golden run
    field.simulate(steps - first)
    got = scenarios.collect(field)
    gold = np.load(os.path.join(GOLDEN, name + '.npz'))
    assert_same(got, {k: gold[k] for k in gold.files if k != 'versions'}, name)
    # the wide scenarios exist to pin the streaming kernels to the real reference
    expected_kernel = {'acoustic2d_wide': 'stream2d_kernel<acoustic2d',
                       'acoustic2d_signal_lines': 'stream2d_kernel<acoustic2d',
                       'acoustic2d_lossy_wide': 'streamv_kernel<acoustic2d',
                       'acoustic3daxi_lossy_wide': 'streamv_kernel<acoustic3daxi',
                       'acoustic3daxi_lossless_wide': 'stream2d_kernel<acoustic3daxi',
                       'thermal2d_wide': 'stream2d_kernel<thermal2d',
                       'thermal3daxi_wide': 'stream2d_kernel<thermal3daxi'}.get(name)
    if expected_kernel:
        launches, spl, kernel = field.__dict__['_engine_state'].engine.last_launch_info()
        assert expected_kernel in kernel and spl >= 2, (name, kernel, spl)


@pytest.mark.parametrize('name', ['acoustic1d_lossy', 'acoustic1d_lossless', 'acoustic1d_long',
                                  'thermal1d'])
def test_1d_kernels_match_reference_golden_bitwise(library, name):
    """The default 1-D path is the register-resident warp kernel (fds_line1d.cuh); kernel = 1 selects
    the shared-memory kernel (fds_step1d.cuh). Both must reproduce the reference."""
    gold = np.load(os.path.join(GOLDEN, name + '.npz'))
    expected = {k: gold[k] for k in gold.files if k != 'versions'}
    for kernel, wanted in ((0, 'line1d_kernel'), (1, 'step1d_kernel')):
        field, steps = scenarios.SCENARIOS[name](fds)
        field.device_kernel = kernel
        first = steps // 3
        field.simulate(first)
        field.simulate(steps - first)
        assert_same(scenarios.collect(field), expected, '{} kernel {}'.format(name, kernel))
        assert wanted in field.__dict__['_engine_state'].engine.last_launch_info()[2]


@pytest.mark.parametrize('name', ['acoustic2d_wide', 'acoustic1d_long', 'thermal2d'])
def test_field_without_outputs_still_advances_several_steps_per_launch(library, name):
    """Probe records bound the steps of one chunk; a field without any Output must not fall back to
    one step per launch (and must still match the run that has outputs)."""
    field, steps = scenarios.SCENARIOS[name](fds)
    twin, _ = scenarios.SCENARIOS[name](fds)
    for component in scenarios.component_names(field):
        getattr(field, component).outputs = []
    field.simulate(steps)
    twin.simulate(steps)
    for component in scenarios.component_names(field):
        assert np.array_equal(bits(getattr(field, component).values),
                              bits(getattr(twin, component).values)), component
    launches, spl, kernel = field.__dict__['_engine_state'].engine.last_launch_info()
    if name != 'thermal2d':         # 44 cells wide: one-step kernel
        assert spl >= 4 and launches <= -(-steps // 4) + 1, (kernel, launches, spl)


@pytest.mark.parametrize('name', ['acoustic2d_lossless', 'acoustic1d_lossy', 'thermal2d'])
def test_single_steps_equal_one_run(library, name):
    """sim_step() is the coupling seam (pyfds/coupling.py:81-87): host values must be coherent before
    and after every single step, and n single steps must equal one n-step run."""
    field, steps = scenarios.SCENARIOS[name](fds)
    steps = min(steps, 25)
    for _ in range(steps):
        field.sim_step()
        field.step += 1
    other, _ = scenarios.SCENARIOS[name](fds)
    other.simulate(steps)
    names = scenarios.component_names(field)
    for n in names:
        assert np.array_equal(bits(getattr(field, n).values), bits(getattr(other, n).values)), n


def _vs_oracle(field, steps, context):
    stepper = restate.stepper_for(field)
    stepper.run(steps)
    field.simulate(steps)
    assert_same(scenarios.collect(field), scenarios.collect_stepper(stepper), context)


@pytest.mark.parametrize('lossy', [False, True])
def test_acoustic2d_medium_grid_vs_oracle(library, lossy):
    field, _ = scenarios._acoustic2d(fds, lossy=lossy, nx=512, ny=384, steps=40, seed=11)
    _vs_oracle(field, 40, 'acoustic2d 512x384 lossy={}'.format(lossy))


def test_acoustic2d_odd_sizes_vs_oracle(library):
    field, _ = scenarios._acoustic2d(fds, lossy=False, nx=333, ny=127, steps=50, seed=12)
    _vs_oracle(field, 50, 'acoustic2d 333x127')


def test_axisymmetric_medium_grid_vs_oracle(library):
    field, _ = scenarios._acoustic2d(fds, lossy=True, nx=256, ny=200, steps=30, seed=13,
                                     klass='Acoustic3DAxi')
    _vs_oracle(field, 30, 'acoustic3daxi 256x200')


def test_thermal2d_medium_grid_vs_oracle(library):
    field, _ = scenarios._thermal2d(fds, 'Thermal2D', 320, 256, 60, seed=14)
    _vs_oracle(field, 60, 'thermal2d 320x256')


def test_config2_shape_1024_vs_oracle(library):
    """BASELINE.json config 2 (two materials, point source, rigid line, 4 probes) at 1024^2 with a
    random initial state so that every cell is exercised ('parity variant')."""
    field, _ = scenarios._acoustic2d(fds, lossy=False, nx=1024, ny=1024, steps=20, seed=15)
    _vs_oracle(field, 20, 'acoustic2d 1024^2')


def test_signal_too_short_raises_before_launch(library):
    field, steps = scenarios.acoustic2d_lossless(fds)
    field.simulate(steps)
    with pytest.raises(IndexError):
        field.simulate(1)          # the source signal has exactly `steps` samples


def test_reset_and_rerun_reproduces(library):
    field, steps = scenarios.acoustic1d_lossy(fds)
    field.simulate(steps)
    first = scenarios.collect(field)
    field.reset()
    for component in (field.pressure, field.velocity):
        for output in component.outputs:
            output.signals = []
    field.simulate(steps)
    assert_same(scenarios.collect(field), first, 'rerun after reset')


def test_round_trip_properties_at_full_size(library):
    """Size-independent properties at BASELINE.json's config-2 size (4096^2), where the oracle is too
    slow for the default suite: (1) a zero field with no source stays exactly zero; (2) with a point
    source the run is deterministic and the probes record the source cell exactly as the boundary
    wrote it; (3) segmented and unsegmented runs agree bitwise."""
    n = 4096
    steps = 12
    def build():
        f = fds.Acoustic2D(t_delta=1e-7, t_samples=steps, x_delta=1e-3, x_samples=n, y_delta=1e-3,
                           y_samples=n, material=fds.AcousticMaterial(1500, 1000))
        f.add_material_region(f.get_rect_region((1024 * 1e-3, 1024 * 1e-3, 1024 * 1e-3, 1024 * 1e-3)),
                              fds.AcousticMaterial(1200, 900))
        return f
    quiet = build()
    quiet.simulate(3)
    assert not quiet.pressure.values.any() and not quiet.velocity_x.values.any()

    def run(segments):
        f = build()
        src = f.get_point_region((2048 * 1e-3, 2048 * 1e-3))
        signal = np.sin(0.1 * np.arange(steps)) + 1.0
        f.pressure.add_boundary(src, value=signal)           # not additive: p[src] = signal[step]
        f.pressure.add_output(src)
        f.velocity_x.add_boundary(f.get_line_region((0, 0, 0, (n - 1) * 1e-3)))
        for count in segments:
            f.simulate(count)
        return f, signal
    a, signal = run([steps])
    b, _ = run([5, 4, 3])
    assert np.array_equal(np.asarray(a.pressure.outputs[0].signals[0]), signal)
    for name in ('pressure', 'velocity_x', 'velocity_y'):
        assert np.array_equal(bits(getattr(a, name).values), bits(getattr(b, name).values)), name
    # the disturbance travels at most one cell per step (leapfrog stencil reach)
    p = a.pressure.values.reshape(n, n)
    far = np.ones((n, n), dtype=bool)
    far[2048 - steps - 1:2048 + steps + 2, 2048 - steps - 1:2048 + steps + 2] = False
    assert not p[far].any()
    assert p[2048, 2048 + steps - 2] != 0.0


# ---- streaming multi-step kernel vs one-step kernel -------------------------------------------------

def _stream_case(nx, ny, steps, seed, kernel, klass='Acoustic2D', lossy=False):
    """Boundaries, sources and probes placed on strip seams (x = 119, 120, 121, 240), on the halo
    lanes, on the first/last rows and on chunk seams (coordinates clamped to small grids)."""
    def X(k):
        return min(k, nx - 1) * 1e-3

    def Y(k):
        return min(k, ny - 1) * 1e-3

    f = getattr(fds, klass)(t_delta=1e-7, t_samples=steps, x_delta=1e-3, x_samples=nx,
                            y_delta=1e-3, y_samples=ny,
                            material=fds.AcousticMaterial(1500, 1000,
                                                          shear_viscosity=1e-3 if lossy else 0))
    f.device_kernel = kernel
    f.add_material_region(f.get_rect_region((X(100), Y(10), X(145) - X(100), Y(40) - Y(10))),
                          fds.AcousticMaterial(1200, 900, absorption_coef=7.7 if lossy else None))
    f.add_material_region(f.get_rect_region((X(119), Y(20), X(121) - X(119), Y(60) - Y(20))),
                          fds.AcousticMaterial(1350, 950, absorption_coef=300 if lossy else None))
    scenarios._randomise(f, ('pressure', 'velocity_x', 'velocity_y'), seed)
    top = Y(ny)
    f.velocity_x.add_boundary(f.get_line_region((0, 0, 0, top)))
    f.velocity_x.add_boundary(f.get_line_region((X(120), 0, X(120), top)), value=1e-6,
                              additive=True)
    f.velocity_y.add_boundary(f.get_line_region((0, Y(32), X(nx), Y(32))))
    f.pressure.add_boundary(f.get_line_region((0, 0, X(nx), 0)))
    f.pressure.add_boundary(f.get_line_region((0, top, X(nx), top)), value=2e-4)
    for k, x in enumerate((119, 120, 121, 240, 0, nx - 1)):
        f.pressure.add_boundary(f.get_point_region((X(x), Y(5 + 9 * k))),
                                value=scenarios._pulse(steps, 10 + k, 6), additive=True)
        f.pressure.add_output(f.get_point_region((X(x), Y(5 + 9 * k))))
        f.velocity_x.add_output(f.get_point_region((X(x), Y(6 + 9 * k))))
        f.velocity_y.add_output(f.get_point_region((X(x), Y(32))))
    f.pressure.add_output(f.get_line_region((X(100), Y(31), X(140), Y(31))))
    f.velocity_y.add_output(f.get_point_region((0, 0)))
    f.velocity_y.add_output(f.get_point_region((X(nx), top)))
    return f


@pytest.mark.parametrize('nx,ny,steps', [(256, 70, 23), (128, 33, 9), (484, 150, 16)])
def test_streaming_kernel_equals_one_step_kernel(library, nx, ny, steps):
    results = []
    for kernel in (1, 2):
        f = _stream_case(nx, ny, steps, seed=31, kernel=kernel)
        f.simulate(steps // 2)
        f.simulate(steps - steps // 2)
        results.append(scenarios.collect(f))
        name = f.__dict__['_engine_state'].engine.last_launch_info()[2]
        assert ('stream2d' in name) == (kernel == 2), name
    assert_same(results[1], results[0], 'stream vs step {}x{}'.format(nx, ny))


def test_streaming_kernel_vs_oracle_with_seams(library):
    f = _stream_case(364, 90, 18, seed=32, kernel=2)
    _vs_oracle(f, 18, 'stream2d 364x90')


@pytest.mark.parametrize('max_k,chunk', [(1, 0), (2, 16), (3, 7), (4, 5)])
def test_streaming_kernel_step_counts_and_chunking(library, max_k, chunk, monkeypatch):
    """Every K (steps per launch) and odd chunk heights must give the same bits."""
    monkeypatch.setenv('FDS_MAX_K', str(max_k))
    monkeypatch.setenv('FDS_CHUNK_ROWS', str(chunk))
    f = _stream_case(256, 61, 13, seed=33, kernel=2)
    f.simulate(13)
    g = _stream_case(256, 61, 13, seed=33, kernel=1)
    g.simulate(13)
    assert_same(scenarios.collect(f), scenarios.collect(g), 'K={} chunk={}'.format(max_k, chunk))


# ---- the branch-free (steady-row) variants of the streaming kernel and the rows between them --------

def _steady_case(pattern, nx, ny, steps, seed, kernel, klass='Acoustic2D', lossy=False):
    """Maps laid out for the kernel's geometry (56 owned cells per strip, seams at x = 56, 112, ...):
    long stretches of rows with the same map word per lane -- walls, constant sources and material
    interfaces ALONG y -- separated by a few rows that break them."""
    mm = 1e-3
    f = getattr(fds, klass)(t_delta=1e-7, t_samples=steps, x_delta=mm, x_samples=nx, y_delta=mm,
                            y_samples=ny,
                            material=fds.AcousticMaterial(1500, 1000,
                                                          shear_viscosity=1e-3 if lossy else 0))
    f.device_kernel = kernel
    scenarios._randomise(f, ('pressure', 'velocity_x', 'velocity_y'), seed)
    top = (ny - 1) * mm
    second = fds.AcousticMaterial(1200, 900, absorption_coef=7.7 if lossy else None)
    third = fds.AcousticMaterial(1350, 950, absorption_coef=300 if lossy else None)

    def column(x, y0=0, y1=None):
        return f.get_line_region((x * mm, y0 * mm, x * mm, top if y1 is None else y1 * mm))

    if pattern == 'plain':
        pass
    elif pattern == 'vx_walls':          # component 1: rigid wall at x = 0, constant source on a seam
        f.velocity_x.add_boundary(column(0))
        f.velocity_x.add_boundary(column(56), value=1e-6, additive=True)
        f.velocity_x.add_boundary(column(nx - 1), value=-2e-6)
    elif pattern == 'p_columns':         # component 0: pressure-release columns, one on a halo lane
        f.pressure.add_boundary(column(55))
        f.pressure.add_boundary(column(113), value=3e-4)
        f.pressure.add_boundary(column(170), value=1e-5, additive=True)
    elif pattern == 'vy_columns':        # component 2
        f.velocity_y.add_boundary(column(111), value=2e-6, additive=True)
        f.velocity_y.add_boundary(column(5))
    elif pattern == 'signal_columns':    # sources along y driven by a signal: classes whose value
        f.pressure.add_boundary(column(60), value=scenarios._pulse(steps, 5, 3), additive=True)
        f.pressure.add_boundary(column(nx - 56, 10, ny - 20), value=-scenarios._pulse(steps, 7, 2))
    elif pattern == 'signal_vx':         # is a sample per step (component 1; a row of them: general)
        f.velocity_x.add_boundary(column(113), value=1e-3 * scenarios._pulse(steps, 4, 3),
                                  additive=True)
        f.velocity_x.add_boundary(f.get_line_region((20 * mm, 50 * mm, (nx - 26) * mm, 50 * mm)),
                                  value=1e-3 * scenarios._pulse(steps, 6, 3), additive=True)
    elif pattern == 'interfaces_y':      # several materials, interfaces inside a strip and on seams
        f.add_material_region(f.get_rect_region((50 * mm, 0, 80 * mm, top)), second)
        f.add_material_region(f.get_rect_region((112 * mm, 0, 3 * mm, top)), third)
    elif pattern == 'interface_and_wall':    # operations AND several materials in one strip: general
        f.add_material_region(f.get_rect_region((20 * mm, 0, 60 * mm, top)), second)
        f.velocity_x.add_boundary(column(0))
        f.pressure.add_boundary(column(130), value=1e-4)
    elif pattern == 'two_components':    # operations on two components of one strip: general
        f.velocity_x.add_boundary(column(0))
        f.pressure.add_boundary(column(3))
        f.velocity_y.add_boundary(column(60), value=1e-6, additive=True)
        f.velocity_x.add_boundary(column(62))
    elif pattern == 'partial_height':    # everything starts and ends somewhere: transitions
        f.velocity_x.add_boundary(column(0, 9, ny - 14))
        f.pressure.add_boundary(column(100, 20, 61))
        f.add_material_region(f.get_rect_region((40 * mm, 17 * mm, 90 * mm, 40 * mm)), second)
        f.add_material_region(f.get_rect_region((150 * mm, 30 * mm, 30 * mm, 31 * mm)), third)
        f.velocity_y.add_boundary(f.get_line_region((0, 45 * mm, (nx - 1) * mm, 45 * mm)))
    else:
        raise ValueError(pattern)
    # what every case has: a source with a signal and probes -- single rows of the general path
    f.pressure.add_boundary(f.get_point_region(((nx // 2) * mm, (ny // 2) * mm)),
                            value=scenarios._pulse(steps, 6, 4), additive=True)
    f.pressure.add_output(f.get_point_region((57 * mm, (ny // 3) * mm)))
    f.velocity_x.add_output(f.get_point_region((1 * mm, (ny - 2) * mm)))
    f.velocity_y.add_output(f.get_point_region(((nx - 1) * mm, 1 * mm)))
    return f


STEADY_PATTERNS = ['plain', 'vx_walls', 'p_columns', 'vy_columns', 'interfaces_y',
                   'interface_and_wall', 'two_components', 'partial_height', 'signal_columns',
                   'signal_vx']


@pytest.mark.parametrize('pattern', STEADY_PATTERNS)
@pytest.mark.parametrize('chunk', [0, 23])
def test_streaming_kernel_steady_variants_equal_one_step_kernel(library, pattern, chunk,
                                                                monkeypatch):
    """All five branch-free instantiations, the general rows between them, even and odd numbers of
    rows per task: bit for bit what the one-step kernel computes."""
    monkeypatch.setenv('FDS_CHUNK_ROWS', str(chunk))
    monkeypatch.setenv('FDS_STREAM_STATS', '1')
    results = []
    for kernel in (1, 2):
        f = _steady_case(pattern, 256, 118, 13, seed=71, kernel=kernel)
        f.simulate(9)         # launches of 4 + 4 + 1 steps
        f.simulate(4)
        results.append(scenarios.collect(f))
        engine = f.__dict__['_engine_state'].engine
        name = engine.last_launch_info()[2]
        assert ('stream2d' in name) == (kernel == 2), name
        if kernel == 2:
            stats = engine.stream_stats()
    assert_same(results[1], results[0], 'steady {} chunk={}'.format(pattern, chunk))
    # the case must really have gone through the variant it is named after
    plain, comp0, comp1, comp2, materials, general, rows = stats[:7]
    assert plain > 0 and rows > 0, stats
    expected = {'vx_walls': comp1, 'p_columns': comp0, 'vy_columns': comp2, 'interfaces_y': materials,
                'partial_height': comp1 + comp0 + materials, 'signal_columns': comp0,
                'signal_vx': comp1}
    if pattern in expected:
        assert expected[pattern] > 0, (pattern, stats)
    if pattern in ('plain', 'vx_walls', 'p_columns', 'vy_columns', 'interfaces_y',
                   'signal_columns', 'signal_vx') and chunk == 0:
        assert general < 0.25 * rows, (pattern, stats)     # only the rows around source and probes
    if pattern in ('interface_and_wall', 'two_components'):
        assert general > 0, (pattern, stats)


@pytest.mark.parametrize('pattern', ['vx_walls', 'interfaces_y', 'partial_height', 'signal_columns',
                                     'signal_vx'])
def test_streaming_kernel_steady_variants_vs_oracle(library, pattern):
    f = _steady_case(pattern, 192, 97, 11, seed=72, kernel=2)
    _vs_oracle(f, 11, 'steady {} vs oracle'.format(pattern))


@pytest.mark.parametrize('klass', ['Thermal2D', 'Thermal3DAxi'])
@pytest.mark.parametrize('chunk', [0, 17])
def test_streaming_kernel_thermal_equals_one_step_kernel(library, chunk, klass, monkeypatch):
    """Thermal2D / Thermal3DAxi on the streaming kernel: Dirichlet temperature columns (component 0
    operations), zero-flux rows, an anisotropic material block."""
    monkeypatch.setenv('FDS_CHUNK_ROWS', str(chunk))
    results = []
    for kernel in (1, 2):
        f, steps = scenarios._thermal2d(fds, klass, 256, 101, 14, seed=73)
        f.device_kernel = kernel
        f.simulate(6)
        f.simulate(8)
        results.append(scenarios.collect(f))
        name = f.__dict__['_engine_state'].engine.last_launch_info()[2]
        assert ('stream2d' in name) == (kernel == 2), name
    assert_same(results[1], results[0], 'thermal stream vs step chunk={} {}'.format(chunk, klass))


def test_streaming_kernel_thermal_axisymmetric_vs_oracle(library):
    f, steps = scenarios._thermal2d(fds, 'Thermal3DAxi', 192, 75, 13, seed=78)
    f.device_kernel = 2
    _vs_oracle(f, steps, 'stream2d thermal3daxi 192x75')


# ---- shared-memory tile kernel vs one-step kernel --------------------------------------------------

@pytest.mark.parametrize('builder,args', [
    ('_acoustic2d', dict(lossy=True, nx=264, ny=75, steps=12, seed=51)),
    ('_acoustic2d', dict(lossy=False, nx=136, ny=50, steps=12, seed=52, klass='Acoustic3DAxi')),
    ('_acoustic2d', dict(lossy=True, nx=392, ny=61, steps=10, seed=53, klass='Acoustic3DAxi')),
    ('_thermal2d', dict(klass='Thermal2D', nx=272, ny=131, steps=14, seed=54)),
    ('_thermal2d', dict(klass='Thermal3DAxi', nx=128, ny=70, steps=14, seed=55)),
])
def test_tile_kernel_equals_one_step_kernel(library, builder, args, monkeypatch):
    monkeypatch.setenv('FDS_TILE_ROWS', '9')         # several tiles in y even on small grids
    results = []
    for kernel in (1, 3):
        f, steps = getattr(scenarios, builder)(fds, **args)
        f.device_kernel = kernel
        f.simulate(steps // 2)
        f.simulate(steps - steps // 2)
        results.append(scenarios.collect(f))
        name = f.__dict__['_engine_state'].engine.last_launch_info()[2]
        assert ('tile2d' in name) == (kernel == 3), name
    assert_same(results[1], results[0], 'tile vs step {}'.format(args))


# ---- streaming kernel of the viscous / axisymmetric models vs one-step kernel ---------------------

@pytest.mark.parametrize('klass,lossy,max_k,chunk', [
    ('Acoustic2D', True, 1, 0), ('Acoustic2D', True, 1, 11), ('Acoustic2D', True, 1, 6),
    ('Acoustic3DAxi', False, 1, 0), ('Acoustic3DAxi', True, 1, 0), ('Acoustic3DAxi', True, 1, 9),
    ('Acoustic2D', True, 2, 0), ('Acoustic2D', True, 2, 11), ('Acoustic2D', True, 2, 6),
    ('Acoustic3DAxi', False, 2, 0), ('Acoustic3DAxi', False, 2, 7), ('Acoustic3DAxi', True, 2, 0),
    ('Acoustic3DAxi', True, 2, 9)])
def test_viscous_streaming_kernel_equals_one_step_kernel(library, klass, lossy, max_k, chunk,
                                                         monkeypatch):
    monkeypatch.setenv('FDS_MAX_K', str(max_k))
    monkeypatch.setenv('FDS_CHUNK_ROWS', str(chunk))
    results = []
    for kernel in (1, 2):
        f = _stream_case(376, 83, 15, seed=61, kernel=kernel, klass=klass, lossy=lossy)
        f.simulate(7)
        f.simulate(8)
        results.append(scenarios.collect(f))
        name = f.__dict__['_engine_state'].engine.last_launch_info()[2]
        # lossy models: the viscous kernel; the lossless axisymmetric model: stream2d_kernel<AXI>
        assert (('streamv' if lossy else 'stream2d') in name) == (kernel == 2), name
    assert_same(results[1], results[0], 'streamv vs step {} lossy={}'.format(klass, lossy))


@pytest.mark.parametrize('pattern', STEADY_PATTERNS)
@pytest.mark.parametrize('chunk', [0, 23])
def test_streaming_kernel_axisymmetric_lossless_four_steps(library, pattern, chunk, monkeypatch):
    """Lossless Acoustic3DAxi runs on stream2d_kernel<AXI> with up to 4 steps per launch: per-column
    a_p_vx / r, the divergence of vx * r and the + 0 * vx term (pyfds/acoustics.py:205-225)."""
    monkeypatch.setenv('FDS_CHUNK_ROWS', str(chunk))
    results = []
    for kernel in (1, 2):
        f = _steady_case(pattern, 256, 118, 13, seed=76, kernel=kernel, klass='Acoustic3DAxi')
        f.simulate(9)         # launches of 4 + 4 + 1 steps
        f.simulate(4)
        results.append(scenarios.collect(f))
        launches, spl, name = f.__dict__['_engine_state'].engine.last_launch_info()
        assert ('stream2d_kernel<acoustic3daxi' in name) == (kernel == 2), name
        if kernel == 2:
            assert spl == 4, (name, spl)
    assert_same(results[1], results[0], 'axi lossless K=4 {} chunk={}'.format(pattern, chunk))


def test_streaming_kernel_axisymmetric_lossless_vs_oracle(library):
    f = _stream_case(364, 90, 18, seed=77, kernel=2, klass='Acoustic3DAxi', lossy=False)
    _vs_oracle(f, 18, 'stream2d axi lossless 364x90')


def test_viscous_streaming_kernel_vs_oracle(library):
    f = _stream_case(256, 70, 12, seed=62, kernel=2, klass='Acoustic3DAxi', lossy=True)
    _vs_oracle(f, 12, 'streamv axi lossy 256x70')


@pytest.mark.parametrize('klass,lossy', [('Acoustic2D', True), ('Acoustic3DAxi', False),
                                         ('Acoustic3DAxi', True)])
@pytest.mark.parametrize('pattern', STEADY_PATTERNS)
@pytest.mark.parametrize('max_k,chunk', [(2, 0), (2, 23), (1, 0)])
def test_viscous_streaming_kernel_steady_variants(library, klass, lossy, pattern, max_k, chunk,
                                                  monkeypatch):
    """The branch-free bodies of the viscous / axisymmetric kernel (one material with class operations on
    no component or on exactly one; several materials without operations), the general rows between
    such runs and strips that never qualify: bit for bit what the one-step kernel computes, with the
    kernel's counters proving which path the rows took."""
    monkeypatch.setenv('FDS_MAX_K', str(max_k))
    monkeypatch.setenv('FDS_CHUNK_ROWS', str(chunk))
    monkeypatch.setenv('FDS_STREAM_STATS', '1')
    results = []
    for kernel in (1, 2):
        f = _steady_case(pattern, 256, 118, 9, seed=74, kernel=kernel, klass=klass, lossy=lossy)
        f.simulate(5)         # launches of 2 + 2 + 1 steps
        f.simulate(4)
        results.append(scenarios.collect(f))
        engine = f.__dict__['_engine_state'].engine
        launches, spl, name = engine.last_launch_info()
        assert (('streamv' if lossy else 'stream2d') in name) == (kernel == 2), name
        if kernel == 2:
            assert spl == max_k, (name, spl)
            stats = engine.stream_stats()
    assert_same(results[1], results[0], 'streamv steady {} {} lossy={}'.format(pattern, klass, lossy))
    plain, comp0, comp1, comp2, materials, general, rows = stats[:7]
    assert plain > 0 and rows > 0, stats
    # the lossy axisymmetric model keeps strips of several materials on the general row iteration
    one_material_only = lossy and klass == 'Acoustic3DAxi'
    if one_material_only:
        assert materials == 0, stats
    expected = {'vx_walls': comp1, 'p_columns': comp0, 'vy_columns': comp2,
                'interfaces_y': general if one_material_only else materials,
                'partial_height': comp1 + comp0, 'signal_columns': comp0, 'signal_vx': comp1}
    if pattern in expected:
        assert expected[pattern] > 0, (pattern, stats)
    if pattern in ('plain', 'vx_walls', 'p_columns', 'vy_columns', 'signal_columns',
                   'signal_vx') and chunk == 0:
        assert general < 0.3 * rows, (pattern, stats)
    if pattern in ('interface_and_wall', 'two_components'):
        assert general > 0, (pattern, stats)


@pytest.mark.parametrize('klass,lossy,pattern', [('Acoustic2D', True, 'vx_walls'),
                                                 ('Acoustic3DAxi', True, 'partial_height'),
                                                 ('Acoustic3DAxi', False, 'p_columns'),
                                                 ('Acoustic2D', True, 'signal_columns'),
                                                 ('Acoustic3DAxi', True, 'signal_vx')])
def test_viscous_streaming_kernel_steady_variants_vs_oracle(library, klass, lossy, pattern):
    f = _steady_case(pattern, 192, 97, 11, seed=75, kernel=2, klass=klass, lossy=lossy)
    _vs_oracle(f, 11, 'streamv steady {} {} vs oracle'.format(klass, pattern))


# ---- the axisymmetric kernel's division by r^2 (sv_quotient in fds_streamv.cuh) -------------------------

def _division_ranges_case(kernel, pattern='plain', steps=9, nx=256, ny=118):
    """Lossy Acoustic3DAxi whose velocity_x holds everything the quotient sequence must tell apart:
    zeros of both signs, denormals, values either side of the 2^-493 and 2^477 limits of the fast
    sequence, tiny and ordinary normals -- in row blocks and scattered."""
    f = _steady_case(pattern, nx, ny, steps, seed=91, kernel=kernel, klass='Acoustic3DAxi',
                     lossy=True)
    rng = np.random.default_rng(92)
    vx = np.array(f.velocity_x.values, dtype=np.float64).reshape(ny, nx)
    sign = np.where(rng.random((ny, nx)) < 0.5, -1.0, 1.0)
    mant = 1.0 + rng.random((ny, nx))
    vx[50:57] = 0.0 * sign[50:57]                                   # +0.0 and -0.0
    vx[57:63] = sign[57:63] * rng.integers(1, 2 ** 40, (6, nx)) * 5e-324    # denormals
    expo = rng.integers(-497, -489, (8, nx))
    vx[63:71] = sign[63:71] * np.ldexp(mant[63:71], expo)           # around 2^-493
    vx[71:76] = sign[71:76] * mant[71:76] * 1e-300
    expo = rng.integers(474, 480, (7, nx))
    vx[76:83] = sign[76:83] * np.ldexp(mant[76:83], expo)           # around 2^477
    scatter = rng.random((7, nx))
    tail = vx[83:90]
    tail[scatter < 0.2] = 0.0
    tail[(scatter >= 0.2) & (scatter < 0.3)] = -0.0
    tail[(scatter >= 0.3) & (scatter < 0.4)] = 3e-320
    tail[(scatter >= 0.4) & (scatter < 0.45)] = -2.0 ** -493
    tail[(scatter >= 0.45) & (scatter < 0.5)] = np.nextafter(2.0 ** -493, 0.0)
    f.velocity_x.values = vx.reshape(-1)
    return f


@pytest.mark.parametrize('pattern', ['plain', 'vx_walls', 'signal_vx'])
def test_axisymmetric_division_fast_and_exact_iterations(library, pattern, monkeypatch):
    """Bitwise against the one-step kernel (IEEE division in every cell) and against the CPU
    restatement; the counters prove that both kinds of iteration ran."""
    monkeypatch.setenv('FDS_STREAM_STATS', '1')
    results = []
    for kernel in (1, 2):
        f = _division_ranges_case(kernel, pattern)
        f.simulate(5)
        f.simulate(4)
        results.append(scenarios.collect(f))
        if kernel == 2:
            engine = f.__dict__['_engine_state'].engine
            assert 'streamv' in engine.last_launch_info()[2]
            stats = engine.stream_stats()
    assert_same(results[1], results[0], 'division ranges, streamv vs one-step kernel')
    assert sum(stats[:4]) > 0 and stats[6] > 0, stats
    exact_iterations = stats[7]
    assert 0 < exact_iterations < stats[6] // 2, stats     # some, not all row pairs
    g = _division_ranges_case(2, pattern)
    _vs_oracle(g, 9, 'division ranges vs oracle')


def test_axisymmetric_division_quiet_field_stays_on_fast_iterations(library, monkeypatch):
    """A field at rest (all zeros) with a source: zeros are covered by the fast sequence, so only the
    rows the leading edge of the pulse reaches (values decaying towards the denormals) may take the
    exact iterations."""
    monkeypatch.setenv('FDS_STREAM_STATS', '1')
    f = _steady_case('plain', 256, 118, 12, seed=93, kernel=2, klass='Acoustic3DAxi', lossy=True)
    for name in ('pressure', 'velocity_x', 'velocity_y'):
        getattr(f, name).values = np.zeros(f.num_points)
    stepper = restate.stepper_for(f)
    stepper.run(12)
    f.simulate(12)
    assert_same(scenarios.collect(f), scenarios.collect_stepper(stepper), 'quiet axisymmetric field')
    stats = f.__dict__['_engine_state'].engine.stream_stats()
    assert stats[7] < stats[6] // 4, stats

