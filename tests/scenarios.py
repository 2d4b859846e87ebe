"""Simulation scenarios written against the *pyfds public API only*, so that the very same builder
runs on the real reference (``oracle/gen_golden.py``, in the build container), on ``pyfds_b200`` (the
GPU parity tests) and feeds the CPU restatement in ``oracle/restate.py``.

Every builder takes the package module (``pyfds`` or ``pyfds_b200``) and returns
``(field, n_steps)``; nothing here is random except through the fixed seeds below.
"""

import numpy as np


def _pulse(t_samples, centre, width, omega=0.1):
    k = np.arange(t_samples)
    return np.sin(omega * k) * np.exp(-((k - centre) / width) ** 2)


def _randomise(field, names, seed, scale=1e-3):
    rng = np.random.default_rng(seed)
    for name in names:
        component = getattr(field, name)
        component.values = scale * rng.standard_normal(component.values.shape[0])


def acoustic1d_lossy(fds):
    """Config 1 of BASELINE.json scaled down (doc/ex_acoustics.rst:14-37): viscous fluid, rigid and
    pressure-release ends, additive pulse, one probe; plus a second material in the middle."""
    fld = fds.Acoustic1D(t_delta=1e-7, t_samples=400, x_delta=1e-3, x_samples=300,
                         material=fds.AcousticMaterial(700, 0.01, shear_viscosity=1e-3))
    fld.add_material_region(fld.get_line_region((120e-3, 170e-3)),
                            fds.AcousticMaterial(650, 0.012, bulk_viscosity=2e-3))
    fld.velocity.add_boundary(fld.get_point_region(0))
    fld.pressure.add_boundary(fld.get_point_region(299 * 1e-3))
    fld.pressure.add_boundary(fld.get_point_region(100 * 1e-3), value=_pulse(400, 60, 20),
                              additive=True)
    fld.pressure.add_output(fld.get_point_region(200 * 1e-3))
    fld.velocity.add_output(fld.get_line_region((10 * 1e-3, 14 * 1e-3)))
    return fld, 400


def acoustic1d_lossless(fds):
    """Lossless 1-D line with a random initial state: every cell and both global ends active."""
    fld = fds.Acoustic1D(t_delta=1e-7, t_samples=250, x_delta=1e-3, x_samples=257,
                         material=fds.AcousticMaterial(700, 0.01))
    _randomise(fld, ('pressure', 'velocity'), seed=1)
    fld.pressure.add_output(fld.get_point_region(0))
    fld.pressure.add_output(fld.get_point_region(256 * 1e-3))
    return fld, 250


def acoustic1d_long(fds):
    """Longer than one CTA tile can hold: exercises the multi-tile 1-D path with its halos."""
    fld = fds.Acoustic1D(t_delta=1e-7, t_samples=150, x_delta=1e-3, x_samples=20000,
                         material=fds.AcousticMaterial(700, 0.01, shear_viscosity=1e-3))
    fld.add_material_region(fld.get_line_region((3.9, 4.3)), fds.AcousticMaterial(600, 0.02))
    _randomise(fld, ('pressure', 'velocity'), seed=2)
    fld.velocity.add_boundary(fld.get_point_region(0))
    fld.pressure.add_boundary(fld.get_point_region(3968 * 1e-3), value=_pulse(150, 40, 15),
                              additive=True)
    fld.pressure.add_output(fld.get_point_region(3967 * 1e-3))
    fld.pressure.add_output(fld.get_point_region(3969 * 1e-3))
    fld.velocity.add_output(fld.get_point_region(19999 * 1e-3))
    return fld, 150


def _acoustic2d(fds, lossy, nx=48, ny=40, steps=120, seed=3, klass='Acoustic2D'):
    """Config 2 of BASELINE.json scaled down: two materials, additive point source, rigid line at
    x=0 (doc/ex_acoustics.rst:73-74), four pressure probes; random initial state."""
    main = fds.AcousticMaterial(1500, 1000, shear_viscosity=1e-3 if lossy else 0)
    fld = getattr(fds, klass)(t_delta=1e-7, t_samples=steps, x_delta=1e-3, x_samples=nx,
                              y_delta=1e-3, y_samples=ny, material=main)
    qx, qy = nx // 4, ny // 4
    second = fds.AcousticMaterial(1200, 900, absorption_coef=7.7 if lossy else None)
    fld.add_material_region(fld.get_rect_region((qx * 1e-3, qy * 1e-3, qx * 1e-3, qy * 1e-3)),
                            second)
    _randomise(fld, ('pressure', 'velocity_x', 'velocity_y'), seed=seed)
    fld.pressure.add_boundary(fld.get_point_region(((nx // 2) * 1e-3, (ny // 2) * 1e-3)),
                              value=_pulse(steps, 40, 15), additive=True)
    fld.velocity_x.add_boundary(fld.get_line_region((0, 0, 0, (ny - 1) * 1e-3)))
    for m in range(1, 5):
        fld.pressure.add_output(fld.get_point_region(((m * nx // 5) * 1e-3, (m * ny // 5) * 1e-3)))
    return fld, steps


def acoustic2d_lossless(fds):
    return _acoustic2d(fds, lossy=False)


def acoustic2d_lossy(fds):
    return _acoustic2d(fds, lossy=True)


def acoustic2d_wide(fds):
    """Wide enough (nx = 400) for the streaming kernel to run several strips side by side, odd ny."""
    return _acoustic2d(fds, lossy=False, nx=400, ny=75, steps=60, seed=4)


def acoustic2d_boundaries(fds):
    """Every boundary flavour at once (pyfds/regions.py:136-145), overlapping and in list order, on
    the first and last cell of the grid, with probes on the very same cells."""
    nx, ny, steps = 37, 29, 80
    fld = fds.Acoustic2D(t_delta=1e-7, t_samples=steps, x_delta=1e-3, x_samples=nx,
                         y_delta=1e-3, y_samples=ny, material=fds.AcousticMaterial(1500, 1000))
    _randomise(fld, ('pressure', 'velocity_x', 'velocity_y'), seed=5)
    top = fld.get_line_region((0, (ny - 1) * 1e-3, (nx - 1) * 1e-3, (ny - 1) * 1e-3))
    left = fld.get_line_region((0, 0, 0, (ny - 1) * 1e-3))
    diag = fld.get_line_region((3e-3, 2e-3, 30e-3, 20e-3))
    first, last = fld.get_point_region((0, 0)), fld.get_point_region(((nx - 1) * 1e-3,
                                                                       (ny - 1) * 1e-3))
    # pressure: fixed value, then an additive scalar on an overlapping line, then a signal
    fld.pressure.add_boundary(top, value=0)
    fld.pressure.add_boundary(left, value=2.5e-4, additive=True)
    fld.pressure.add_boundary(diag, value=_pulse(steps, 30, 12), additive=True)
    fld.pressure.add_boundary(last, value=1e-3)
    fld.pressure.add_boundary(first, value=_pulse(steps, 20, 9, omega=0.3))
    # velocity_y: one signal per point, additive; region with a duplicated index (last write wins)
    region = fds.regions.LineRegion([5 + 4 * nx, 6 + 4 * nx, 7 + 4 * nx, 6 + 4 * nx],
                                    (0, 0, 0, 0), 'explicit')
    signals = [_pulse(steps, 25, 10, omega=0.1 * (k + 1)) for k in range(4)]
    fld.velocity_y.add_boundary(region, value=signals, additive=True)
    fld.velocity_x.add_boundary(left)
    fld.velocity_x.add_boundary(top, value=-1e-5, additive=True)
    for component in (fld.pressure, fld.velocity_x, fld.velocity_y):
        component.add_output(first)
        component.add_output(last)
        component.add_output(diag)
    fld.velocity_y.add_output(region)
    return fld, steps


def acoustic3daxi_lossy(fds):
    """Config 3 of BASELINE.json scaled down: viscous medium, lossy sponge regions along three edges
    (the reference has no absorbing boundary), Dirichlet lines, additive line source, line probe."""
    nx, ny, steps = 40, 36, 100
    main = fds.AcousticMaterial(1500, 1000, shear_viscosity=1e-3)
    sponge = fds.AcousticMaterial(1500, 1000, absorption_coef=500)
    fld = fds.Acoustic3DAxi(t_delta=1e-7, t_samples=steps, x_delta=1e-3, x_samples=nx,
                            y_delta=1e-3, y_samples=ny, material=main)
    w = 6
    X, Y = (nx - 1) * 1e-3, (ny - 1) * 1e-3
    fld.add_material_region(fld.get_rect_region(((nx - w) * 1e-3, 0, (w - 1) * 1e-3, Y)), sponge)
    fld.add_material_region(fld.get_rect_region((0, 0, X, (w - 1) * 1e-3)), sponge)
    fld.add_material_region(fld.get_rect_region((0, (ny - w) * 1e-3, X, (w - 1) * 1e-3)), sponge)
    _randomise(fld, ('pressure', 'velocity_x', 'velocity_y'), seed=6)
    for line in ((X, 0, X, Y), (0, 0, X, 0), (0, Y, X, Y)):
        fld.pressure.add_boundary(fld.get_line_region(line))
    fld.velocity_x.add_boundary(fld.get_line_region((0, 0, 0, Y)))
    fld.pressure.add_boundary(fld.get_line_region((0, 18e-3, 8e-3, 18e-3)),
                              value=_pulse(steps, 35, 14), additive=True)
    fld.pressure.add_output(fld.get_line_region((0, 25e-3, 12e-3, 25e-3)))
    fld.velocity_x.add_output(fld.get_point_region((5e-3, 5e-3)))
    return fld, steps


def acoustic3daxi_lossless(fds):
    fld, steps = _acoustic2d(fds, lossy=False, nx=33, ny=41, steps=90, seed=7,
                             klass='Acoustic3DAxi')
    return fld, steps


def _flow_for_periods(periods, ny):
    """Flow velocities (m/s) whose ``flow_t_deltas`` (pyfds/acoustic_flow.py:34: dx / flow / dt cut to
    an integer) are the given periods, repeated over the rows; with dx = 1e-3 and dt = 1e-7."""
    wanted = np.array([periods[n % len(periods)] for n in range(ny)], dtype=float)
    return 1e4 / (wanted + np.sign(wanted) * 0.5)


def _acoustic_flow2d(fds, nx, ny, steps, seed, periods):
    """``AcousticFlow2D`` (pyfds/acoustic_flow.py): the config-2 layout in a medium whose rows move at
    different speeds, one row against the x direction (the shift itself ignores the sign)."""
    import warnings
    main = fds.AcousticMaterial(1500, 1000)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')      # "Flow velocity may be to high" for the period-1 rows
        fld = fds.AcousticFlow2D(_flow_for_periods(periods, ny), t_delta=1e-7, t_samples=steps,
                                 x_delta=1e-3, x_samples=nx, y_delta=1e-3, y_samples=ny,
                                 material=main)
    assert [int(f) for f in fld.flow_t_deltas[:len(periods)]] == list(periods)
    qx, qy = nx // 4, ny // 4
    fld.add_material_region(fld.get_rect_region((qx * 1e-3, qy * 1e-3, qx * 1e-3, qy * 1e-3)),
                            fds.AcousticMaterial(1200, 900))
    _randomise(fld, ('pressure', 'velocity_x', 'velocity_y'), seed=seed)
    fld.pressure.add_boundary(fld.get_point_region(((nx // 2) * 1e-3, (ny // 2) * 1e-3)),
                              value=_pulse(steps, 40, 15), additive=True)
    fld.velocity_x.add_boundary(fld.get_line_region((0, 0, 0, (ny - 1) * 1e-3)))
    for m in range(1, 5):
        fld.pressure.add_output(fld.get_point_region(((m * nx // 5) * 1e-3, (m * ny // 5) * 1e-3)))
    fld.velocity_x.add_output(fld.get_point_region((0, 3e-3)))
    return fld, steps


def acoustic_flow2d(fds):
    """Small grid (one-step kernel); rows moving after every step, every 2nd ... 33rd step."""
    return _acoustic_flow2d(fds, 48, 40, 90, seed=12, periods=(3, 7, 1, 33, -5, 1000, 2, 11))


def acoustic_flow2d_wide(fds):
    """Streaming kernel: launches of 4 steps that have to end where a row moves (steps 8, 10, 12, 16,
    20 ... -> launches of 4, 4, 2, 2, 4 ... steps)."""
    return _acoustic_flow2d(fds, 400, 75, 60, seed=13, periods=(8, 12, -20, 1000, 10))


def _thermal2d(fds, klass, nx, ny, steps, seed):
    """Config 4 of BASELINE.json scaled down: two materials (one anisotropic), Dirichlet temperature
    on x=0 / x=max, adiabatic (zero flux) y=0 / y=max, probes on temperature and flux."""
    fld = getattr(fds, klass)(t_delta=1e-3, t_samples=steps, x_delta=1e-3, x_samples=nx,
                              y_delta=1e-3, y_samples=ny,
                              material=fds.ThermalMaterial(900, 2700, 200))
    fld.add_material_region(
        fld.get_rect_region(((nx // 3) * 1e-3, (ny // 3) * 1e-3, (nx // 4) * 1e-3,
                             (ny // 4) * 1e-3)), fds.ThermalMaterial(450, 7800, (50, 30)))
    rng = np.random.default_rng(seed)
    fld.temperature.values = 20 + rng.standard_normal(nx * ny)
    X, Y = (nx - 1) * 1e-3, (ny - 1) * 1e-3
    fld.temperature.add_boundary(fld.get_line_region((0, 0, 0, Y)), value=100)
    fld.temperature.add_boundary(fld.get_line_region((X, 0, X, Y)), value=0)
    fld.heat_flux_y.add_boundary(fld.get_line_region((0, 0, X, 0)), value=0)
    fld.heat_flux_y.add_boundary(fld.get_line_region((0, Y, X, Y)), value=0)
    fld.heat_flux_x.add_boundary(fld.get_point_region((5e-3, 5e-3)), value=1e-2, additive=True)
    for m in range(1, 5):
        fld.temperature.add_output(
            fld.get_point_region(((m * nx // 5) * 1e-3, (m * ny // 5) * 1e-3)))
    fld.heat_flux_x.add_output(fld.get_line_region((1e-3, 3e-3, 6e-3, 3e-3)))
    fld.heat_flux_y.add_output(fld.get_point_region((0, 0)))
    return fld, steps


def thermal2d(fds):
    return _thermal2d(fds, 'Thermal2D', 44, 32, 150, seed=8)


def thermal3daxi(fds):
    return _thermal2d(fds, 'Thermal3DAxi', 30, 35, 120, seed=9)


def thermal1d(fds):
    fld = fds.Thermal1D(t_delta=1e-3, t_samples=300, x_delta=1e-3, x_samples=181,
                        material=fds.ThermalMaterial(900, 2700, 200))
    fld.add_material_region(fld.get_line_region((60e-3, 99e-3)), fds.ThermalMaterial(450, 7800, 50))
    rng = np.random.default_rng(10)
    fld.temperature.values = 20 + rng.standard_normal(181)
    fld.temperature.add_boundary(fld.get_point_region(0), value=100)
    fld.heat_flux.add_boundary(fld.get_point_region(180e-3), value=0)
    fld.temperature.add_output(fld.get_line_region((50e-3, 53e-3)))
    fld.heat_flux.add_output(fld.get_point_region(90e-3))
    return fld, 300


# ---- wide grids (nx >= 128, a multiple of 4): every model on its streaming kernel ------------------

def acoustic2d_lossy_wide(fds):
    """Viscous streaming kernel (2 steps per launch, three strips side by side)."""
    return _acoustic2d(fds, lossy=True, nx=160, ny=53, steps=45, seed=13)


def acoustic3daxi_lossy_wide(fds):
    return _acoustic2d(fds, lossy=True, nx=144, ny=47, steps=41, seed=14, klass='Acoustic3DAxi')


def acoustic3daxi_lossless_wide(fds):
    """Axisymmetric instantiation of the lossless streaming kernel (4 steps per launch)."""
    return _acoustic2d(fds, lossy=False, nx=152, ny=49, steps=43, seed=15, klass='Acoustic3DAxi')


def thermal2d_wide(fds):
    return _thermal2d(fds, 'Thermal2D', 136, 45, 50, seed=16)


def thermal3daxi_wide(fds):
    return _thermal2d(fds, 'Thermal3DAxi', 140, 43, 50, seed=17)


def acoustic2d_signal_lines(fds):
    """Source lines driven by ONE signal each (the device applies them as classes, not through the
    boundary table): a transducer face along y, one along x, two lines sharing a signal, a
    non-additive signal line on velocity_x, and a per-point signal line (table path) across them."""
    nx, ny, steps = 168, 57, 48
    mm = 1e-3
    fld = fds.Acoustic2D(t_delta=1e-7, t_samples=steps, x_delta=mm, x_samples=nx, y_delta=mm,
                         y_samples=ny, material=fds.AcousticMaterial(1500, 1000))
    fld.add_material_region(fld.get_rect_region((100 * mm, 10 * mm, 30 * mm, 30 * mm)),
                            fds.AcousticMaterial(1200, 900))
    _randomise(fld, ('pressure', 'velocity_x', 'velocity_y'), seed=18)
    top = (ny - 1) * mm
    burst = _pulse(steps, 20, 8)
    fld.pressure.add_boundary(fld.get_line_region((0, 5 * mm, 0, 45 * mm)), value=burst,
                              additive=True)
    fld.pressure.add_boundary(fld.get_line_region((60 * mm, 0, 60 * mm, top)), value=burst.copy(),
                              additive=True)
    fld.pressure.add_boundary(fld.get_line_region((10 * mm, 30 * mm, 150 * mm, 30 * mm)),
                              value=_pulse(steps, 12, 5), additive=True)
    fld.velocity_x.add_boundary(fld.get_line_region((111 * mm, 0, 111 * mm, top)),
                                value=1e-4 * _pulse(steps, 25, 10))
    fld.velocity_y.add_boundary(fld.get_line_region((20 * mm, 50 * mm, 27 * mm, 50 * mm)),
                                value=[1e-4 * _pulse(steps, 15 + k, 6) for k in range(8)],
                                additive=True)
    fld.velocity_x.add_boundary(fld.get_line_region((nx * mm - mm, 0, nx * mm - mm, top)))
    for m in range(1, 4):
        fld.pressure.add_output(fld.get_point_region(((m * nx // 4) * mm, (m * ny // 4) * mm)))
    fld.velocity_x.add_output(fld.get_line_region((58 * mm, 20 * mm, 62 * mm, 20 * mm)))
    return fld, steps


SCENARIOS = {
    'acoustic1d_lossy': acoustic1d_lossy,
    'acoustic1d_lossless': acoustic1d_lossless,
    'acoustic1d_long': acoustic1d_long,
    'acoustic2d_lossless': acoustic2d_lossless,
    'acoustic2d_lossy': acoustic2d_lossy,
    'acoustic2d_wide': acoustic2d_wide,
    'acoustic2d_boundaries': acoustic2d_boundaries,
    'acoustic3daxi_lossy': acoustic3daxi_lossy,
    'acoustic3daxi_lossless': acoustic3daxi_lossless,
    'thermal2d': thermal2d,
    'thermal3daxi': thermal3daxi,
    'thermal1d': thermal1d,
    'acoustic_flow2d': acoustic_flow2d,
    'acoustic_flow2d_wide': acoustic_flow2d_wide,
    'acoustic2d_lossy_wide': acoustic2d_lossy_wide,
    'acoustic3daxi_lossy_wide': acoustic3daxi_lossy_wide,
    'acoustic3daxi_lossless_wide': acoustic3daxi_lossless_wide,
    'thermal2d_wide': thermal2d_wide,
    'thermal3daxi_wide': thermal3daxi_wide,
    'acoustic2d_signal_lines': acoustic2d_signal_lines,
}


def component_names(field):
    for names in (('pressure', 'velocity'), ('pressure', 'velocity_x', 'velocity_y'),
                  ('temperature', 'heat_flux'), ('temperature', 'heat_flux_x', 'heat_flux_y')):
        if all(hasattr(field, n) for n in names) and len(names) == (2 if not hasattr(field, 'y')
                                                                   else 3):
            return names
    raise TypeError('unknown field type')


def collect(field):
    """Final values and probe signals of a field that was stepped through the pyfds API, as a flat
    dict of arrays (the golden file layout)."""
    out = {}
    for name in component_names(field):
        component = getattr(field, name)
        out['values/' + name] = np.asarray(component.values, dtype=np.float64)
        for k, output in enumerate(component.outputs):
            out['signals/{}/{}'.format(name, k)] = np.asarray(output.signals, dtype=np.float64)
    out['step'] = np.asarray(field.step)
    return out


def collect_stepper(stepper):
    """Same layout from an ``oracle.restate`` stepper."""
    out = {}
    for name in stepper.components:
        out['values/' + name] = stepper.values(name)
        for k, signals in enumerate(stepper.signals(name)):
            out['signals/{}/{}'.format(name, k)] = signals
    out['step'] = np.asarray(stepper.step)
    return out


# ---- region index maps (pyfds/fields.py:391-533) -------------------------------------------------

def region_cases(fds):
    """Named region constructions whose ``indices`` (values *and* order) are pinned by goldens."""
    f = fds.fields.Field2D(61, 1e-3, 47, 5e-4, 10, 1.0, int(5))
    g = fds.fields.Field1D(50, 0.1, 10, 1.0, int(5))
    cases = {
        'line_h': f.get_line_region((2e-3, 3e-3, 40e-3, 3e-3)),
        'line_h_rev': f.get_line_region((40e-3, 3e-3, 2e-3, 3e-3)),
        'line_v': f.get_line_region((7e-3, 1e-3, 7e-3, 20e-3)),
        'line_diag': f.get_line_region((0, 0, 23e-3, 23 * 5e-4)),
        'line_diag_rev': f.get_line_region((30e-3, 2 * 5e-4, 10e-3, 22 * 5e-4)),
        'line_shallow': f.get_line_region((1e-3, 1e-3, 58e-3, 9e-3)),
        'line_steep': f.get_line_region((50e-3, 22e-3, 41e-3, 5e-4)),
        'line_half': f.get_line_region((0, 0, 10e-3, 5 * 5e-4)),
        'rect': f.get_rect_region((5e-3, 2e-3, 12e-3, 6e-3)),
        'rect_neg': f.get_rect_region((30e-3, 10e-3, -7e-3, -4e-3)),
        'rect_full': f.get_rect_region((0, 0, 60e-3, 23e-3)),
        'tri_cw': f.get_tri_region((5e-3, 2e-3, 30e-3, 20e-3, 50e-3, 4e-3)),
        'tri_ccw': f.get_tri_region((5e-3, 2e-3, 50e-3, 4e-3, 30e-3, 20e-3)),
        'ellipse': f.get_ellipse_region((30e-3, 11e-3), (12e-3, 6e-3)),
        'circle': f.get_ellipse_region((25e-3, 10e-3), 5e-3),
        'point': f.get_point_region((13e-3, 7 * 5e-4)),
        'line_1d': g.get_line_region((0.3, 2.5)),
        'point_1d': g.get_point_region(4.9),
    }
    return cases


# ---- coupled fields (pyfds/coupling.py, pyfds/coupled_fields.py) -----------------------------------
# Builders return (group, n_steps); `collect_group` flattens every member field.

def linear_transfer(fds, scale):
    """``lambda values: scale * values`` -- as the object the drop-in recognises as a linear transfer
    function where it offers one (``pyfds_b200.coupling.linear``), as the plain callable elsewhere."""
    factory = getattr(getattr(fds, 'coupling', None), 'linear', None)
    if factory is not None:
        return factory(scale)
    return lambda values: scale * values


def _thermoacoustic(fds, stepping, steps=240, nx=300):
    grp = fds.ThermoAcoustic1D(x_samples=nx, x_delta=1e-3, t_samples=steps, t_delta=1e-7,
                               thermal_material=fds.ThermalMaterial(900, 2700, 200),
                               acoustic_material=fds.AcousticMaterial(700, 0.01,
                                                                     shear_viscosity=1e-3),
                               stepping=stepping)
    sound, heat = grp.fields
    sound.add_material_region(sound.get_line_region((120e-3, 170e-3)),
                              fds.AcousticMaterial(650, 0.012, bulk_viscosity=2e-3))
    heat.add_material_region(heat.get_line_region((140e-3, 199e-3)),
                             fds.ThermalMaterial(450, 7800, 50))
    # The heating of one step is ~1e-25 K at these amplitudes (the reference's expression carries
    # dt / dx twice): the temperature starts at exactly zero so that every bit of it is the coupling's.
    _randomise(sound, ('pressure', 'velocity'), seed=31, scale=1.0)
    sound.velocity.add_boundary(sound.get_point_region(0))
    sound.pressure.add_boundary(sound.get_point_region((nx - 1) * 1e-3))
    sound.pressure.add_boundary(sound.get_point_region(100e-3), value=_pulse(steps, 60, 20),
                                additive=True)
    heat.temperature.add_boundary(heat.get_point_region(0), value=0)
    sound.pressure.add_output(sound.get_point_region(200e-3))
    heat.temperature.add_output(heat.get_line_region((98e-3, 103e-3)))
    heat.heat_flux.add_output(heat.get_point_region(150e-3))
    return grp, steps


def thermoacoustic1d(fds):
    """The preset of pyfds/coupled_fields.py: viscous losses of the sound field heat the medium,
    delivered after every step."""
    return _thermoacoustic(fds, stepping=1)


def thermoacoustic1d_stepping(fds):
    """... accumulated over three steps between deliveries (accumulate = True)."""
    return _thermoacoustic(fds, stepping=3)


def _two_line_fields(fds, steps, nx=220):
    sound = fds.Acoustic1D(t_delta=1e-7, t_samples=steps, x_delta=1e-3, x_samples=nx,
                           material=fds.AcousticMaterial(700, 0.01, shear_viscosity=1e-3))
    heat = fds.Thermal1D(t_delta=1e-3, t_samples=steps, x_delta=1e-3, x_samples=nx,
                         material=fds.ThermalMaterial(900, 2700, 200))
    heat.add_material_region(heat.get_line_region((60e-3, 99e-3)), fds.ThermalMaterial(450, 7800, 50))
    _randomise(sound, ('pressure', 'velocity'), seed=33, scale=1e-2)
    rng = np.random.default_rng(34)
    heat.temperature.values = 20 + 3 * np.sin(np.arange(nx) / 17.0) + 0.1 * rng.standard_normal(nx)
    sound.velocity.add_boundary(sound.get_point_region(0))
    sound.pressure.add_boundary(sound.get_point_region(70e-3), value=_pulse(steps, 50, 18),
                                additive=True)
    heat.temperature.add_boundary(heat.get_point_region(0), value=40)
    sound.pressure.add_output(sound.get_point_region(150e-3))
    sound.velocity.add_output(sound.get_point_region(151e-3))
    heat.temperature.add_output(heat.get_point_region(30e-3))
    return sound, heat


def boundary_coupling_linear(fds):
    """Two BoundaryCouplings with linear transfer functions (pyfds/coupling.py:90-140): temperature
    feeds the pressure additively after every step; the velocity REPLACES the heat flux every fourth
    step with what accumulated in between."""
    steps = 150
    sound, heat = _two_line_fields(fds, steps)
    a = fds.BoundaryCoupling(heat.temperature, sound.pressure, linear_transfer(fds, 1e-4))
    b = fds.BoundaryCoupling(sound.velocity, heat.heat_flux, linear_transfer(fds, -2.5),
                             additive=False, accumulate=True, stepping=4)
    return fds.SynchronizedFields([sound, heat], [a, b]), steps


def boundary_coupling_stepping(fds):
    """additive with stepping 3 and no accumulation; non-additive after every step."""
    steps = 100
    sound, heat = _two_line_fields(fds, steps)
    a = fds.BoundaryCoupling(heat.temperature, sound.velocity, linear_transfer(fds, 3e-6),
                             additive=True, accumulate=False, stepping=3)
    b = fds.BoundaryCoupling(sound.pressure, heat.heat_flux, linear_transfer(fds, 0.5),
                             additive=False)
    return fds.SynchronizedFields([sound, heat], [a, b]), steps


def material_coupling_exponential(fds):
    """MaterialCouplingExponential (pyfds/coupling.py:218-259): the smooth temperature profile scales
    the sound velocity point by point -- every cell its own material -- re-assembled every 5th step."""
    steps = 90
    sound, heat = _two_line_fields(fds, steps)
    law = fds.MaterialCouplingExponential(heat.temperature, sound, 'sound_velocity', a=0.8, b=-0.01,
                                          stepping=5)
    return fds.SynchronizedFields([sound, heat], [law]), steps


def material_coupling_powerlaw(fds):
    """MaterialCouplingPowerLaw (pyfds/coupling.py:262-300) on the density of the thermal field, driven
    by the pressure, re-assembled only when the factors moved by more than 2 %."""
    steps = 90
    sound, heat = _two_line_fields(fds, steps)
    law = fds.MaterialCouplingPowerLaw(sound.pressure, heat, 'density', power=2, factor=40.0,
                                       rel_change_threshold=0.02)
    return fds.SynchronizedFields([sound, heat], [law]), steps


def material_coupling_2d(fds):
    """A 2-D target: the smooth temperature field of a Thermal2D scales the sound velocity of an
    Acoustic2D cell by cell (thousands of distinct materials), re-assembled every 4th step; a lossy
    main material, so that the 5-diagonal operator is per-cell as well."""
    steps, nx, ny = 30, 136, 44
    axes = dict(t_samples=steps, x_delta=1e-3, x_samples=nx, y_delta=1e-3, y_samples=ny)
    sound = fds.Acoustic2D(t_delta=1e-7, material=fds.AcousticMaterial(1500, 1000,
                                                                      shear_viscosity=1e-3), **axes)
    heat = fds.Thermal2D(t_delta=1e-3, material=fds.ThermalMaterial(900, 2700, 200), **axes)
    sound.add_material_region(sound.get_rect_region((30e-3, 10e-3, 40e-3, 20e-3)),
                              fds.AcousticMaterial(1200, 900, absorption_coef=7.7))
    _randomise(sound, ('pressure', 'velocity_x', 'velocity_y'), seed=35)
    yy, xx = np.mgrid[0:ny, 0:nx]
    heat.temperature.values = (20 + 5 * np.sin(xx / 9.0) * np.cos(yy / 7.0)).reshape(-1)
    sound.pressure.add_boundary(sound.get_point_region((68e-3, 22e-3)), value=_pulse(steps, 12, 5),
                                additive=True)
    sound.velocity_x.add_boundary(sound.get_line_region((0, 0, 0, (ny - 1) * 1e-3)))
    heat.temperature.add_boundary(heat.get_line_region((0, 0, 0, (ny - 1) * 1e-3)), value=40)
    sound.pressure.add_output(sound.get_point_region((100e-3, 30e-3)))
    heat.temperature.add_output(heat.get_point_region((10e-3, 10e-3)))
    law = fds.MaterialCouplingExponential(heat.temperature, sound, 'sound_velocity', a=0.8, b=-0.01,
                                          stepping=4)
    return fds.SynchronizedFields([sound, heat], [law]), steps


COUPLED_SCENARIOS = {
    'material_coupling_2d': material_coupling_2d,
    'thermoacoustic1d': thermoacoustic1d,
    'thermoacoustic1d_stepping': thermoacoustic1d_stepping,
    'boundary_coupling_linear': boundary_coupling_linear,
    'boundary_coupling_stepping': boundary_coupling_stepping,
    'material_coupling_exponential': material_coupling_exponential,
    'material_coupling_powerlaw': material_coupling_powerlaw,
}


def collect_group(group):
    """Final values and probe signals of every member of a ``SynchronizedFields`` group."""
    out = {}
    for f, field in enumerate(group.fields):
        for key, value in collect(field).items():
            out['field{}/{}'.format(f, key)] = value
    return out
