"""The reference's own unit tests (test/test_fields.py, test/test_acoustics.py, test/test_regions.py)
restated against the drop-in: same constructions, same expected values. CPU only -- they cover the host
logic (axes, regions, materials, boundary/probe semantics, lazy scipy operators)."""

import pickle

import numpy as np
import pytest

import pyfds_b200 as fds
import pyfds_b200.fields as fls
import pyfds_b200.regions as reg


# ---- test/test_fields.py ---------------------------------------------------------------------

def test_dimension():                                              # test_fields.py:5-9
    dim = fls.Dimension(3, 0.1)
    assert np.allclose(dim.vector, [0, 0.1, 0.2])
    assert dim.get_index(0.1) == 1


def test_dimension_snap_failures():                                # fields.py:568-569
    dim = fls.Dimension(1000, 1e-3)
    assert dim.get_index(999 * 1e-3) == 999
    with pytest.raises(AssertionError):
        dim.get_index(0.5e-3)
    with pytest.raises(AssertionError):
        dim.get_index(2.0)


def test_boundary_scalar():                                        # test_fields.py:12-18
    fc = fls.FieldComponent(100)
    fc.values = np.random.rand(100)
    fc.boundaries = [reg.Boundary(reg.LineRegion([5, 6, 7], [0, 0.2], 'test boundary'))]
    fc.boundaries[0].value = 23
    fc.apply_bounds(step=0)
    assert np.allclose(fc.values[[5, 6, 7]], [23, 23, 23])


def test_boundary_signal_per_point_additive():                     # test_fields.py:21-28
    fc = fls.FieldComponent(100)
    fc.values = np.ones(100)
    fc.boundaries = [reg.Boundary(reg.LineRegion([5, 6, 7], [0, 0.2], 'test boundary'))]
    fc.boundaries[0].value = [np.arange(0, 3) * 23, np.arange(0, 3) * 42, np.arange(0, 3) * 23]
    fc.boundaries[0].additive = True
    fc.apply_bounds(step=1)
    assert np.allclose(fc.values[[5, 6, 7]], [24, 43, 24])


def test_boundary_single_signal_additive():                        # test_fields.py:31-38
    fc = fls.FieldComponent(100)
    fc.values = np.ones(100)
    fc.boundaries = [reg.Boundary(reg.LineRegion([5, 6, 7], [0, 0.2], 'test boundary'))]
    fc.boundaries[0].value = np.arange(0, 3) * 23
    fc.boundaries[0].additive = True
    fc.apply_bounds(step=2)
    assert np.allclose(fc.values[[5, 6, 7]], [47, 47, 47])


def test_output_layout():                                          # test_fields.py:41-47
    fc = fls.FieldComponent(100)
    fc.outputs = [reg.Output(reg.LineRegion([0, 1, 2], [0, 0.2], 'test output'))]
    fc.write_outputs()
    fc.write_outputs()
    assert np.allclose(fc.outputs[0].signals, [[0, 0], [0, 0], [0, 0]])
    assert np.allclose(fc.outputs[0].mean_signal, np.zeros(2))


def test_material_vector_duck_typed():                             # test_fields.py:50-54, 70-75
    fld = fls.Field1D(100, 0.1, 100, 0.1, int(5))
    assert np.allclose(fld.material_vector('real'), 5)
    fld = fls.Field2D(100, 0.1, 100, 0.1, 100, 0.1, int(5))
    assert np.allclose(fld.material_vector('real'), 5)
    assert np.size(fld.material_vector('real')) == 10000
    with pytest.raises(KeyError):
        fld.material_vector('no_such_parameter')


def test_operators_1d():                                           # test_fields.py:57-67
    fld = fls.Field1D(3, 1, 3, 1, int(5))
    assert np.allclose(fld.d_x().toarray(), [[-1, 1, 0], [0, -1, 1], [0, 0, -1]])
    assert np.allclose(fld.d_x(variant='backward').toarray(), [[1, 0, 0], [-1, 1, 0], [0, -1, 1]])
    assert np.allclose(fld.d_x(variant='central').toarray(),
                       [[0, 0.5, 0], [-0.5, 0, 0.5], [0, -0.5, 0]])
    assert np.allclose(fld.d_x2().toarray(), [[-2, 1, 0], [1, -2, 1], [0, 1, -2]])
    with pytest.raises(ValueError):
        fld.d_x(variant='sideways')


def test_operators_2d():                                           # test_fields.py:78-107
    fld = fls.Field2D(2, 1, 2, 1, 10, 1, int(5))
    assert np.allclose(fld.d_x().toarray(),
                       [[-1, 1, 0, 0], [0, -1, 1, 0], [0, 0, -1, 1], [0, 0, 0, -1]])
    assert np.allclose(fld.d_x(variant='backward').toarray(),
                       [[1, 0, 0, 0], [-1, 1, 0, 0], [0, -1, 1, 0], [0, 0, -1, 1]])
    assert np.allclose(fld.d_x(variant='central').toarray(),
                       [[0, 0.5, 0, 0], [-0.5, 0, 0.5, 0], [0, -0.5, 0, 0.5], [0, 0, -0.5, 0]])
    assert np.allclose(fld.d_x2().toarray(),
                       [[-2, 1, 0, 0], [1, -2, 1, 0], [0, 1, -2, 1], [0, 0, 1, -2]])
    assert np.allclose(fld.d_y().toarray(),
                       [[-1, 0, 1, 0], [0, -1, 0, 1], [0, 0, -1, 0], [0, 0, 0, -1]])
    assert np.allclose(fld.d_y(variant='backward').toarray(),
                       [[1, 0, 0, 0], [0, 1, 0, 0], [-1, 0, 1, 0], [0, -1, 0, 1]])
    assert np.allclose(fld.d_y(variant='central').toarray(),
                       [[0, 0, 0.5, 0], [0, 0, 0, 0.5], [-0.5, 0, 0, 0], [0, -0.5, 0, 0]])
    assert np.allclose(fld.d_y2().toarray(),
                       [[-2, 0, 1, 0], [0, -2, 0, 1], [1, 0, -2, 0], [0, 1, 0, -2]])


def test_index_position():                                         # test_fields.py:110-122
    fld = fls.Field2D(4, 0.1, 3, 0.1, 1, 1, int(5))
    assert fld.get_index((0.2, 0.1)) == 6
    assert np.allclose(fld.get_position(fld.get_index((0.2, 0.1))), (0.2, 0.1))
    line = fls.Field1D(4, 0.1, 1, 1, int(5))
    assert np.allclose(line.get_position(line.get_index(0.1)), 0.1)


def test_line_regions():                                           # test_fields.py:125-139
    fld = fls.Field1D(4, 0.1, 1, 1, int(5))
    fld.material_regions.append(reg.MaterialRegion(fld.get_line_region((0.1, 0.2)), int(23)))
    assert np.allclose(fld.material_vector('real'), [5, 23, 23, 5])
    fld = fls.Field2D(3, 1, 4, 0.5, 1, 1, int(5))
    assert list(fld.get_line_region((1, 0, 1, 1.5)).indices) == [1, 4, 7, 10]
    assert list(fld.get_line_region((0, 0, 2, 0)).indices) == [0, 1, 2]
    assert list(fld.get_line_region((0, 0, 2, 1.5)).indices) == [0, 4, 7, 11]
    assert list(fld.get_line_region((0, 1.5, 2, 0)).indices) == [9, 7, 4, 2]


def test_rect_regions():                                           # test_fields.py:142-148
    fld = fls.Field2D(3, 1, 4, 0.5, 1, 1, int(5))
    assert list(fld.get_rect_region((0, 0, 1, 1)).indices) == [0, 3, 6, 1, 4, 7]
    assert list(fld.get_rect_region((2, 1.5, -1, -1)).indices) == [4, 7, 10, 5, 8, 11]


# ---- test/test_regions.py ----------------------------------------------------------------------

def test_output_mean_signal():                                     # test_regions.py:5-8
    out = reg.Output(reg.LineRegion([0, 1, 2], [0, 0.2], 'test output'))
    out.signals = [np.linspace(0, 1) for _ in range(len(out.region.indices))]
    assert np.allclose(out.mean_signal, np.linspace(0, 1))


# ---- test/test_acoustics.py --------------------------------------------------------------------

def test_acoustic_material():                                      # test_acoustics.py:5-16
    water = fds.AcousticMaterial(1500, 1000)
    water.bulk_viscosity = 1e-3
    water.shear_viscosity = 1e-3
    assert np.isclose(water.absorption_coef, 7e-3 / 3)
    water.absorption_coef = 3e-3
    assert np.isclose(water.absorption_coef, 3e-3)
    assert np.isclose(fds.AcousticMaterial(1500, 1000, absorption_coef=2e-3).absorption_coef, 2e-3)
    # a falsy override falls back to the derived value (acoustics.py:276)
    assert fds.AcousticMaterial(1500, 1000, shear_viscosity=3, absorption_coef=0).absorption_coef == 4


def test_acoustic1d_matrices():                                    # test_acoustics.py:19-26
    fld = fds.Acoustic1D(t_delta=1, t_samples=1, x_delta=1, x_samples=3,
                         material=fds.AcousticMaterial(700, 0.01, bulk_viscosity=1))
    assert fld.a_p_v is None
    fld.assemble_matrices()
    assert fld.matrices_assembled
    assert np.allclose(fld.a_p_v.toarray(), [[-4900, 4900, 0], [0, -4900, 4900], [0, 0, -4900]])
    assert np.allclose(fld.a_v_p.toarray(), [[100, 0, 0], [-100, 100, 0], [0, -100, 100]])
    assert np.allclose(fld.a_v_v.toarray(), [[-200, 100, 0], [100, -200, 100], [0, 100, -200]])


def test_acoustic2d_matrices():                                    # test_acoustics.py:29-46
    fld = fds.Acoustic2D(t_delta=1, t_samples=1, x_delta=1, x_samples=2, y_delta=1, y_samples=2,
                         material=fds.AcousticMaterial(700, 0.01, bulk_viscosity=1))
    fld.assemble_matrices()
    assert np.allclose(fld.a_p_vx.toarray(), [[-4900, 4900, 0, 0], [0, -4900, 4900, 0],
                                              [0, 0, -4900, 4900], [0, 0, 0, -4900]])
    assert np.allclose(fld.a_p_vy.toarray(), [[-4900, 0, 4900, 0], [0, -4900, 0, 4900],
                                              [0, 0, -4900, 0], [0, 0, 0, -4900]])
    assert np.allclose(fld.a_vx_p.toarray(), [[100, 0, 0, 0], [-100, 100, 0, 0],
                                              [0, -100, 100, 0], [0, 0, -100, 100]])
    assert np.allclose(fld.a_vy_p.toarray(), [[100, 0, 0, 0], [0, 100, 0, 0],
                                              [-100, 0, 100, 0], [0, -100, 0, 100]])
    expected = [[-400, 100, 100, 0], [100, -400, 100, 100], [100, 100, -400, 100],
                [0, 100, 100, -400]]
    assert np.allclose(fld.a_vx_vx.toarray(), expected)
    assert np.allclose(fld.a_vy_vy.toarray(), expected)


def test_acoustic3daxi_matrices():                                 # test_acoustics.py:49-66
    fld = fds.Acoustic3DAxi(t_delta=1, t_samples=1, x_delta=1, x_samples=2, y_delta=1,
                            y_samples=2, material=fds.AcousticMaterial(1, 1, bulk_viscosity=1))
    fld.assemble_matrices()
    assert np.allclose(fld.a_p_vx.toarray(), [[-2, 2 / 3, 0, 0], [0, -2 / 3, 2, 0],
                                              [0, 0, -2, 2 / 3], [0, 0, 0, -2 / 3]])
    assert np.allclose(fld.a_p_vy.toarray(), [[-1, 0, 1, 0], [0, -1, 0, 1],
                                              [0, 0, -1, 0], [0, 0, 0, -1]])
    assert np.allclose(fld.a_vx_p.toarray(), [[1, 0, 0, 0], [-1, 1, 0, 0], [0, -1, 1, 0],
                                              [0, 0, -1, 1]])
    assert np.allclose(fld.a_vy_p.toarray(), [[1, 0, 0, 0], [0, 1, 0, 0], [-1, 0, 1, 0],
                                              [0, -1, 0, 1]])
    expected = [[-4, 4 / 3, 1, 0], [0, -4, 2, 1], [1, 2 / 3, -4, 4 / 3], [0, 1, 0, -4]]
    assert np.allclose(fld.a_vx_vx.toarray(), expected)
    assert np.allclose(fld.a_vy_vy.toarray(), expected)
    assert fds.AcousticAxisymmetric is fds.Acoustic3DAxi


def test_thermal1d_matrix():                                       # test_coupling.py:52
    ths = fds.Thermal1D(t_delta=1, t_samples=1, x_delta=1, x_samples=3,
                        material=fds.ThermalMaterial(1, 1, 1))
    ths.assemble_matrices()
    assert np.allclose(ths.a_t_q.toarray(), [[-1, 1, 0], [0, -1, 1], [0, 0, -1]])


def test_thermal_material_conductivity():                          # thermal.py:199-211
    assert fds.ThermalMaterial(1, 2, 3).thermal_conductivity == (3, 3)
    assert fds.ThermalMaterial(1, 2, (4, 5)).thermal_conductivity == (4, 5)
    with pytest.raises(ValueError):
        fds.ThermalMaterial(1, 2, 'x')


# ---- contract of the drop-in seam ----------------------------------------------------------------

def test_field_is_picklable_and_reset_keeps_signals():            # gfx.py:86, fields.py:121-127
    fld = fds.Acoustic2D(t_delta=1e-7, t_samples=5, x_delta=1e-3, x_samples=8, y_delta=1e-3,
                         y_samples=6, material=fds.AcousticMaterial(1500, 1000))
    fld.pressure.add_output(fld.get_point_region((1e-3, 1e-3)))
    fld.pressure.outputs[0].signals = [[1.0, 2.0]]
    fld.pressure.values[:] = 3
    fld.step = 4
    fld.assemble_matrices()
    clone = pickle.loads(pickle.dumps(fld))
    assert clone.step == 4 and clone.matrices_assembled
    assert np.array_equal(clone.pressure.values, fld.pressure.values)
    fld.reset()
    assert fld.step == 0 and not fld.pressure.values.any()
    assert fld.pressure.outputs[0].signals == [[1.0, 2.0]]


def test_simulate_without_gpu_fails_loudly():
    import pyfds_b200._engine as engine
    try:
        count = engine.load_library().fds_device_count()
    except RuntimeError:
        count = 0
    if count:
        pytest.skip('a GPU is present')
    fld = fds.Acoustic1D(t_delta=1e-7, t_samples=5, x_delta=1e-3, x_samples=16,
                         material=fds.AcousticMaterial(700, 0.01))
    with pytest.raises(RuntimeError):
        fld.simulate(2)


def test_subclass_sim_step_override_is_called_per_step():         # acoustic_flow.py:44-47
    calls = []

    class Counting(fds.Acoustic1D):
        def sim_step(self):
            calls.append(self.step)

    fld = Counting(t_delta=1e-7, t_samples=5, x_delta=1e-3, x_samples=16,
                   material=fds.AcousticMaterial(700, 0.01))
    fld.simulate(3)
    assert calls == [0, 1, 2] and fld.step == 3
    fld.simulate()          # falsy -> t.samples steps (fields.py:74-75)
    assert fld.step == 8


# ---- is_stable (pyfds/acoustics.py:54-63, 130-139, 227-236) ------------------------------------------

def _stability_cases():
    """(class name, constructor kwargs, extra region or None, expected) -- expected values follow from
    the reference statement ``all(c < 0.99 * min(dx, dy) / dt)``: strict inequality, 1 % headroom."""
    water = dict(sound_velocity=1500, density=1000)
    line = dict(x_samples=40, x_delta=1e-3, t_samples=10)
    grid = dict(x_samples=24, x_delta=1e-3, y_samples=20, y_delta=2e-3, t_samples=10)
    yield 'Acoustic1D', dict(line, t_delta=6e-7), water, None, True        # limit 1650
    yield 'Acoustic1D', dict(line, t_delta=6.7e-7), water, None, False     # limit 1477.6
    # exactly on the limit: `<` is strict
    limit = 0.99 * 1e-3 / 5e-7
    yield 'Acoustic1D', dict(line, t_delta=5e-7), dict(sound_velocity=limit, density=1000), None, False
    yield 'Acoustic1D', dict(line, t_delta=5e-7), \
        dict(sound_velocity=np.nextafter(limit, 0), density=1000), None, True
    # the finer of the two increments decides in 2-D (dx = 1e-3 < dy = 2e-3)
    yield 'Acoustic2D', dict(grid, t_delta=6e-7), water, None, True
    yield 'Acoustic2D', dict(grid, t_delta=6.7e-7), water, None, False
    yield 'Acoustic3DAxi', dict(grid, t_delta=6e-7), water, None, True
    yield 'Acoustic3DAxi', dict(grid, t_delta=6.7e-7), water, None, False
    # one fast inclusion is enough
    yield 'Acoustic2D', dict(grid, t_delta=6e-7), water, ('rect', 5900), False
    yield 'Acoustic2D', dict(grid, t_delta=6e-7), water, ('rect', 1600), True
    # ... unless a later region paints all of it over (the test is on the painted vector)
    yield 'Acoustic2D', dict(grid, t_delta=6e-7), water, ('hidden', 5900), True


def _stability_field(package, name, kwargs, material, extra):
    fld_ = getattr(package, name)(material=package.AcousticMaterial(**material), **kwargs)
    if extra is not None:
        kind, speed = extra
        region = fld_.get_rect_region((3e-3, 4e-3, 5e-3, 6e-3))
        fld_.add_material_region(region, package.AcousticMaterial(speed, 2000))
        if kind == 'hidden':
            fld_.add_material_region(fld_.get_rect_region((2e-3, 2e-3, 8e-3, 12e-3)),
                                     package.AcousticMaterial(1400, 900))
    return fld_


@pytest.mark.parametrize('name,kwargs,material,extra,expected', list(_stability_cases()))
def test_is_stable_known_answers(name, kwargs, material, extra, expected):
    field = _stability_field(fds, name, kwargs, material, extra)
    assert bool(field.is_stable()) is expected
    # the statement of the reference on the painted vector
    dx = min(field.x.increment, field.y.increment) if hasattr(field, 'y') else field.x.increment
    assert bool(np.all(field.material_vector('sound_velocity')
                       < 0.99 * dx / field.t.increment)) is expected


def test_is_stable_equals_the_reference_where_it_is_installed():
    import os
    import sys
    import types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    staged = os.path.join(root, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(staged, 'pyfds')):
        pytest.skip('reference not staged (baseline/_ref)')
    for mod in ('matplotlib', 'matplotlib.patches', 'matplotlib.pyplot', 'matplotlib.animation'):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    sys.path.insert(0, staged)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import pyfds
        for name, kwargs, material, extra, expected in _stability_cases():
            reference = _stability_field(pyfds, name, kwargs, material, extra)
            assert bool(reference.is_stable()) is expected, (name, kwargs, extra)
