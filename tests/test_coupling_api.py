"""The reference's coupling tests (test/test_coupling.py) restated against the drop-in, plus the flow
model. Parts that only touch host logic run on CPU; anything that steps a field needs the GPU."""

import numpy as np
import pytest

import pyfds_b200 as fds
from conftest import bits
from oracle import restate


def test_synchronized_fields_structure():                          # test_coupling.py:6-22
    acs = fds.Acoustic1D(t_delta=1, t_samples=1, x_delta=1, x_samples=3,
                         material=fds.AcousticMaterial(400, 1))
    ths = fds.Thermal1D(t_delta=1, t_samples=1, x_delta=1, x_samples=3,
                        material=fds.ThermalMaterial(1, 1, 1))
    with pytest.raises(RuntimeError):
        fds.SynchronizedFields([acs, acs], [])
    cpl = fds.SynchronizedFields([acs, ths], [])
    assert acs.velocity in vars(cpl).values()
    assert acs.pressure in vars(cpl).values()
    assert ths.temperature in vars(cpl).values()
    assert ths.heat_flux in vars(cpl).values()
    assert len(cpl.material_regions) == 2
    cpl.assemble_matrices()
    assert acs.matrices_assembled and ths.matrices_assembled


def test_boundary_coupling():                                      # test_coupling.py:30-44
    comp1 = fds.fields.FieldComponent(num_points=12)
    comp1.values = 2 * np.ones(12)
    comp2 = fds.fields.FieldComponent(num_points=12)
    comp2.values = np.ones(12)
    fds.BoundaryCoupling(comp1, comp2, lambda x: x ** 2).apply(0)
    assert np.allclose(comp2.values, 5 * np.ones(12))
    accu = fds.BoundaryCoupling(comp1, comp2, lambda x: x, False, True, 3)
    for step in (1, 2, 3, 4):
        accu.apply(step)
    assert np.allclose(comp2.values, 6 * np.ones(12))


def test_material_coupling_reassembly():                           # test_coupling.py:47-84
    base = np.array([[-1, 1, 0], [0, -1, 1], [0, 0, -1]])
    ths1 = fds.Thermal1D(t_delta=1, t_samples=1, x_delta=1, x_samples=3,
                         material=fds.ThermalMaterial(1, 1, 1))
    ths1.assemble_matrices()
    assert np.allclose(ths1.a_t_q.toarray(), base)
    comp1 = fds.fields.FieldComponent(num_points=3)
    comp1.values = np.ones(3)
    coupling = fds.MaterialCoupling(comp1, ths1, 'density', lambda x: 1 + x ** 2, stepping=2)
    assert np.allclose(ths1.material_vector('density'), 2 * np.ones(3))
    assert np.allclose(ths1.material_vector('thermal_conductivity_x'), np.ones(3))
    coupling.apply(0)
    assert np.allclose(ths1.a_t_q.toarray(), base / 2)
    comp1.values = 2 * np.ones(3)
    coupling.apply(1)                       # stepping = 2: nothing happens on odd steps
    assert np.allclose(ths1.a_t_q.toarray(), base / 2)
    coupling.apply(2)
    assert np.allclose(ths1.a_t_q.toarray(), base / 5)

    comp1.values = np.zeros(3)
    ths2 = fds.Thermal1D(t_delta=1, t_samples=1, x_delta=1, x_samples=3,
                         material=fds.ThermalMaterial(1, 1, 1))
    threshold = fds.MaterialCoupling(comp1, ths2, 'density', lambda x: 1 + x ** 2,
                                     rel_change_threshold=0.4)
    threshold.apply(0)
    assert np.allclose(ths2.a_t_q.toarray(), base)
    comp1.values = np.ones(3)
    threshold.apply(1)
    assert np.allclose(ths2.a_t_q.toarray(), base / 2)
    comp1.values = 1.5 * np.ones(3)
    threshold.apply(2)                      # below the threshold: unchanged
    assert np.allclose(ths2.a_t_q.toarray(), base / 2)


def test_dense_material_snapshot_bakes_distinct_combinations():
    from pyfds_b200 import _bake
    ths = fds.Thermal1D(t_delta=1e-3, t_samples=4, x_delta=1e-3, x_samples=12,
                        material=fds.ThermalMaterial(900, 2700, 200))
    source = fds.fields.FieldComponent(num_points=12)
    source.values = np.repeat([0.0, 1.0, 2.0], 4)
    fds.MaterialCoupling(source, ths, 'density', lambda x: 1 + x)
    ths.assemble_matrices()
    snapshot = ths._baked['snapshot']
    assert isinstance(snapshot, _bake.DenseSnapshot)
    ids, values = _bake.material_ids(snapshot, 12, 12, 0, 12)
    assert np.array_equal(values['density'][ids], ths.material_vector('density'))
    assert len(values['density']) == 4


@pytest.mark.gpu
def test_synchronized_fields_step(library):                        # test_coupling.py:23-27
    acs = fds.Acoustic1D(t_delta=1, t_samples=1, x_delta=1, x_samples=3,
                         material=fds.AcousticMaterial(400, 1))
    ths = fds.Thermal1D(t_delta=1, t_samples=1, x_delta=1, x_samples=3,
                        material=fds.ThermalMaterial(1, 1, 1))
    cpl = fds.SynchronizedFields([acs, ths], [])
    cpl.simulate(1)
    assert acs.step == 1 and ths.step == 1 and cpl.step == 1


@pytest.mark.gpu
def test_thermoacoustic_preset_matches_host_emulation(library):
    """ThermoAcoustic1D (coupled_fields.py) on the device seam vs the same coupling stepped with the
    CPU restatement of both fields."""
    def build():
        f = fds.ThermoAcoustic1D(x_samples=80, x_delta=1e-3, t_samples=60, t_delta=1e-7,
                                 thermal_material=fds.ThermalMaterial(900, 2700, 200),
                                 acoustic_material=fds.AcousticMaterial(700, 0.01, shear_viscosity=1e-3))
        k = np.arange(60)
        f.fields[0].pressure.add_boundary(f.fields[0].get_point_region(40e-3),
                                          value=np.sin(0.3 * k), additive=True)
        return f
    device = build()
    device.simulate(60)

    host = build()
    host.assemble_matrices()      # the interaction reads the acoustic field's a_v_p operator
    steppers = [restate.stepper_for(field) for field in host.fields]
    for step in range(60):
        for field, stepper in zip(host.fields, steppers):
            stepper.run(1)
            for name in stepper.components:
                getattr(field, name).values = stepper.values(name)
        for interaction in host.interactions:
            interaction.apply(step)
        # interactions rebind / modify host arrays: hand them back to the restatement
        for field, stepper in zip(host.fields, steppers):
            for name in stepper.components:
                stepper.comp[name].values = np.array(getattr(field, name).values, dtype=np.float64)
    for d_field, h_field in zip(device.fields, host.fields):
        for name in d_field._device_components:
            assert np.array_equal(bits(getattr(d_field, name).values),
                                  bits(getattr(h_field, name).values)), name
    assert device.temperature.values.any()


@pytest.mark.gpu
def test_acoustic_flow_matches_host_emulation(library):
    """AcousticFlow2D: device leapfrog + device row shift vs CPU restatement + the host row shift."""
    def build():
        f = fds.AcousticFlow2D(300.0, t_delta=1e-7, t_samples=40, x_delta=1e-3, x_samples=32,
                               y_delta=1e-3, y_samples=20, material=fds.AcousticMaterial(1500, 1000))
        rng = np.random.default_rng(3)
        for name in ('pressure', 'velocity_x', 'velocity_y'):
            getattr(f, name).values = 1e-3 * rng.standard_normal(640)
        return f
    device = build()
    assert np.all(device.flow_t_deltas == 33)
    device.simulate(40)

    host = build()
    stepper = restate.Acoustic2D(host)          # the plain leapfrog step; the shift follows below
    for step in range(40):
        stepper.run(1)
        for name in stepper.components:
            getattr(host, name).values = stepper.values(name)
        host.step = step
        host.apply_flow()
    for name in ('pressure', 'velocity_x', 'velocity_y'):
        assert np.array_equal(bits(getattr(device, name).values), bits(getattr(host, name).values))


def test_flow_periods_for_the_device():
    """|flow_t_deltas| per row; the most negative integer (zero flow) only ever divides step 0."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        f = fds.AcousticFlow2D([2000.0, -300.0, 1e4, 0.0], t_delta=1e-7, t_samples=4, x_delta=1e-3,
                               x_samples=8, y_delta=1e-3, y_samples=4,
                               material=fds.AcousticMaterial(1500, 1000))
    assert list(f.flow_t_deltas[:3]) == [5, -33, 1]
    periods = f._device_flow()
    assert periods.dtype == np.int64
    assert list(periods[:3]) == [5, 33, 1]
    assert periods[3] == np.iinfo(np.int64).max
    assert f._uses_device()

    class Custom(fds.AcousticFlow2D):
        def apply_flow(self):
            pass
    g = Custom(10.0, t_delta=1e-7, t_samples=4, x_delta=1e-3, x_samples=8, y_delta=1e-3,
               y_samples=4, material=fds.AcousticMaterial(1500, 1000))
    assert g._device_flow() is None and not g._uses_device()


def test_oracle_flow_shift_equals_reference_statement():
    """The vectorised row shift of the restatement against the literal loop of apply_flow."""
    import scenarios
    a, _ = scenarios.acoustic_flow2d(fds)
    b, _ = scenarios.acoustic_flow2d(fds)
    stepper = restate.stepper_for(a)
    assert type(stepper).__name__ == 'AcousticFlow2D'
    plain = restate.Acoustic2D(b)
    for step in range(12):
        stepper.run(1)
        plain.run(1)
        for name in plain.components:
            getattr(b, name).values = plain.values(name)
        b.step = step
        b.apply_flow()
        for name in plain.components:
            plain.comp[name].values = np.array(getattr(b, name).values)
            assert np.array_equal(bits(stepper.values(name)), bits(plain.values(name))), (step, name)


@pytest.mark.gpu
@pytest.mark.parametrize('kernel', [0, 1, 3])
def test_flow_launch_structure(library, kernel):
    """The streaming kernel keeps advancing several steps per launch in a flowing medium and ends a
    launch where a row moves; the number of shift passes equals the number of steps after which some
    row is due."""
    import scenarios
    field, steps = scenarios.acoustic_flow2d_wide(fds)
    field.device_kernel = kernel
    stepper = restate.stepper_for(field).run(steps)
    field.simulate(steps)
    for name in stepper.components:
        assert np.array_equal(bits(getattr(field, name).values), bits(stepper.values(name))), name
    engine = field.__dict__['_engine_state'].engine
    due = sum(1 for s in range(steps) if any(s % abs(int(f)) == 0 for f in field.flow_t_deltas))
    assert engine.last_flow_shifts() == due
    launches, spl, name = engine.last_launch_info()
    if kernel == 0:
        assert 'stream2d' in name and spl == 4 and launches < steps // 2


@pytest.mark.gpu
def test_flow_override_runs_on_the_host(library):
    """A subclass that overrides apply_flow is stepped once per step with its override."""
    calls = []

    class Halved(fds.AcousticFlow2D):
        def apply_flow(self):
            calls.append(self.step)
            self.pressure.values *= 0.5

    f = Halved(300.0, t_delta=1e-7, t_samples=6, x_delta=1e-3, x_samples=32, y_delta=1e-3,
               y_samples=8, material=fds.AcousticMaterial(1500, 1000))
    f.pressure.values = np.ones(256)
    g = fds.Acoustic2D(t_delta=1e-7, t_samples=6, x_delta=1e-3, x_samples=32, y_delta=1e-3,
                       y_samples=8, material=fds.AcousticMaterial(1500, 1000))
    g.pressure.values = np.ones(256)
    f.simulate(3)
    for _ in range(3):
        g.simulate(1)
        g.pressure.values *= 0.5
    assert calls == [0, 1, 2]
    assert np.array_equal(bits(f.pressure.values), bits(g.pressure.values))


# ---- device session of SynchronizedFields.simulate -------------------------------------------------

def _coupled_2d(stepping=2):
    acs = fds.Acoustic2D(t_delta=1e-7, t_samples=40, x_delta=1e-3, x_samples=64, y_delta=1e-3,
                         y_samples=24, material=fds.AcousticMaterial(1500, 1000))
    ths = fds.Thermal2D(t_delta=1e-7, t_samples=40, x_delta=1e-3, x_samples=64, y_delta=1e-3,
                        y_samples=24, material=fds.ThermalMaterial(900, 2700, 200))
    rng = np.random.default_rng(17)
    for component in (acs.pressure, acs.velocity_x, acs.velocity_y):
        component.values = 1e-3 * rng.standard_normal(64 * 24)
    ths.temperature.values = 20 + rng.standard_normal(64 * 24)
    acs.pressure.add_output(acs.get_point_region((10e-3, 5e-3)))
    ths.temperature.add_output(ths.get_line_region((3e-3, 4e-3, 9e-3, 4e-3)))
    ths.heat_flux_x.add_output(ths.get_point_region((30e-3, 12e-3)))
    heating = fds.BoundaryCoupling(acs.pressure, ths.temperature, lambda p: 1e-3 * p ** 2,
                                   additive=True, accumulate=stepping > 1, stepping=stepping)
    feedback = fds.BoundaryCoupling(ths.temperature, acs.velocity_y, lambda t: 1e-9 * t,
                                    additive=False, accumulate=False, stepping=5)
    return fds.SynchronizedFields([acs, ths], [heating, feedback])


def test_session_plan():
    cpl = _coupled_2d()
    plan = cpl._session_plan()
    assert [(s, t) for _, s, t in plan] == [((0, 0), (1, 0)), ((1, 0), (0, 2))]
    cpl.device_session = False
    assert cpl._session_plan() is None

    other = _coupled_2d()
    stray = fds.fields.FieldComponent(num_points=64 * 24)
    other.interactions.append(fds.BoundaryCoupling(stray, other.fields[0].pressure, lambda x: x))
    assert other._session_plan() is None                    # a component no field owns

    material = _coupled_2d()
    material.interactions.append(fds.MaterialCouplingPowerLaw(
        material.fields[0].pressure, material.fields[1], 'density', 2, 1e-3))
    assert material._session_plan() is None                 # re-assembles mid-run: per-step loop

    class Own(fds.SynchronizedFields):
        def sim_step(self):
            super().sim_step()
    own = _coupled_2d()
    own.__class__ = Own
    assert own._session_plan() is None


def _collect_coupled(cpl):
    import scenarios
    out = {}
    for k, field in enumerate(cpl.fields):
        for key, value in scenarios.collect(field).items():
            out['{}/{}'.format(k, key)] = np.asarray(value)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['2d', '2d_every_step', 'thermoacoustic', 'thermoacoustic_accumulating'])
def test_device_session_equals_per_step_loop(library, case):
    """simulate() with the state resident on the device against the per-step seam (full host
    coherence around every step), in two segments: values, steps and probe signals, bit for bit."""
    def build():
        if case == '2d':
            return _coupled_2d()
        if case == '2d_every_step':
            return _coupled_2d(stepping=1)
        f = fds.ThermoAcoustic1D(x_samples=96, x_delta=1e-3, t_samples=40, t_delta=1e-7,
                                 thermal_material=fds.ThermalMaterial(900, 2700, 200),
                                 acoustic_material=fds.AcousticMaterial(700, 0.01,
                                                                        shear_viscosity=1e-3),
                                 stepping=3 if case.endswith('accumulating') else 1)
        f.fields[0].pressure.add_boundary(f.fields[0].get_point_region(40e-3),
                                          value=np.sin(0.3 * np.arange(40)), additive=True)
        f.fields[1].temperature.add_output(f.fields[1].get_point_region(41e-3))
        f.fields[0].velocity.add_output(f.fields[0].get_point_region(39e-3))
        return f
    session = build()
    assert session._session_plan() is not None
    session.simulate(13)
    session.simulate(27)
    loop = build()
    loop.device_session = False
    loop.simulate(13)
    loop.simulate(27)
    assert session.step == loop.step == 40
    got, expected = _collect_coupled(session), _collect_coupled(loop)
    assert sorted(got) == sorted(expected)
    for key in expected:
        assert np.array_equal(bits(got[key]), bits(expected[key])), (case, key)
    assert any(np.any(expected[k]) for k in expected if 'temperature' in k)
