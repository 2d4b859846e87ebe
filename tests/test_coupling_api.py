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
    """AcousticFlow2D: device leapfrog + host row shift vs CPU restatement + the same row shift."""
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
    stepper = restate.stepper_for(host)
    for step in range(40):
        stepper.run(1)
        for name in stepper.components:
            getattr(host, name).values = stepper.values(name)
        host.step = step
        host.apply_flow()
    for name in ('pressure', 'velocity_x', 'velocity_y'):
        assert np.array_equal(bits(getattr(device, name).values), bits(getattr(host, name).values))
