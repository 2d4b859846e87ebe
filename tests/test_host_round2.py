"""Host-side logic added in round 2 that needs no GPU: which interactions a coupled group hands to the
device, the fingerprints that decide whether tables are re-sent, how a field is spread over devices,
the vectorised host statement of the medium flow, and the dependency lists of overlapped sweeps
(restated here: the C side builds the same lists in build_stream_plan)."""

import os

import numpy as np
import pytest

import pyfds_b200 as fds
import scenarios
from pyfds_b200 import _engine, coupling, parallel


def test_group_plan_recognises_builtin_interactions():
    kinds = {}
    for name, builder in scenarios.COUPLED_SCENARIOS.items():
        group, _ = builder(fds)
        plan = group._group_plan()
        kinds[name] = None if plan is None else [entry[0] for entry in plan]
    assert kinds['thermoacoustic1d'] == ['heating']
    assert kinds['thermoacoustic1d_stepping'] == ['heating']
    assert kinds['boundary_coupling_linear'] == ['linear', 'linear']
    assert kinds['boundary_coupling_stepping'] == ['linear', 'linear']
    assert kinds['material_coupling_exponential'] == ['law']
    assert kinds['material_coupling_powerlaw'] == ['law']
    assert kinds['material_coupling_2d'] is None          # 2-D members: host-side law, per-step seam


def test_group_plan_leaves_opaque_transfer_functions_to_the_host():
    sound, heat = scenarios._two_line_fields(fds, 20)
    opaque = fds.BoundaryCoupling(heat.temperature, sound.pressure, lambda values: 1e-4 * values)
    group = fds.SynchronizedFields([sound, heat], [opaque])
    assert group._group_plan() is None
    assert group._session_plan() is not None              # still device-resident members
    # a subclass with its own transfer function is opaque as well
    class Damped(fds.MaterialCouplingExponential):
        def transfer_function(self, values):
            return 0.5 * super().transfer_function(values)
    sound2, heat2 = scenarios._two_line_fields(fds, 20)
    law = Damped(heat2.temperature, sound2, 'sound_velocity', a=0.8, b=-0.01)
    assert fds.SynchronizedFields([sound2, heat2], [law])._group_plan() is None
    group.device_session = False
    assert group._group_plan() is None and group._session_plan() is None


def test_linear_transfer_function_is_the_lambda_it_replaces():
    values = np.linspace(-2, 3, 11)
    assert np.array_equal(coupling.linear(2.5)(values), 2.5 * values)
    target = fds.fields.FieldComponent(num_points=11)
    source = fds.fields.FieldComponent(num_points=11)
    source.values = values.copy()
    fds.BoundaryCoupling(source, target, coupling.linear(-0.5)).apply(0)
    assert np.array_equal(target.values, -0.5 * values)


def test_tables_are_resent_on_content_not_identity():
    a = np.arange(10, dtype=np.int64)
    b = np.linspace(0, 1, 10)
    assert _engine._same_arrays((a, b), (a.copy(), b.copy()))
    assert not _engine._same_arrays(None, (a, b))
    assert not _engine._same_arrays((a, b), (a, b + 1e-16 * (np.arange(10) == 3)))
    assert not _engine._same_arrays((a, b), (a.astype(np.int32), b))
    assert not _engine._same_arrays((a, b), (a[:9], b))
    assert not _engine._same_arrays((np.zeros(2),), (-np.zeros(2),))      # -0.0 is not 0.0 here
    assert _engine._same_arrays((np.zeros(0),), (np.zeros(0),))
    mark = _engine._fingerprint(a, b)
    assert mark == _engine._fingerprint(a.copy(), b.copy())
    assert mark != _engine._fingerprint(a.astype(np.int32), b)


def test_slab_devices_selection(monkeypatch):
    field, _ = scenarios._acoustic2d(fds, lossy=False, nx=128, ny=64, steps=4)
    monkeypatch.delenv('FDS_DEVICES', raising=False)
    assert _engine._slab_devices(field) is None
    field.devices = [0, 1]
    assert _engine._slab_devices(field) == (0, 1)
    field.devices = 4
    assert _engine._slab_devices(field) == (0, 1, 2, 3)
    field.devices = 16                                   # 4 rows per slab: thinner than the edge bands
    assert _engine._slab_devices(field) is None
    field.devices = [3]
    assert _engine._slab_devices(field) is None
    del field.devices
    monkeypatch.setenv('FDS_DEVICES', '2')
    assert _engine._slab_devices(field) == (0, 1)
    line = scenarios.acoustic1d_lossy(fds)[0]
    assert _engine._slab_devices(line) is None            # 1-D problems stay on one GPU


def test_flow_host_statement_moves_due_rows_only():
    """pyfds/acoustic_flow.py:49-57: rows whose period divides the step move one cell towards +x."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        field = fds.AcousticFlow2D(scenarios._flow_for_periods((3, 1, -5, 1000, 2, 7), 6),
                                   t_delta=1e-7, t_samples=40, x_delta=1e-3, x_samples=16,
                                   y_delta=1e-3, y_samples=6, material=fds.AcousticMaterial(1500, 1000))
    periods = [int(f) for f in field.flow_t_deltas]
    assert periods == [3, 1, -5, 1000, 2, 7]
    rng = np.random.default_rng(3)
    for step in (0, 1, 2, 3, 5, 6, 10, 14, 15):
        before = {}
        for name in ('pressure', 'velocity_x', 'velocity_y'):
            getattr(field, name).values = rng.standard_normal(96)
            before[name] = getattr(field, name).values.reshape(6, 16).copy()
        field.step = step
        field.apply_flow()
        for name, old in before.items():
            new = getattr(field, name).values.reshape(6, 16)
            for row, period in enumerate(periods):
                if step % period == 0:
                    assert np.array_equal(new[row, 1:], old[row, :-1]) and new[row, 0] == 0
                else:
                    assert np.array_equal(new[row], old[row])


def _dependencies(tasks, n_strips, lag):
    """What build_stream_plan computes (fds_abi.cu): the tasks of the previous sweep a task must wait
    for -- strips s-1, s, s+1 (cyclic) whose rows come within lag + 1 of its own."""
    deps = []
    for s, ys, ye in tasks:
        near = {(s - 1) % n_strips, s, (s + 1) % n_strips}
        deps.append([k for k, (s2, ys2, ye2) in enumerate(tasks)
                     if s2 in near and ys2 < ye + lag + 1 and ye2 > ys - lag - 1])
    return deps


def test_sweep_dependencies_cover_read_and_overwrite_hazards():
    """For a task table as the plan builds it (strips x chunks), a task's dependency list must contain
    every task whose OUTPUT it reads (rows ys - lag .. ye + lag, strip +- 4 cells, one more row for the
    flat-index wrap) and every task that READS what it overwrites -- the same set."""
    n_strips, rows, height, lag = 7, 100, 23, 4
    tasks = [(s, ys, min(ys + height, rows)) for ys in range(0, rows, height) for s in range(n_strips)]
    deps = _dependencies(tasks, n_strips, lag)
    width = 56
    for k, (s, ys, ye) in enumerate(tasks):
        read = (ys - lag - 1, ye + lag + 1, s * width - 4, s * width + width + 4)
        for k2, (s2, ys2, ye2) in enumerate(tasks):
            # column ranges are cyclic over n_strips * width cells (rows wrap into each other)
            total = n_strips * width
            cols = {c % total for c in range(read[2], read[3])}
            owned = set(range(s2 * width, s2 * width + width))
            overlaps = bool(cols & owned) and ys2 < read[1] and ye2 > read[0]
            if overlaps:
                assert k2 in deps[k], (tasks[k], tasks[k2])
            # symmetric: if k2's read footprint covers what k writes, k must wait for k2 as well
            read2 = (ys2 - lag - 1, ye2 + lag + 1)
            cols2 = {c % total for c in range(s2 * width - 4, s2 * width + width + 4)}
            writes = set(range(s * width, s * width + width))
            if bool(cols2 & writes) and ys < read2[1] and ye > read2[0]:
                assert k2 in deps[k], (tasks[k], tasks[k2])
    assert max(len(d) for d in deps) <= 9


def test_partition_and_halo_for_local_slabs():
    field, _ = scenarios._acoustic2d(fds, lossy=False, nx=256, ny=300, steps=4)
    assert parallel.halo_rows_for(field, 8) == parallel.STREAM_STEPS
    parts = parallel.partition_rows(300, 8)
    assert min(rows for _, rows in parts) >= _engine.MIN_SLAB_ROWS
    slabs = parallel.LocalSlabs(field, range(8))
    assert slabs.devices == tuple(range(8)) and slabs.engines == []
