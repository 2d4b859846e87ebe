"""Host logic of the drop-in that the reference leaves untested but the device path relies on:
region index maps against goldens from the real reference (values AND order, bit exact), the baking of
materials, boundaries and probes into device tables."""

import os

import numpy as np
import pytest

import pyfds_b200 as fds
import scenarios
from pyfds_b200 import _bake

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_region_index_maps_match_reference():
    gold = np.load(os.path.join(GOLDEN, 'regions.npz'))
    cases = scenarios.region_cases(fds)
    assert sorted(cases) == sorted(gold.files)
    for name, region in cases.items():
        got = np.asarray(region.indices)
        assert got.dtype.kind == 'i', name
        assert np.array_equal(got, gold[name]), name
        assert len(region) == gold[name].shape[0], name


def test_zero_length_line_fails_like_reference():
    f = fds.fields.Field2D(10, 1e-3, 10, 1e-3, 1, 1.0, int(5))
    with pytest.raises(ValueError):
        f.get_line_region((2e-3, 2e-3, 2e-3, 2e-3))


def test_material_vector_matches_painting_order():
    fld = fds.Acoustic2D(t_delta=1e-7, t_samples=1, x_delta=1e-3, x_samples=12, y_delta=1e-3,
                         y_samples=9, material=fds.AcousticMaterial(1500, 1000))
    fld.add_material_region(fld.get_rect_region((2e-3, 1e-3, 5e-3, 4e-3)),
                            fds.AcousticMaterial(1200, 900))
    fld.add_material_region(fld.get_rect_region((4e-3, 3e-3, 6e-3, 5e-3)),
                            fds.AcousticMaterial(1000, 800))
    # explicit painting with index lists, reference style
    expected = np.zeros(fld.num_points)
    for mat_reg in fld.material_regions:
        expected[np.asarray(mat_reg.region.indices)] = mat_reg.materials[0].density
    assert np.array_equal(fld.material_vector('density'), expected)

    fld.assemble_matrices()
    ids, values = _bake.material_ids(fld._baked['snapshot'], fld.num_points, 12, 0, fld.num_points)
    assert np.array_equal(values['density'][ids], expected)
    assert ids.min() == 1 and len(values['density']) == 4
    # a slab window with halo rows: out-of-grid cells are void (id 0)
    ids_w, _ = _bake.material_ids(fld._baked['snapshot'], fld.num_points, 12, -24, 5 * 12)
    assert not ids_w[:24].any() and np.array_equal(ids_w[24:], ids[:60])


def test_material_ids_when_a_region_defines_only_some_parameters():
    class OnlyDensity:
        density = 7.0

    fld = fds.Acoustic2D(t_delta=1e-7, t_samples=1, x_delta=1e-3, x_samples=6, y_delta=1e-3,
                         y_samples=5, material=fds.AcousticMaterial(1500, 1000))
    fld.add_material_region(fld.get_rect_region((1e-3, 1e-3, 2e-3, 2e-3)), OnlyDensity())
    fld.assemble_matrices()
    ids, values = _bake.material_ids(fld._baked['snapshot'], 30, 6, 0, 30)
    for name in ('density', 'sound_velocity', 'absorption_coef'):
        assert np.array_equal(values[name][ids], fld.material_vector(name)), name


def test_coefficient_tables_equal_reference_factor_vectors():
    """Per-material evaluation must give the very bits the per-cell reference expression gives."""
    fld, _ = scenarios.acoustic3daxi_lossy(fds)
    fld.assemble_matrices()
    nx = fld.x.samples
    ids, values = _bake.material_ids(fld._baked['snapshot'], fld.num_points, nx, 0, fld.num_points)
    tables = fld._coefficient_tables(values)
    col = np.arange(fld.num_points) % nx
    a_p_vx = fld.a_p_vx        # scipy operator built from per-cell vectors, reference style
    fx = np.vstack((np.zeros((1, nx)), tables['column_tables']['FX']))[ids, col]
    assert np.array_equal(a_p_vx.data[1], fx)
    a_vv = fld.a_vx_vx
    offsets = list(a_vv.offsets)
    vm1 = np.vstack((np.zeros((1, nx)), tables['column_tables']['VM1']))[ids, col]
    vp1 = np.vstack((np.zeros((1, nx)), tables['column_tables']['VP1']))[ids, col]
    v0 = np.concatenate(([0.0], tables['tables']['V0']))[ids]
    assert np.array_equal(a_vv.data[offsets.index(-1)], vm1)
    assert np.array_equal(a_vv.data[offsets.index(1)], vp1)
    assert np.array_equal(a_vv.data[offsets.index(0)], v0)
    assert tables['lossy']


def test_boundary_table_order_duplicates_and_windows():
    fld, steps = scenarios.acoustic2d_boundaries(fds)
    nx = fld.x.samples
    signals = []
    table = _bake.boundary_table(fld.pressure.boundaries, 10, 20, 0, fld.num_points, signals)
    assert np.all(np.diff(table.cells) > 0)
    assert table.offsets[0] == 0 and table.offsets[-1] == table.alpha.shape[0]
    # cell 0 is hit by 'left' (additive scalar) and then by 'first' (signal, not additive)
    k = np.searchsorted(table.cells, 0)
    ops = slice(table.offsets[k], table.offsets[k + 1])
    assert list(table.alpha[ops]) == [1.0, 0.0]
    assert table.signal[ops][0] == -1 and table.signal[ops][1] >= 0
    assert table.value[ops][0] == 2.5e-4
    # signal windows start at first_step
    full = fld.pressure.boundaries[2].value
    assert np.array_equal(signals[0], full[10:30])
    # too short a signal fails before anything is launched (regions.py:141)
    with pytest.raises(IndexError):
        _bake.boundary_table(fld.pressure.boundaries, steps - 5, 20, 0, fld.num_points, [])
    # duplicated index in one region: only the last occurrence survives, with its own signal
    sig = []
    tvy = _bake.boundary_table(fld.velocity_y.boundaries, 0, steps, 0, fld.num_points, sig)
    assert list(tvy.cells) == [5 + 4 * nx, 6 + 4 * nx, 7 + 4 * nx]
    assert list(tvy.signal) == [0, 3, 2]
    # a slab window keeps only its cells and shifts them
    part = _bake.boundary_table(fld.velocity_x.boundaries, 0, steps, 10 * nx, 20 * nx, [])
    assert part.cells.min() >= 0 and part.cells.max() < 10 * nx


def test_equal_region_signals_share_one_window():
    """Boundaries driven by bit-equal signals get ONE signal index (the device applies single-signal
    cells through at most 7 classes per component); per-point signals keep one index per point."""
    fld, steps = scenarios.acoustic2d_signal_lines(fds)
    signals = []
    table = _bake.boundary_table(fld.pressure.boundaries, 0, steps, 0, fld.num_points, signals)
    # three boundaries on pressure: two lines with the same burst (one of them a copy), one other pulse
    assert len(signals) == 2
    assert sorted(set(table.signal.tolist())) == [0, 1]
    nx = fld.x.samples
    first = table.signal[table.offsets[np.searchsorted(table.cells, 5 * nx)]]          # (0, 5)
    second = table.signal[table.offsets[np.searchsorted(table.cells, 60)]]             # (60, 0)
    assert first == second
    # a different window of the same signals is still shared, and a second component appends
    table_vy = _bake.boundary_table(fld.velocity_y.boundaries, 0, steps, 0, fld.num_points, signals)
    assert len(signals) == 2 + 8 and len(set(table_vy.signal.tolist())) == 8


def test_probe_table_slots():
    fld, _ = scenarios.acoustic2d_boundaries(fds)
    cells, slots, nxt = _bake.probe_table(fld.pressure.outputs, 0, 0, fld.num_points)
    counts = [len(np.asarray(o.region.indices)) for o in fld.pressure.outputs]
    assert nxt == sum(counts)
    assert np.all(np.diff(cells) >= 0)
    order = np.argsort(slots)
    flat = np.concatenate([np.asarray(o.region.indices) for o in fld.pressure.outputs])
    assert np.array_equal(cells[order], flat)
