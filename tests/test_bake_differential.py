"""The tables the device reads (``pyfds_b200/_bake.py``: material id map + per-material values,
boundary operations in CSR form with signal windows, probe slots) EMULATED on the CPU for seeded
random scenarios and compared with the statement they were baked from -- the restatement's
``apply_bounds`` / ``write_outputs`` (pyfds/fields.py:591-611, pyfds/regions.py:136-145) and the
painted material vectors (pyfds/fields.py:34-57) -- bit for bit, on the whole grid and on random
slab windows (what a rank of a multi-GPU run bakes)."""

import numpy as np
import pytest

import pyfds_b200 as fds
from conftest import bits
from oracle import restate
from pyfds_b200 import _bake
from test_oracle_differential import MODELS, random_scenario

CASES = [(klass, lossy) for klass in sorted(MODELS)
         for lossy in ((False, True) if MODELS[klass][1] == 'acoustic' else (False,))]


def windows(field, rng):
    """The whole grid, and for 2-D fields two random row windows (slabs)."""
    n = field.num_points
    yield 0, n
    if hasattr(field, 'y'):
        nx, ny = field.x.samples, field.y.samples
        for _ in range(2):
            a, b = sorted(int(k) for k in rng.choice(ny + 1, 2, replace=False))
            yield a * nx, b * nx


def apply_table(table, signals, values, step_in_window):
    """What the kernels' slow path does with a boundary table: the operations of a cell in order,
    v = alpha * v + (sample of its signal | value)."""
    for k, cell in enumerate(table.cells):
        v = values[cell]
        for o in range(table.offsets[k], table.offsets[k + 1]):
            sample = signals[table.signal[o]][step_in_window] if table.signal[o] >= 0 \
                else table.value[o]
            v = table.alpha[o] * v + sample
        values[cell] = v


@pytest.mark.parametrize('klass,lossy', CASES)
@pytest.mark.parametrize('seed', range(3))
def test_baked_boundaries_and_probes_emulated_equal_the_restatement(klass, lossy, seed):
    field, steps = random_scenario(fds, klass, 8000 + 10 * seed + lossy, lossy)
    rng = np.random.default_rng(8500 + seed)
    first = int(rng.integers(0, steps // 2))
    n_steps = steps - first
    for name in MODELS[klass][2]:
        component = getattr(field, name)
        expected = restate._Component(component)      # the statement: copy of the host component
        for lo, hi in windows(field, rng):
            signals = [np.full(n_steps, np.nan)]      # tables of other components come first
            table = _bake.boundary_table(component.boundaries, first, n_steps, lo, hi, signals)
            cells, slots, next_slot = _bake.probe_table(component.outputs, 5, lo, hi)
            assert next_slot == 5 + sum(len(np.asarray(o.region.indices).reshape(-1))
                                        for o in component.outputs)
            assert np.all(np.diff(table.cells) > 0) and np.all(np.diff(cells) >= 0)
            if table.cells.size:
                assert table.cells.min() >= 0 and table.cells.max() < hi - lo
            state = np.random.default_rng(8600 + seed).standard_normal(field.num_points)
            for step in range(first, first + n_steps):
                expected.values = state.copy()
                expected.apply_bounds(step)
                local = state[lo:hi].copy()
                apply_table(table, signals, local, step - first)
                assert np.array_equal(bits(local), bits(expected.values[lo:hi])), (name, lo, hi, step)
                # probes: slot k of the component = point k of its outputs in list order
                record = np.full(next_slot, np.nan)
                record[slots] = local[cells]
                flat = np.concatenate([np.asarray(o.region.indices, dtype=np.int64).reshape(-1)
                                       for o in component.outputs] or [np.zeros(0, np.int64)])
                inside = (flat >= lo) & (flat < hi)
                want = np.full(next_slot, np.nan)
                want[5:][inside] = expected.values[flat[inside]]
                assert np.array_equal(bits(record), bits(want)), (name, lo, hi, step)
                state = 0.5 * expected.values + 0.125


@pytest.mark.parametrize('klass,lossy', CASES)
@pytest.mark.parametrize('seed', range(3))
def test_baked_material_map_equals_the_painted_vectors(klass, lossy, seed):
    field, _ = random_scenario(fds, klass, 8000 + 10 * seed + lossy, lossy)
    rng = np.random.default_rng(8700 + seed)
    field.assemble_matrices()
    nx = field.x.samples
    kind = MODELS[klass][1]
    parameters = ('sound_velocity', 'density', 'absorption_coef') if kind == 'acoustic' else \
        ('density', 'heat_capacity', 'thermal_conductivity_x') + \
        (('thermal_conductivity_y',) if hasattr(field, 'y') else ())
    for lo, hi in windows(field, rng):
        pad = nx if hasattr(field, 'y') else 3        # a halo beyond the grid is void (id 0)
        ids, values = _bake.material_ids(field._baked['snapshot'], field.num_points, nx, lo - pad,
                                         hi + pad)
        inside = np.arange(lo - pad, hi + pad)
        valid = (inside >= 0) & (inside < field.num_points)
        assert not ids[~valid].any() and ids[valid].min() >= 1
        for name in parameters:
            painted = field.material_vector(name)
            assert values[name][0] == 0.0
            assert np.array_equal(bits(values[name][ids][valid]), bits(painted[inside[valid]])), name
