"""Baking: turns the host-side description of a field (material regions, boundaries, outputs) into
the compact tables the step kernels read.

* Materials (``Field.material_vector`` + ``assemble_matrices``, ``pyfds/fields.py:34-57``,
  ``pyfds/acoustics.py:89-109``): regions are piecewise constant, so instead of the reference's
  per-cell fp64 factor vectors (4-5 operators x 2-5 diagonals x N doubles) the device gets one
  *material id* byte per cell and a table of coefficients per id. The coefficients are evaluated by the
  model classes with the reference's own NumPy expressions on a vector with one entry per material,
  which gives bit-identical values.
* Boundaries (``FieldComponent.apply_bounds`` / ``Boundary.apply``, ``pyfds/fields.py:591-600``,
  ``pyfds/regions.py:125-145``): an ordered operation table per component, grouped by cell.
* Outputs (``FieldComponent.write_outputs``, ``pyfds/fields.py:602-611``): a probe-point table per
  component; probe k of the whole field is column k of every step's probe record.
"""

import numpy as np

from . import regions as reg

MAX_MATERIALS = 31


class MaterialSnapshot:
    """What ``assemble_matrices`` would have read from the material regions, frozen at that moment:
    per region (in list order) the value of every requested parameter its materials define."""

    def __init__(self, field, params):
        self.params = tuple(params)
        self.entries = []
        found = set()
        for mat_reg in field.material_regions:
            values = {}
            for mat in mat_reg.materials:
                for name in self.params:
                    if hasattr(mat, name):
                        value = getattr(mat, name)
                        if np.ndim(value) != 0:
                            raise NotImplementedError(
                                'Material parameter {} is not a scalar; per-cell material '
                                'parameters are not supported by the device engine.'.format(name))
                        values[name] = value
                        found.add(name)
            self.entries.append((mat_reg.region, values))
        for name in self.params:
            if name not in found:
                # same failure as Field.material_vector (pyfds/fields.py:54-55)
                raise KeyError('Material parameter {} not found in set materials.'.format(name))

    def uniform(self):
        """True if every region defines every parameter (the normal case): one owner map serves all
        parameters."""
        return all(len(values) == len(self.params) for _, values in self.entries)


class DenseSnapshot:
    """Material description of a field whose ``material_vector`` was replaced on the instance (the
    reference's ``MaterialCoupling`` does that, ``pyfds/coupling.py:179-180``): the per-point vectors
    themselves, frozen at ``assemble_matrices`` time. Distinct value combinations become materials."""

    def __init__(self, field, params):
        self.params = tuple(params)
        self.vectors = {p: np.asarray(field.material_vector(p), dtype=np.float64).reshape(-1)
                        for p in self.params}


def distinct_combinations(snapshot):
    """Number of distinct per-point parameter combinations of a ``DenseSnapshot``."""
    stacked = np.stack([snapshot.vectors[p] for p in snapshot.params], axis=1)
    return len(np.unique(stacked, axis=0))


def _dense_material_ids(snapshot, num_points, cell_lo, cell_hi):
    lo, hi = max(cell_lo, 0), min(cell_hi, num_points)
    ids = np.zeros(cell_hi - cell_lo, dtype=np.uint8)
    stacked = np.stack([snapshot.vectors[p][lo:hi] for p in snapshot.params], axis=1)
    combos, inverse = np.unique(stacked, axis=0, return_inverse=True)
    if len(combos) > MAX_MATERIALS:
        raise NotImplementedError(
            'The per-point material vectors hold {} distinct parameter combinations; the device '
            'engine supports {}.'.format(len(combos), MAX_MATERIALS))
    ids[lo - cell_lo:hi - cell_lo] = (inverse.reshape(-1) + 1).astype(np.uint8)
    values = {p: np.concatenate(([0.0], combos[:, k])) for k, p in enumerate(snapshot.params)}
    return ids, values


def _local_cells(region, nx, cell_lo, cell_hi):
    """Flat indices of ``region`` that fall into [cell_lo, cell_hi), shifted to start at 0."""
    idx = region.index_array() if isinstance(region, reg.Region) else \
        np.asarray(region.indices, dtype=np.int64).reshape(-1)
    keep = (idx >= cell_lo) & (idx < cell_hi)
    return idx[keep] - cell_lo


def _paint_local(target, region, value, nx, cell_lo, cell_hi):
    """``target[region.indices] = value`` restricted to the cell window [cell_lo, cell_hi)."""
    d = region.descriptor if isinstance(region, reg.Region) else None
    if d is not None and d[0] == 'rect' and d[5] == nx and cell_lo % nx == 0 and cell_hi % nx == 0:
        _, x0, x1, y0, y1, _ = d
        r_lo, r_hi = cell_lo // nx, cell_hi // nx
        ya, yb = max(y0, r_lo), min(y1 + 1, r_hi)
        if ya < yb:
            target.reshape(-1, nx)[ya - r_lo:yb - r_lo, x0:x1 + 1] = value
    elif d is not None and d[0] == 'range':
        a, b = max(d[1], cell_lo), min(d[2], cell_hi)
        if a < b:
            target[a - cell_lo:b - cell_lo] = value
    else:
        target[_local_cells(region, nx, cell_lo, cell_hi)] = value


def material_ids(snapshot, num_points, nx, cell_lo, cell_hi):
    """Material id (1..n) of every cell in the global window [cell_lo, cell_hi) -- cells outside
    [0, num_points) get id 0 (void) -- and, per parameter, the value of each id.

    Returns ``(ids uint8[cell_hi - cell_lo], {param: float64[n + 1]})`` with entry 0 of every value
    vector unused (the void material has no physical parameters; its coefficients are all zero).
    """
    if isinstance(snapshot, DenseSnapshot):
        return _dense_material_ids(snapshot, num_points, cell_lo, cell_hi)
    lo, hi = max(cell_lo, 0), min(cell_hi, num_points)
    ids = np.zeros(cell_hi - cell_lo, dtype=np.uint8)
    inner = ids[lo - cell_lo:hi - cell_lo]

    if snapshot.uniform():
        # distinct parameter tuples -> ids; regions with equal materials share one id
        tuples = []
        region_id = []
        for _, values in snapshot.entries:
            key = tuple(float(values[p]) for p in snapshot.params)
            if key not in tuples:
                tuples.append(key)
            region_id.append(tuples.index(key) + 1)
        if len(tuples) > MAX_MATERIALS:
            raise NotImplementedError('More than {} distinct materials.'.format(MAX_MATERIALS))
        for (region, _), rid in zip(snapshot.entries, region_id):
            _paint_local(inner, region, rid, nx, lo, hi)
        values = {p: np.array([0.0] + [t[k] for t in tuples], dtype=np.float64)
                  for k, p in enumerate(snapshot.params)}
        # cells no region covers keep id 0 in the reference too (np.zeros), but there their parameters
        # are 0 and the coefficients become inf/nan; refuse instead of silently diverging
        if (inner == 0).any():
            raise ValueError('Some cells are not covered by any material region.')
        return ids, values

    # general case: a region may define only some parameters, so each parameter has its own owner map
    owners = []
    for p in snapshot.params:
        owner = np.zeros(hi - lo, dtype=np.int32)
        for k, (region, values) in enumerate(snapshot.entries):
            if p in values:
                _paint_local(owner, region, k + 1, nx, lo, hi)
        owners.append(owner)
    stacked = np.stack(owners, axis=1)
    combos, inverse = np.unique(stacked, axis=0, return_inverse=True)
    if len(combos) > MAX_MATERIALS:
        raise NotImplementedError('More than {} distinct materials.'.format(MAX_MATERIALS))
    inner[:] = (inverse.reshape(-1) + 1).astype(np.uint8)
    values = {}
    for k, p in enumerate(snapshot.params):
        column = [0.0]
        for combo in combos:
            # owner 0 = no region painted this parameter: the reference leaves it at 0.0
            column.append(float(snapshot.entries[combo[k] - 1][1][p]) if combo[k] else 0.0)
        values[p] = np.array(column, dtype=np.float64)
    return ids, values


class BoundaryTable:
    """Boundary operations of one component in the CSR form of ``fds_upload_boundaries``."""

    def __init__(self, cells, offsets, alpha, value, signal):
        self.cells = cells
        self.offsets = offsets
        self.alpha = alpha
        self.value = value
        self.signal = signal


def _window(signal, first_step, n_steps):
    """Samples first_step .. first_step+n_steps-1 of a signal as float64; ``IndexError`` if it is too
    short (the reference fails at ``value[step]``, ``pyfds/regions.py:141,144``)."""
    signal = np.asarray(signal)
    if signal.ndim != 1:
        raise NotImplementedError('Boundary signals must be one-dimensional.')
    if first_step + n_steps > signal.shape[0]:
        raise IndexError('index {} is out of bounds for axis 0 with size {}'.format(
            signal.shape[0], signal.shape[0]))
    return signal[first_step:first_step + n_steps].astype(np.float64)


def boundary_table(boundaries, first_step, n_steps, cell_lo, cell_hi, signals):
    """Bakes the ``boundaries`` list of one component for the global cell window
    [cell_lo, cell_hi). New signal windows are appended to the list ``signals``."""
    cells, order, alpha, value, signal = [], [], [], [], []
    for position, bound in enumerate(boundaries):
        idx = np.asarray(bound.region.indices, dtype=np.int64).reshape(-1)
        kind = bound.kind()
        n = idx.shape[0]
        if kind == 'scalar':
            op_value = np.full(n, float(bound.value))
            op_signal = np.full(n, -1, dtype=np.int32)
        elif kind == 'signal' and np.ndim(bound.value) == 1:
            # one signal for the whole region; boundaries driven by equal signals share one window
            # (the device applies such sources through a handful of classes, see fds_upload_boundaries)
            window = _window(bound.value, first_step, n_steps)
            shared = next((k for k, other in enumerate(signals)
                           if other.shape == window.shape and
                           np.array_equal(other.view(np.int64), window.view(np.int64))), None)
            if shared is None:
                signals.append(window)
                shared = len(signals) - 1
            op_value = np.zeros(n)
            op_signal = np.full(n, shared, dtype=np.int32)
        else:
            # one signal per point: a list of signals, or a 2-D array [step][point]
            per_point = list(np.asarray(bound.value).T) if kind == 'signal' else list(bound.value)
            if len(per_point) != n:
                raise ValueError('shape mismatch: {} signals for {} boundary points'.format(
                    len(per_point), n))
            base = len(signals)
            op_value = np.zeros(n)
            op_signal = np.arange(base, base + n, dtype=np.int32)
            keep_local = (idx >= cell_lo) & (idx < cell_hi)
            for k, sig in enumerate(per_point):
                # only windows of points inside the slab are needed, but ids stay dense
                signals.append(_window(sig, first_step, n_steps) if keep_local[k]
                               else np.zeros(n_steps))
        # duplicate indices inside one region: NumPy fancy assignment keeps the last write
        if n > 1:
            _, last = np.unique(idx[::-1], return_index=True)
            keep = np.sort(n - 1 - last)
        else:
            keep = np.arange(n)
        keep = keep[(idx[keep] >= cell_lo) & (idx[keep] < cell_hi)]
        cells.append(idx[keep] - cell_lo)
        order.append(np.full(keep.shape[0], position, dtype=np.int64))
        alpha.append(np.full(keep.shape[0], float(bound.additive)))
        value.append(op_value[keep])
        signal.append(op_signal[keep])

    if not cells or sum(c.shape[0] for c in cells) == 0:
        return BoundaryTable(np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0), np.zeros(0),
                             np.zeros(0, np.int32))
    cells = np.concatenate(cells)
    order = np.concatenate(order)
    perm = np.lexsort((order, cells))       # by cell, then list position
    cells = cells[perm]
    unique_cells, start = np.unique(cells, return_index=True)
    offsets = np.append(start, cells.shape[0]).astype(np.int32)
    return BoundaryTable(unique_cells.astype(np.int64), offsets,
                         np.concatenate(alpha)[perm].astype(np.float64),
                         np.concatenate(value)[perm].astype(np.float64),
                         np.concatenate(signal)[perm].astype(np.int32))


def probe_table(outputs, slot_base, cell_lo, cell_hi):
    """Probe points of one component inside the global cell window. Slot numbers count the points of
    all outputs in list order starting at ``slot_base`` (whether or not they fall into the window).

    Returns ``(cells int64 ascending, slots int32, next_slot_base)``."""
    cells, slots = [], []
    base = slot_base
    for output in outputs:
        idx = np.asarray(output.region.indices, dtype=np.int64).reshape(-1)
        keep = (idx >= cell_lo) & (idx < cell_hi)
        cells.append(idx[keep] - cell_lo)
        slots.append((base + np.arange(idx.shape[0], dtype=np.int64))[keep])
        base += idx.shape[0]
    if cells:
        cells = np.concatenate(cells)
        slots = np.concatenate(slots)
        perm = np.argsort(cells, kind='stable')
        return cells[perm].astype(np.int64), slots[perm].astype(np.int32), base
    return np.zeros(0, np.int64), np.zeros(0, np.int32), base
