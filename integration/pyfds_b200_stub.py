"""What a pyfds maintainer would add to the reference as ``pyfds/_b200.py`` (option B):
a ctypes binding of ``libfdsb200.so`` (C ABI: ``include/fdsb200.h``) that runs ``Field.simulate`` of the
REFERENCE's own ``Acoustic1D`` / ``Acoustic2D`` / ``Thermal2D`` objects on the GPU.

Nothing of ``pyfds_b200``'s Python layer is used: the stub reads the reference objects through their
public attributes (``material_vector``, ``boundaries``, ``outputs``, ``values``, ``step``), evaluates
the coefficient expressions of the reference's ``assemble_matrices`` on one entry per distinct material,
and talks to the library through the entry points a C / cgo / JNI binding would use as well.

    import pyfds, pyfds_b200_stub
    pyfds_b200_stub.install(pyfds, '/path/to/libfdsb200.so')     # patches Field.simulate
    field.simulate(1000)                                         # now one fds_simulate call

``tests/test_reference_stub.py`` drives the unmodified reference through it and compares with goldens
the same reference produced on its scipy path.
"""

import ctypes as ct

import numpy as np

_lib = None


class fds_desc(ct.Structure):          # struct fds_desc, include/fdsb200.h
    _fields_ = [('model', ct.c_int32), ('device', ct.c_int32), ('nx', ct.c_int64), ('ny', ct.c_int64),
                ('row0', ct.c_int64), ('rows', ct.c_int64), ('halo_rows', ct.c_int32),
                ('lossy', ct.c_int32), ('n_materials', ct.c_int32), ('kernel', ct.c_int32)]


# enum fds_table
GX, GY, FX, FY, VM1, VP1, VMN, VPN, V0 = range(9)


def _ptr(array):
    return array.ctypes.data_as(ct.c_void_p)


def _check(ctx, rc):
    if rc != 0:
        raise RuntimeError(_lib.fds_last_error(ctx).decode())


# ---- per model: components, material parameters, coefficient tables ---------------------------------
# The expressions are those of the reference's assemble_matrices, evaluated by NumPy on a vector with one
# entry per distinct material -- identical bits to the per-cell factor vectors the reference builds.

def _acoustic1d(f, c, rho, mu):
    dt, dx = f.t.increment, f.x.increment                         # pyfds/acoustics.py:27-38
    h = dt / dx ** 2 * mu / rho
    return {FX: dt / dx * c ** 2 * rho, GX: dt / dx / rho, VM1: h, V0: -2 * h, VP1: h}, np.any(mu != 0)


def _acoustic2d(f, c, rho, mu):
    dt, dx, dy = f.t.increment, f.x.increment, f.y.increment      # pyfds/acoustics.py:89-109
    hx, hy = dt / dx ** 2 * mu / rho, dt / dy ** 2 * mu / rho
    zero = np.zeros_like(hx)                                      # (d_x2 + d_y2).todia(): 0 + a + b
    return {GX: dt / dx / rho, GY: dt / dy / rho, FX: dt / dx * c ** 2 * rho,
            FY: dt / dy * c ** 2 * rho, VM1: zero + hx, VP1: zero + hx, VMN: zero + hy,
            VPN: zero + hy, V0: (zero + -2 * hx) + -2 * hy}, np.any(mu != 0)


def _thermal2d(f, rho, cp, kx, ky):
    dt, dx, dy = f.t.increment, f.x.increment, f.y.increment      # pyfds/thermal.py:75-90
    return {FX: dt / dx / rho / cp, FY: dt / dy / rho / cp, GX: 1 / dx * kx, GY: 1 / dy * ky}, False


MODELS = {   # class name -> (model id, components, material parameters, coefficient function)
    'Acoustic1D': (1, ('pressure', 'velocity'),
                   ('sound_velocity', 'density', 'absorption_coef'), _acoustic1d),
    'Acoustic2D': (2, ('pressure', 'velocity_x', 'velocity_y'),
                   ('sound_velocity', 'density', 'absorption_coef'), _acoustic2d),
    'Thermal2D': (5, ('temperature', 'heat_flux_x', 'heat_flux_y'),
                  ('density', 'heat_capacity', 'thermal_conductivity_x', 'thermal_conductivity_y'),
                  _thermal2d),
}


def _boundary_table(component, first_step, n_steps, signals):
    """``FieldComponent.boundaries`` -> the CSR table of fds_upload_boundaries: distinct cells
    ascending, the operations of a cell in list order, v = alpha * v + (value | signal[step])."""
    cells, order, alpha, value, signal = [], [], [], [], []
    for position, bound in enumerate(component.boundaries):
        idx = np.asarray(bound.region.indices, dtype=np.int64).reshape(-1)
        n = idx.shape[0]
        if np.ndim(bound.value) == 0:                              # pyfds/regions.py:137-138
            val, sig = np.full(n, float(bound.value)), np.full(n, -1, dtype=np.int32)
        else:
            if isinstance(bound.value, np.ndarray) and np.ndim(bound.value) == 1:
                rows = [bound.value] * n                           # one signal for all points (:140)
            elif isinstance(bound.value, np.ndarray):
                rows = list(bound.value.T)
            else:
                rows = list(bound.value)                           # one signal per point (:143-145)
            base = len(signals)
            for row in rows:
                window = np.asarray(row, dtype=np.float64)[first_step:first_step + n_steps]
                if window.shape[0] != n_steps:
                    raise IndexError('boundary signal shorter than the simulated steps')
                signals.append(window)
            val, sig = np.zeros(n), np.arange(base, base + n, dtype=np.int32)
        # duplicate indices inside one region: NumPy fancy assignment keeps the last write
        _, last = np.unique(idx[::-1], return_index=True)
        keep = np.sort(n - 1 - last)
        cells.append(idx[keep])
        order.append(np.full(keep.shape[0], position))
        alpha.append(np.full(keep.shape[0], float(bound.additive)))
        value.append(val[keep])
        signal.append(sig[keep])
    if not cells:
        return (np.zeros(0, np.int64), np.zeros(1, np.int32), np.zeros(0), np.zeros(0),
                np.zeros(0, np.int32))
    cells, order = np.concatenate(cells), np.concatenate(order)
    perm = np.lexsort((order, cells))
    unique_cells, start = np.unique(cells[perm], return_index=True)
    offsets = np.append(start, cells.shape[0]).astype(np.int32)
    return (unique_cells.astype(np.int64), offsets, np.concatenate(alpha)[perm],
            np.concatenate(value)[perm], np.concatenate(signal)[perm].astype(np.int32))


def simulate(field, num_steps, device=0):
    """``num_steps`` x ``field.sim_step()`` (pyfds/fields.py:87-93) as ONE library call."""
    model, names, params, coefficients = MODELS[type(field).__name__]
    nx = field.x.samples
    ny = field.y.samples if hasattr(field, 'y') else 1
    n = field.num_points
    components = [getattr(field, name) for name in names]

    # 1. materials: distinct parameter combinations become material ids 1..m (0 = outside the grid)
    vectors = np.stack([np.asarray(field.material_vector(p), dtype=np.float64) for p in params])
    distinct, ids = np.unique(vectors, axis=1, return_inverse=True)
    tables, lossy = coefficients(field, *distinct)
    ctx = ct.c_void_p()
    desc = fds_desc(model, device, nx, ny, 0, ny, 0, int(lossy), distinct.shape[1], 0)
    if _lib.fds_create(ct.byref(desc), ct.byref(ctx)) != 0:
        raise RuntimeError(_lib.fds_last_error(None).decode())
    try:
        ids8 = np.ascontiguousarray(ids.reshape(-1) + 1, dtype=np.uint8)
        _check(ctx, _lib.fds_upload_material_map(ctx, _ptr(ids8), ct.c_int64(n)))
        for which, column in tables.items():
            column = np.ascontiguousarray(np.concatenate(([0.0], column)))    # entry 0: void
            _check(ctx, _lib.fds_upload_table(ctx, which, _ptr(column), ct.c_int64(column.size)))

        # 2. boundaries, signals and probes of every component
        signals = []
        for c, component in enumerate(components):
            cells, offsets, alpha, value, signal = (
                np.ascontiguousarray(a) for a in
                _boundary_table(component, field.step, num_steps, signals))
            _check(ctx, _lib.fds_upload_boundaries(
                ctx, c, _ptr(cells), _ptr(offsets), ct.c_int64(cells.size), _ptr(alpha), _ptr(value),
                _ptr(signal), ct.c_int64(alpha.size)))
        samples = np.ascontiguousarray(signals, dtype=np.float64).reshape(len(signals), num_steps)
        _check(ctx, _lib.fds_upload_signals(ctx, _ptr(samples), ct.c_int64(len(signals)),
                                            ct.c_int64(num_steps), ct.c_int64(field.step)))
        n_slots = sum(len(o.region.indices) for comp in components for o in comp.outputs)
        slot = 0
        layout = []
        for c, component in enumerate(components):
            cells, slots = [], []
            for output in component.outputs:
                idx = np.asarray(output.region.indices, dtype=np.int64).reshape(-1)
                cells.append(idx)
                slots.append(slot + np.arange(idx.shape[0]))
                layout.append((output, slot, idx.shape[0]))
                slot += idx.shape[0]
            cells = np.concatenate(cells) if cells else np.zeros(0, np.int64)
            slots = np.concatenate(slots) if slots else np.zeros(0, np.int64)
            perm = np.argsort(cells, kind='stable')
            cells = np.ascontiguousarray(cells[perm], dtype=np.int64)
            slots = np.ascontiguousarray(slots[perm], dtype=np.int32)
            _check(ctx, _lib.fds_upload_probes(ctx, c, _ptr(cells), _ptr(slots),
                                               ct.c_int64(cells.size), ct.c_int64(n_slots)))

        # 3. values in, steps, values and probe records out -- one call
        arrays = [np.ascontiguousarray(comp.values, dtype=np.float64) for comp in components]
        pointers = (ct.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        records = np.zeros((num_steps, n_slots))
        _check(ctx, _lib.fds_simulate(ctx, ct.c_int64(field.step), ct.c_int64(num_steps), pointers,
                                      pointers, _ptr(records) if n_slots else None))
        for comp, array in zip(components, arrays):
            comp.values = array
        for output, first, count in layout:                        # pyfds/fields.py:606-611
            block = records[:, first:first + count]
            if not output.signals:
                output.signals = block.T.tolist()
            else:
                for k, sig in enumerate(output.signals[:count]):
                    sig.extend(block[:, k].tolist())
        field.step += num_steps
    finally:
        _lib.fds_destroy(ctx)


def install(pyfds, library_path):
    """Loads the library and routes ``Field.simulate`` of the supported models through it. Classes the
    stub does not know (and subclasses that override ``sim_step``) keep the reference's Python loop."""
    global _lib
    _lib = ct.CDLL(library_path)
    _lib.fds_last_error.restype = ct.c_char_p
    _lib.fds_last_error.argtypes = [ct.c_void_p]
    original = pyfds.fields.Field.simulate

    def patched(self, num_steps=None):
        known = MODELS.get(type(self).__name__)
        if known is None or type(self).sim_step is not getattr(pyfds, type(self).__name__).sim_step:
            return original(self, num_steps)
        simulate(self, int(num_steps) if num_steps else self.t.samples)

    pyfds.fields.Field.simulate = patched
    return original
