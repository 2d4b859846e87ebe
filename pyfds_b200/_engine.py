"""ctypes binding of ``libfdsb200.so`` (C ABI: ``include/fdsb200.h``) and the host driver that runs
``Field.simulate`` / ``sim_step`` on it.

The division of labour follows the reference seam (``pyfds/fields.py:59-95``): the model class says
*what* to simulate (materials, boundaries, outputs, values), this module moves that description to the
device once per call and lets the CUDA engine execute all requested steps back to back.

There is deliberately no CPU path here: if the library or a GPU is missing, ``run`` raises.
"""

import ctypes as ct
import os
import weakref

import numpy as np

from . import _bake

MODEL_IDS = {
    'acoustic1d': 1, 'acoustic2d': 2, 'acoustic3daxi': 3,
    'thermal1d': 4, 'thermal2d': 5, 'thermal3daxi': 6,
}

# enum fds_table / fds_column_table / fds_column_vector
TAB = {'GX': 0, 'GY': 1, 'FX': 2, 'FY': 3, 'VM1': 4, 'VP1': 5, 'VMN': 6, 'VPN': 7, 'V0': 8, 'EB': 9}
CTAB = {'FX': 0, 'VM1': 1, 'VP1': 2}
CVEC = {'R': 0, 'RR': 1}

#: a field is only cut into slabs of at least this many rows (twice the 4 halo rows a sweep consumes)
MIN_SLAB_ROWS = 8

#: upper bound for one probe drain buffer (steps per fds_step call are chunked to stay below it)
MAX_PROBE_BYTES = 256 << 20


class fds_desc(ct.Structure):
    _fields_ = [
        ('model', ct.c_int32), ('device', ct.c_int32),
        ('nx', ct.c_int64), ('ny', ct.c_int64), ('row0', ct.c_int64), ('rows', ct.c_int64),
        ('halo_rows', ct.c_int32), ('lossy', ct.c_int32), ('n_materials', ct.c_int32),
        ('kernel', ct.c_int32),
    ]


_LIB = None


def library_path():
    # FDS_LIBRARY_PATH: load an alternative build of the same library (kernel tuning experiments)
    override = os.environ.get('FDS_LIBRARY_PATH')
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libfdsb200.so')


def load_library():
    """Loads the CUDA engine. Raises ``RuntimeError`` if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            'libfdsb200.so is missing ({}). Build it with `python -m pyfds_b200._build`; this '
            'package has no CPU implementation of the time-stepping path.'.format(path))
    lib = ct.CDLL(path)
    p = ct.c_void_p
    i32, i64 = ct.c_int32, ct.c_int64
    dptr = ct.POINTER(ct.c_double)
    sigs = {
        'fds_create': (ct.c_int, [ct.POINTER(fds_desc), ct.POINTER(p)]),
        'fds_destroy': (None, [p]),
        'fds_last_error': (ct.c_char_p, [p]),
        'fds_device_count': (ct.c_int, []),
        'fds_upload_material_map': (ct.c_int, [p, p, i64]),
        'fds_upload_table': (ct.c_int, [p, i32, p, i64]),
        'fds_upload_cell_table': (ct.c_int, [p, i32, p, i64]),
        'fds_upload_column_table': (ct.c_int, [p, i32, p, i64]),
        'fds_upload_column_vector': (ct.c_int, [p, i32, p, i64]),
        'fds_upload_boundaries': (ct.c_int, [p, i32, p, p, i64, p, p, p, i64]),
        'fds_upload_signals': (ct.c_int, [p, p, i64, i64, i64]),
        'fds_upload_probes': (ct.c_int, [p, i32, p, p, i64, i64]),
        'fds_upload_state': (ct.c_int, [p, i32, p, i64]),
        'fds_download_state': (ct.c_int, [p, i32, p, i64]),
        'fds_host_register': (ct.c_int, [p, i64]),
        'fds_host_unregister': (ct.c_int, [p]),
        'fds_reset_state': (ct.c_int, [p]),
        'fds_step': (ct.c_int, [p, i64, i64, p]),
        'fds_step_async': (ct.c_int, [p, i64, i64]),
        'fds_step_prepare': (ct.c_int, [p, i64, i64]),
        'fds_simulate': (ct.c_int, [p, i64, i64, ct.POINTER(p), ct.POINTER(p), p]),
        'fds_last_pipeline_bands': (ct.c_int, [p, ct.POINTER(i64)]),
        'fds_sync': (ct.c_int, [p]),
        'fds_set_flow': (ct.c_int, [p, p, i64]),
        'fds_last_flow_shifts': (ct.c_int, [p, ct.POINTER(i64)]),
        'fds_snapshot_async': (ct.c_int, [p, i32, i32, i32, i32]),
        'fds_snapshot_wait': (ct.c_int, [p, i32, p, i64]),
        'fds_comm_unique_id': (ct.c_int, [p]),
        'fds_comm_init': (ct.c_int, [p, p, i32, i32]),
        'fds_peer_export': (ct.c_int, [p, p]),
        'fds_peer_import': (ct.c_int, [p, i32, p, i64]),
        'fds_slab_init': (ct.c_int, [p, i32, i32]),
        'fds_peer_connect': (ct.c_int, [p, i32, p]),
        'fds_last_step_ms': (ct.c_int, [p, dptr]),
        'fds_last_launch_info': (ct.c_int, [p, ct.POINTER(i64), ct.POINTER(i64),
                                            ct.POINTER(ct.c_char_p)]),
        'fds_device_bytes': (i64, [p]),
        'fds_group_create': (ct.c_int, [ct.POINTER(p), i32, ct.POINTER(p)]),
        'fds_group_destroy': (None, [p]),
        'fds_group_last_error': (ct.c_char_p, [p]),
        'fds_group_add_linear': (ct.c_int, [p, i32, i32, i32, i32, ct.c_double, i32, i32, i64, p]),
        'fds_group_add_viscous_heating': (ct.c_int, [p, i32, i32, p, p, p, ct.c_double, i32, i64, p]),
        'fds_group_add_material_law': (ct.c_int, [p, i32, i32, i32, i32, i32, ct.c_double,
                                                   ct.c_double, i32, ct.c_double, i64, p, p, p]),
        'fds_group_step': (ct.c_int, [p, i64, i64, ct.POINTER(p)]),
        'fds_group_read': (ct.c_int, [p, i32, p, ct.POINTER(i64)]),
        'fds_stream_stats': (ct.c_int, [p, ct.POINTER(i64)]),
    }
    for name, (restype, argtypes) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def _ptr(array):
    return array.ctypes.data_as(ct.c_void_p)


def _c(array, dtype):
    return np.ascontiguousarray(array, dtype=dtype)


class EngineError(RuntimeError):
    pass


class Engine:
    """One device context = one y-slab of one field (the whole grid on a single GPU)."""

    def __init__(self, model, nx, ny, n_materials, lossy, row0=0, rows=None, halo_rows=0,
                 device=0, kernel=0):
        self.lib = load_library()
        rows = ny if rows is None else rows
        self.desc = fds_desc(MODEL_IDS[model], device, nx, ny, row0, rows, halo_rows, int(lossy),
                             n_materials, kernel)
        self.model = model
        self.nx, self.ny, self.row0, self.rows, self.halo_rows = nx, ny, row0, rows, halo_rows
        self.n_materials = n_materials
        self.lossy = bool(lossy)
        self.owned = rows * nx
        self.ncomp = 2 if model.endswith('1d') else 3
        handle = ct.c_void_p()
        if self.lib.fds_create(ct.byref(self.desc), ct.byref(handle)) != 0:
            raise EngineError(self.lib.fds_last_error(None).decode())
        self.handle = handle
        self._finalizer = weakref.finalize(self, self.lib.fds_destroy, handle)

    def close(self):
        self._finalizer()

    def _check(self, rc):
        if rc != 0:
            raise EngineError(self.lib.fds_last_error(self.handle).decode())

    # window of global cells held by this context, halo rows included
    @property
    def cell_lo(self):
        return (self.row0 - self.halo_rows) * self.nx

    @property
    def cell_hi(self):
        return (self.row0 + self.rows + self.halo_rows) * self.nx

    def upload_material_map(self, ids):
        ids = _c(ids, np.uint8)
        self._check(self.lib.fds_upload_material_map(self.handle, _ptr(ids), ids.size))

    def upload_table(self, which, values):
        values = _c(values, np.float64)
        self._check(self.lib.fds_upload_table(self.handle, which, _ptr(values), values.size))

    def upload_cell_table(self, which, values):
        """Per-cell coefficients of a 1-D field (``None``: back to the per-material tables)."""
        if values is None:
            self._check(self.lib.fds_upload_cell_table(self.handle, 0, None, 0))
            return
        values = _c(values, np.float64)
        self._check(self.lib.fds_upload_cell_table(self.handle, which, _ptr(values), values.size))

    def upload_column_table(self, which, values):
        values = _c(values, np.float64)
        self._check(self.lib.fds_upload_column_table(self.handle, which, _ptr(values), values.size))

    def upload_column_vector(self, which, values):
        values = _c(values, np.float64)
        self._check(self.lib.fds_upload_column_vector(self.handle, which, _ptr(values),
                                                      values.size))

    def upload_boundaries(self, component, table):
        cells, offsets = _c(table.cells, np.int64), _c(table.offsets, np.int32)
        alpha, value = _c(table.alpha, np.float64), _c(table.value, np.float64)
        signal = _c(table.signal, np.int32)
        self._check(self.lib.fds_upload_boundaries(
            self.handle, component, _ptr(cells), _ptr(offsets), cells.size, _ptr(alpha),
            _ptr(value), _ptr(signal), alpha.size))

    def upload_signals(self, samples, first_step):
        samples = _c(samples, np.float64)
        n_signals, n_steps = samples.shape if samples.ndim == 2 else (0, 0)
        self._check(self.lib.fds_upload_signals(self.handle, _ptr(samples), n_signals, n_steps,
                                                first_step))

    def upload_probes(self, component, cells, slots, n_slots_total):
        cells, slots = _c(cells, np.int64), _c(slots, np.int32)
        self._check(self.lib.fds_upload_probes(self.handle, component, _ptr(cells), _ptr(slots),
                                               cells.size, n_slots_total))

    def upload_state(self, component, values):
        values = _c(values, np.float64)
        self._check(self.lib.fds_upload_state(self.handle, component, _ptr(values), values.size))

    def download_state(self, component, out=None):
        if out is None:
            out = np.empty(self.owned, dtype=np.float64)
        self._check(self.lib.fds_download_state(self.handle, component, _ptr(out), out.size))
        return out

    def reset_state(self):
        self._check(self.lib.fds_reset_state(self.handle))

    def step(self, first_step, n_steps, n_slots):
        """Runs n_steps steps synchronously; returns the probe records [n_steps][n_slots]."""
        probes = np.zeros((n_steps, n_slots), dtype=np.float64)
        self._check(self.lib.fds_step(self.handle, first_step, n_steps,
                                      _ptr(probes) if n_slots else None))
        return probes

    def simulate(self, first_step, n_steps, values_in, values_out, n_slots):
        """Upload, ``n_steps`` steps and download in one call (``fds_simulate``: overlapped by row
        bands where that pays); returns the probe records [n_steps][n_slots]."""
        probes = np.zeros((n_steps, n_slots), dtype=np.float64)
        ins = (ct.c_void_p * len(values_in))(*[a.ctypes.data for a in values_in])
        outs = (ct.c_void_p * len(values_out))(*[a.ctypes.data for a in values_out])
        self._check(self.lib.fds_simulate(self.handle, first_step, n_steps, ins, outs,
                                          _ptr(probes) if n_slots else None))
        return probes

    def last_pipeline_bands(self):
        bands = ct.c_int64()
        self._check(self.lib.fds_last_pipeline_bands(self.handle, ct.byref(bands)))
        return bands.value

    def step_prepare(self, first_step, n_steps):
        """All allocations and loads of a following ``step`` over the same steps, without stepping."""
        self._check(self.lib.fds_step_prepare(self.handle, first_step, n_steps))

    def step_async(self, first_step, n_steps):
        self._check(self.lib.fds_step_async(self.handle, first_step, n_steps))

    def sync(self):
        self._check(self.lib.fds_sync(self.handle))

    def last_step_ms(self):
        ms = ct.c_double()
        self._check(self.lib.fds_last_step_ms(self.handle, ct.byref(ms)))
        return ms.value

    def last_launch_info(self):
        launches, spl, name = ct.c_int64(), ct.c_int64(), ct.c_char_p()
        self._check(self.lib.fds_last_launch_info(self.handle, ct.byref(launches), ct.byref(spl),
                                                  ct.byref(name)))
        return launches.value, spl.value, (name.value or b'').decode()

    def set_flow(self, periods):
        """Per-row shift periods of a flowing medium (``None`` switches the flow off)."""
        if periods is None:
            self._check(self.lib.fds_set_flow(self.handle, None, 0))
            return
        periods = _c(periods, np.int64)
        self._check(self.lib.fds_set_flow(self.handle, _ptr(periods), periods.size))

    def last_flow_shifts(self):
        shifts = ct.c_int64()
        self._check(self.lib.fds_last_flow_shifts(self.handle, ct.byref(shifts)))
        return shifts.value

    def frame_shape(self, stride_x, stride_y):
        return (-(-self.rows // stride_y), -(-self.nx // stride_x))

    def snapshot_async(self, component, stride_x, stride_y, slot):
        """Enqueues a decimated snapshot of one component behind the steps enqueued so far."""
        self._check(self.lib.fds_snapshot_async(self.handle, component, stride_x, stride_y, slot))

    def snapshot_wait(self, slot, shape):
        frame = np.empty(shape, dtype=np.float64)
        self._check(self.lib.fds_snapshot_wait(self.handle, slot, _ptr(frame), frame.size))
        return frame

    def device_bytes(self):
        return self.lib.fds_device_bytes(self.handle)

    def stream_stats(self):
        """Counters of the streaming kernel (all zero unless FDS_STREAM_STATS was set at creation):
        entries into the branch-free body by variant [0..4], general rows [5], rows streamed [6], row
        pairs of the lossy axisymmetric model stepped with the IEEE division [7]."""
        out = (ct.c_int64 * 8)()
        self._check(self.lib.fds_stream_stats(self.handle, out))
        return list(out)

    def peer_export(self):
        buf = (ct.c_uint8 * (7 * 64))()
        self._check(self.lib.fds_peer_export(self.handle, buf))
        return bytes(buf)

    def peer_import(self, side, handles, neighbour_rows):
        buf = (ct.c_uint8 * (7 * 64)).from_buffer_copy(bytes(handles))
        self._check(self.lib.fds_peer_import(self.handle, side, buf, neighbour_rows))

    def slab_init(self, rank, world):
        """This context is slab ``rank`` of ``world`` slabs that all live in this process."""
        self._check(self.lib.fds_slab_init(self.handle, rank, world))

    def peer_connect(self, side, neighbour):
        """Wires the neighbour slab's context (same process) to side 0 (lower) / 1 (upper)."""
        self._check(self.lib.fds_peer_connect(self.handle, side, neighbour.handle))

    def comm_init(self, unique_id, rank, world):
        buf = (ct.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._check(self.lib.fds_comm_init(self.handle, buf, rank, world))


class Group:
    """Device form of a ``SynchronizedFields`` group of 1-D fields (``fds_group_*``): the members step
    in lock-step on one stream and the built-in interactions run as kernels in between."""

    def __init__(self, engines):
        self.lib = load_library()
        self.engines = list(engines)
        handles = (ct.c_void_p * len(self.engines))(*[e.handle for e in self.engines])
        handle = ct.c_void_p()
        if self.lib.fds_group_create(handles, len(self.engines), ct.byref(handle)) != 0:
            raise EngineError(self.lib.fds_group_last_error(None).decode())
        self.handle = handle
        self.n = self.engines[0].nx
        self.count = 0
        self._finalizer = weakref.finalize(self, self.lib.fds_group_destroy, handle)

    def close(self):
        self._finalizer()

    def _check(self, rc):
        if rc != 0:
            raise EngineError(self.lib.fds_group_last_error(self.handle).decode())

    @staticmethod
    def _optional(values, n):
        """Array of n doubles, or NULL for the integer 0 the reference starts its sums with."""
        if np.ndim(values) == 0:
            return None, None
        array = _c(np.broadcast_to(np.asarray(values, dtype=np.float64), (n,)), np.float64)
        return array, _ptr(array)

    def add_linear(self, source, target, scale, additive, accumulate, stepping, accumulated=0):
        keep, pointer = self._optional(accumulated, self.n)
        self._check(self.lib.fds_group_add_linear(
            self.handle, source[0], source[1], target[0], target[1], float(scale), int(additive),
            int(accumulate), int(stepping), pointer))
        self.count += 1
        return self.count - 1

    def add_viscous_heating(self, sound, heat, density, gradient_factor, gain, dt, accumulate,
                            stepping, accumulated=0):
        arrays = [_c(a, np.float64) for a in (density, gradient_factor, gain)]
        keep, pointer = self._optional(accumulated, self.n)
        self._check(self.lib.fds_group_add_viscous_heating(
            self.handle, sound, heat, _ptr(arrays[0]), _ptr(arrays[1]), _ptr(arrays[2]), float(dt),
            int(accumulate), int(stepping), pointer))
        self.count += 1
        return self.count - 1

    def add_material_law(self, source, target, parameter, law, p0, p1, threshold, stepping, statics,
                         last, scalars):
        statics = _c(statics, np.float64)
        scalars = _c(scalars, np.float64)
        keep, pointer = self._optional(last, self.n)
        self._check(self.lib.fds_group_add_material_law(
            self.handle, source[0], source[1], target, int(parameter), int(law), float(p0),
            float(p1), int(threshold is not None), float(threshold or 0.0), int(stepping),
            _ptr(statics), pointer, _ptr(scalars)))
        self.count += 1
        return self.count - 1

    def step(self, first_step, n_steps, slots):
        """``n_steps`` common steps; returns the probe records of every member."""
        records = [np.zeros((n_steps, n), dtype=np.float64) for n in slots]
        pointers = (ct.c_void_p * len(records))(
            *[r.ctypes.data if r.size else None for r in records])
        self._check(self.lib.fds_group_step(self.handle, first_step, n_steps, pointers))
        return records

    def read(self, interaction):
        values = np.zeros(self.n, dtype=np.float64)
        count = ct.c_int64()
        self._check(self.lib.fds_group_read(self.handle, interaction, _ptr(values),
                                            ct.byref(count)))
        return values, count.value


def comm_unique_id():
    lib = load_library()
    buf = (ct.c_uint8 * 128)()
    if lib.fds_comm_unique_id(buf) != 0:
        raise EngineError(lib.fds_last_error(None).decode())
    return bytes(buf)


# ---------------------------------------------------------------------------------------------
# host driver
# ---------------------------------------------------------------------------------------------

#: host arrays at least this large are page-locked while a field keeps using them
PIN_THRESHOLD_BYTES = 32 << 20


def _unregister_all(lib, pinned):
    for _, (array, address) in list(pinned.items()):
        lib.fds_host_unregister(ct.c_void_p(address))
    pinned.clear()


class _State:
    """Per-field device state; lives in ``field.__dict__['_engine_state']`` and is never pickled."""

    def __init__(self):
        self.engine = None
        self.epoch = -1
        self.key = None
        # component index -> (array, address): arrays currently page-locked (the reference keeps
        # updating the same ``values`` arrays in place, so the registration is reused call after call;
        # holding the array keeps its memory alive until it is unregistered)
        self.pinned = {}
        self._finalizer = None

    def pin(self, lib, index, array):
        """Page-locks ``array`` for component ``index`` unless it already is; best effort."""
        if array.nbytes < PIN_THRESHOLD_BYTES or not array.flags.writeable:
            return
        address = array.ctypes.data
        current = self.pinned.get(index)
        if current is not None and current[1] == address and current[0].nbytes == array.nbytes:
            return
        if current is not None:
            lib.fds_host_unregister(ct.c_void_p(current[1]))
            del self.pinned[index]
        if lib.fds_host_register(ct.c_void_p(address), array.nbytes) == 0:
            self.pinned[index] = (array, address)
            if self._finalizer is None:
                self._finalizer = weakref.finalize(self, _unregister_all, lib, self.pinned)


def _components(field):
    return [getattr(field, name) for name in field._device_components]


def _grid(field):
    nx = field.x.samples
    ny = field.y.samples if hasattr(field, 'y') else 1
    return nx, ny


def halo_rows_for(field, world):
    """Rows a slab needs from each neighbour for one step: 1 (lossless) or 2 (viscous operator)."""
    if world <= 1 or not hasattr(field, 'y'):
        return 0
    return 2 if field._baked['lossy'] else 1


def prepare(field, device=0, row0=0, rows=None, halo_rows=0, kernel=None, lossy=None,
            per_cell=False, state=None):
    """Creates (or reuses) the device context of ``field`` and uploads what ``assemble_matrices``
    froze: the material map and the coefficient tables. Returns the ``Engine``.

    ``kernel``: 0 automatic, 1 one-step kernel, 2 streaming multi-step kernel (``fds_desc.kernel``);
    defaults to the field attribute ``device_kernel`` (0 if absent). ``lossy``: True if ANY cell of the
    grid is lossy -- every slab of a multi-GPU run must pick the same kernel and the same number of
    steps per launch, whatever materials its own rows hold (default: decided from this window).
    ``per_cell``: give a 1-D field with per-point material vectors per-cell coefficient arrays even if
    its distinct value combinations would fit the material table (targets of a device-side material
    law, whose coefficients are rewritten cell by cell). ``state``: where the context is kept (default:
    the field's own slot; a run over several slabs in one process keeps one per slab)."""
    if kernel is None:
        kernel = getattr(field, 'device_kernel', 0)
    if state is None:
        state = field.__dict__.get('_engine_state')
        if state is None:
            state = field.__dict__['_engine_state'] = _State()
    baked = field._baked
    nx, ny = _grid(field)
    rows = ny if rows is None else rows
    key = (field._device_model, nx, ny, row0, rows, halo_rows, device, kernel, bool(per_cell))
    if state.engine is not None and state.epoch == baked['epoch'] and state.key == key:
        return state.engine

    snapshot = baked['snapshot']
    lo, hi = (row0 - halo_rows) * nx, (row0 + rows + halo_rows) * nx
    one_d = not hasattr(field, 'y')
    # per-cell coefficients: 1-D models, and plain 2-D models on a single slab
    cells_possible = one_d or (field._device_model in ('acoustic2d', 'thermal2d') and
                               row0 == 0 and rows == ny and halo_rows == 0)
    cells = None
    if isinstance(snapshot, _bake.DenseSnapshot) and cells_possible and \
            (per_cell or _bake.distinct_combinations(snapshot) > _bake.MAX_MATERIALS):
        # materials that differ from cell to cell (MaterialCoupling on a smooth source): the kernels
        # read per-cell coefficient arrays, evaluated here by the same expressions on the per-point
        # parameter vectors themselves (2-D: the one-thread-per-cell kernel, one step per launch)
        values = {p: np.concatenate(([0.0], snapshot.vectors[p])) for p in snapshot.params}
        cells = field._coefficient_tables(values)
        ids = np.ones(field.num_points, dtype=np.uint8)
        tables = {'tables': {name: np.zeros(1) for name in cells['tables']},
                  'lossy': cells.get('lossy', False)}
        n_materials = 1
    else:
        ids, values = _bake.material_ids(snapshot, field.num_points, nx, lo, hi)
        tables = field._coefficient_tables(values)
        n_materials = len(next(iter(values.values()))) - 1
    lossy = bool(tables.get('lossy', False)) or bool(lossy)
    baked['lossy'] = lossy

    engine = state.engine
    if engine is None or state.key != key or engine.n_materials != n_materials or \
            engine.lossy != lossy:
        if engine is not None:
            engine.close()
        engine = Engine(field._device_model, nx, ny, n_materials, lossy, row0=row0, rows=rows,
                        halo_rows=halo_rows, device=device, kernel=kernel)
    engine.upload_material_map(ids)
    for name, column in tables['tables'].items():
        engine.upload_table(TAB[name], np.concatenate(([0.0], column)))
    for name, matrix in tables.get('column_tables', {}).items():
        engine.upload_column_table(CTAB[name], np.vstack((np.zeros((1, nx)), matrix)))
    for name, vector in tables.get('column_vectors', {}).items():
        engine.upload_column_vector(CVEC[name], vector)
    if cells_possible:
        engine.upload_cell_table(0, None)
        if cells is not None:
            for name, column in cells['tables'].items():
                engine.upload_cell_table(TAB[name], column)
    state.engine, state.epoch, state.key = engine, baked['epoch'], key
    return engine


def upload_run_tables(field, engine, first_step, n_steps):
    """Bakes and uploads what may change between two ``simulate`` calls: boundary operations, the
    signal window for the step range, and the probe points. Returns ``(n_slots, layout)`` where layout
    lists ``(output, first_slot, n_points)``."""
    nx = engine.nx
    periods = field._device_flow()
    # all grid rows: every slab derives the same launch schedule from them (fds_set_flow)
    engine.set_flow(periods)
    signals = []
    sent = engine.__dict__.setdefault('_sent_tables', {})
    for c, component in enumerate(_components(field)):
        table = _bake.boundary_table(component.boundaries, first_step, n_steps, engine.cell_lo,
                                     engine.cell_hi, signals)
        # the device addresses cells relative to the first *owned* cell
        table.cells = table.cells - engine.halo_rows * nx
        # Unchanged since the last call (an animator calling simulate(20) again and again): nothing to
        # send -- an upload marks the cell flags dirty and with them the task tables of the kernels.
        mark = (table.cells, table.offsets, table.alpha, table.value, table.signal)
        if not _same_arrays(sent.get(('bounds', c)), mark):
            engine.upload_boundaries(c, table)
            sent[('bounds', c)] = mark
    engine.upload_signals(np.array(signals, dtype=np.float64).reshape(len(signals), n_steps)
                          if signals else np.zeros((0, 0)), first_step)

    layout = []
    tables = []
    slot = 0
    own_lo, own_hi = engine.row0 * nx, (engine.row0 + engine.rows) * nx
    for c, component in enumerate(_components(field)):
        base = slot
        for output in component.outputs:
            count = np.asarray(output.region.indices).reshape(-1).shape[0]
            layout.append((output, slot, count))
            slot += count
        cells, slots, _ = _bake.probe_table(component.outputs, base, own_lo, own_hi)
        tables.append((c, cells, slots))
    for c, cells, slots in tables:
        mark = (cells, slots, np.asarray([slot]))
        if not _same_arrays(sent.get(('probes', c)), mark):
            engine.upload_probes(c, cells, slots, slot)
            sent[('probes', c)] = mark
    return slot, layout


def _same_arrays(kept, fresh):
    """True if ``kept`` (the arrays of the last upload, or None) equal ``fresh`` element for element,
    dtype and shape included; floats are compared as bit patterns (a -0.0 is not a 0.0 here)."""
    if kept is None or len(kept) != len(fresh):
        return False
    for a, b in zip(kept, fresh):
        a, b = np.asarray(a), np.asarray(b)
        if a.dtype != b.dtype or a.shape != b.shape:
            return False
        if a.dtype == np.float64:
            a, b = np.ascontiguousarray(a).view(np.int64), np.ascontiguousarray(b).view(np.int64)
        if not np.array_equal(a, b):
            return False
    return True


def _fingerprint(*arrays):
    import hashlib
    digest = hashlib.blake2b(digest_size=16)
    for array in arrays:
        array = np.ascontiguousarray(array)
        digest.update(str((array.dtype.str, array.shape)).encode())
        digest.update(array.tobytes())
    return digest.digest()


def _append_signals(layout, records):
    """``Output.signals[k]`` grows by one sample per step (``pyfds/fields.py:606-611``)."""
    for output, first, count in layout:
        block = records[:, first:first + count]
        if not output.signals:
            output.signals = block.T.tolist()
        else:
            for k, signal in enumerate(output.signals[:count]):
                signal.extend(block[:, k].tolist())


def _host_values(component, num_points):
    values = np.asarray(component.values, dtype=np.float64).reshape(-1)
    if values.shape[0] != num_points:
        raise ValueError('Field component has {} values, the field has {} points.'.format(
            values.shape[0], num_points))
    return values


def upload_values(field, engine):
    """Host ``values`` of every component -> device. Arrays the field keeps reusing are page-locked
    on first use (the reference updates ``values`` in place, so the registration is reused call after
    call). Returns the seconds spent page-locking and copying."""
    import time
    clock = time.perf_counter
    t0 = clock()
    state = field.__dict__['_engine_state']
    host = []
    for c, component in enumerate(_components(field)):
        values = _host_values(component, field.num_points)
        own = component.values
        if isinstance(own, np.ndarray) and own.ctypes.data == values.ctypes.data and \
                own.nbytes == values.nbytes:
            state.pin(engine.lib, c, own)
        host.append(values)
    t1 = clock()
    for c, values in enumerate(host):
        engine.upload_state(c, values)
    return t1 - t0, clock() - t1


def download_values(field, engine, only=None):
    """Device -> host ``values`` (in place where the component holds a suitable array, so that
    page-locked arrays keep being reused). ``only``: component numbers to fetch (default all)."""
    for c, component in enumerate(_components(field)):
        if only is not None and c not in only:
            continue
        target = component.values
        if isinstance(target, np.ndarray) and target.dtype == np.float64 and \
                target.flags.c_contiguous and target.flags.writeable and \
                target.shape == (field.num_points,):
            engine.download_state(c, out=target)
        else:
            component.values = engine.download_state(c)


def run(field, n_steps, progress_logger=None, advance=True):
    """``n_steps`` x ``sim_step`` on the device: upload values, step, download values and probes.
    Wall-clock seconds of the phases are left in ``field.__dict__['_last_run_profile']``."""
    import time
    clock = time.perf_counter
    t0 = clock()
    # optional field attribute ``devices`` (or FDS_DEVICES=n): the CUDA ordinals of ONE process among
    # which a 2-D field is cut into y-slabs -- the same simulate() call then runs on all of them
    devices = _slab_devices(field)
    if devices is not None:
        from . import parallel
        slabs = field.__dict__.get('_local_slabs')
        if slabs is None or slabs.devices != devices:
            if slabs is not None:
                slabs.close()
            slabs = field.__dict__['_local_slabs'] = parallel.LocalSlabs(field, devices)
        slabs.simulate(n_steps, progress_logger)
        if advance:
            field.step += n_steps
        return
    # optional field attribute ``device_index``: CUDA ordinal to run on (default 0)
    engine = prepare(field, device=int(getattr(field, 'device_index', 0)))
    first_step = field.step
    n_slots, layout = upload_run_tables(field, engine, first_step, n_steps)
    t1 = clock()

    chunk = n_steps
    if n_slots:
        chunk = max(1, min(chunk, MAX_PROBE_BYTES // (8 * n_slots)))
    if progress_logger is not None:
        chunk = max(1, min(chunk, -(-n_steps // 20)))
    if chunk == n_steps and getattr(field, 'device_single_call', True):
        # the whole call in one engine call: values up, steps, values down (fds_simulate)
        state = field.__dict__['_engine_state']
        components = _components(field)
        ins, outs = [], []
        for c, component in enumerate(components):
            values = _host_values(component, field.num_points)
            own = component.values
            in_place = isinstance(own, np.ndarray) and own.ctypes.data == values.ctypes.data and \
                own.nbytes == values.nbytes and own.flags.writeable
            if in_place:
                state.pin(engine.lib, c, own)
            ins.append(values)
            outs.append(values if in_place else np.empty(field.num_points, dtype=np.float64))
        t2 = clock()
        records = engine.simulate(first_step, n_steps, ins, outs, n_slots)
        for component, result in zip(components, outs):
            if component.values is not result:
                component.values = result
        t3 = clock()
        if n_slots:
            _append_signals(layout, records)
        if progress_logger is not None:
            for s in range(first_step, first_step + n_steps):
                progress_logger.log(s)
        field.__dict__['_last_run_profile'] = {
            'prepare_and_tables_s': t1 - t0, 'page_lock_s': t2 - t1, 'simulate_call_s': t3 - t2,
            'signals_s': clock() - t3, 'pipeline_bands': engine.last_pipeline_bands()}
        _log_throughput(field, engine, n_steps, t3 - t2)
        if advance:
            field.step += n_steps
        return

    lock_s, upload_s = upload_values(field, engine)
    t2 = clock()
    done = 0
    while done < n_steps:
        count = min(chunk, n_steps - done)
        records = engine.step(first_step + done, count, n_slots)
        if n_slots:
            _append_signals(layout, records)
        if progress_logger is not None:
            for s in range(first_step + done, first_step + done + count):
                progress_logger.log(s)
        done += count
    t3 = clock()

    download_values(field, engine)
    t4 = clock()
    field.__dict__['_last_run_profile'] = {
        'prepare_and_tables_s': t1 - t0, 'page_lock_s': lock_s, 'upload_state_s': upload_s,
        'step_s': t3 - t2,
        'download_state_s': t4 - t3}
    if advance:
        field.step += n_steps


def _slab_devices(field):
    """CUDA ordinals to spread a 2-D field over, or ``None`` for the single-GPU path: the field's
    ``devices`` attribute (a sequence of ordinals, or a count), else the environment variable
    ``FDS_DEVICES`` (a count). Needs slabs at least as tall as the halo the kernels consume."""
    if not hasattr(field, 'y'):
        return None
    wanted = getattr(field, 'devices', None)
    if wanted is None:
        wanted = os.environ.get('FDS_DEVICES')
        if wanted is None:
            return None
    if np.ndim(wanted) == 0:
        wanted = list(range(int(wanted)))
    devices = tuple(int(d) for d in wanted)
    if len(devices) < 2 or field.y.samples // len(devices) < MIN_SLAB_ROWS:
        return None      # slabs thinner than the edge bands of the kernels: one GPU does it
    return devices


def _log_throughput(field, engine, n_steps, seconds):
    """One INFO line per call on the reference's logger (``pyfds``): steps, wall time of the engine
    call (transfers included), cell updates per second and what that is in algorithmic bytes."""
    import logging
    logger = logging.getLogger('pyfds')
    if not logger.isEnabledFor(logging.INFO) or seconds <= 0:
        return
    launches, per_launch, kernel = engine.last_launch_info()
    rate = field.num_points * n_steps / seconds
    bytes_per_update = 16 * (1 if field._device_model.startswith('thermal') else engine.ncomp)
    logger.info('Device run: %d steps of %d cells in %.3f ms (%s, %d launches x %d steps): '
                '%.2f Gcell-updates/s, %.0f GB/s algorithmic, host arrays in and out.',
                n_steps, field.num_points, seconds * 1e3, kernel, launches, per_launch, rate / 1e9,
                rate * bytes_per_update / 1e9)


def reset(field):
    """Drops nothing, but makes sure stale device values can never leak into a reset field: the next
    ``run`` uploads the (zeroed) host values anyway."""
    state = field.__dict__.get('_engine_state')
    if state is not None and state.engine is not None:
        state.engine.reset_state()
    slabs = field.__dict__.get('_local_slabs')
    if slabs is not None:
        for engine in slabs.engines:
            engine.reset_state()
