"""Multi-GPU runs: one process per GPU, the grid cut into contiguous y-slabs.

The reference is single-process and has nothing comparable. The flat cell index
``x + y * nx`` (``pyfds/fields.py:377``) makes a y-slab a contiguous index range and every stencil
neighbour lies within +-nx of a cell, so a slab needs ``reach`` rows of its neighbours per time step:
1 row for the lossless leapfrog, 2 rows when the viscous 5-point operator is active. A launch that
advances k steps therefore consumes ``k * reach`` halo rows; after each launch the outermost rows of all
state components travel to the neighbour slabs (``ncclSend``/``ncclRecv`` pairs inside
``libfdsb200.so``, on a separate stream, while the interior rows are being computed).

``torch.distributed`` is only plumbing here: it hands the NCCL unique id from rank 0 to the other ranks
and gathers probe records; none of the field data moves through it.
"""

import numpy as np

from . import _bake, _engine

#: steps per launch of the streaming kernel = halo rows it needs on a lossless grid
STREAM_STEPS = 4
#: steps per launch of the viscous / axisymmetric streaming kernel (fds_streamv.cuh)
STREAMV_STEPS = 2


def partition_rows(ny, world):
    """Balanced contiguous row ranges: list of ``(row0, rows)`` for ranks 0..world-1."""
    base, extra = divmod(int(ny), int(world))
    parts, row0 = [], 0
    for rank in range(world):
        rows = base + (1 if rank < extra else 0)
        parts.append((row0, rows))
        row0 += rows
    return parts


def is_lossy(field):
    """True if any material region has a non-zero absorption coefficient (acoustic models)."""
    if not field.matrices_assembled:
        field.assemble_matrices()
    snapshot = field._baked['snapshot']
    if isinstance(snapshot, _bake.DenseSnapshot):
        return bool(np.any(snapshot.vectors.get('absorption_coef', 0) != 0))
    return any(values.get('absorption_coef', 0) != 0 for _, values in snapshot.entries)


def stencil_reach(field):
    """Rows of neighbour data one time step needs on either side of a slab."""
    return 2 if is_lossy(field) else 1


def streaming_steps(field):
    """Time steps per launch of the streaming kernels for this field (0: not eligible): 4 for the
    lossless acoustic and the thermal 2-D models (plain and axisymmetric), 2 for the lossy ones."""
    nx = field.x.samples
    if nx % 4 or nx < 128:
        return 0
    model = field._device_model
    if model in ('thermal2d', 'thermal3daxi'):
        return STREAM_STEPS
    if model in ('acoustic2d', 'acoustic3daxi') and not is_lossy(field):
        return STREAM_STEPS
    if model in ('acoustic2d', 'acoustic3daxi'):
        return STREAMV_STEPS
    return 0


def streaming_eligible(field):
    return streaming_steps(field) > 0


def halo_rows_for(field, world, kernel=0):
    """Halo rows per side of a slab context: the reach of one step, times the steps per launch when
    a streaming kernel will run."""
    if world <= 1:
        return 0
    if kernel in (0, 2) and streaming_eligible(field):
        return streaming_steps(field) * stencil_reach(field)
    return stencil_reach(field)


def _broadcast_unique_id(rank):
    import torch.distributed as dist
    payload = [_engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    return payload[0]


class SlabRun:
    """The slab of ``field`` owned by ``rank``: device context, NCCL communicator and a ``simulate``
    that mirrors ``Field.simulate`` for the rows this rank owns.

    Every rank constructs the same ``field`` (same script, SPMD). Host ``values`` arrays stay full
    size; a rank uploads and refreshes only its own rows -- use ``gather`` to assemble whole fields.
    """

    def __init__(self, field, rank, world, device=0, kernel=None, unique_id=None):
        if not hasattr(field, 'y'):
            raise ValueError('Only 2-D fields are partitioned; 1-D problems stay on one GPU.')
        if not field.matrices_assembled:
            field.assemble_matrices()
        if kernel is None:
            kernel = getattr(field, 'device_kernel', 0)
        self.field, self.rank, self.world = field, rank, world
        self.row0, self.rows = partition_rows(field.y.samples, world)[rank]
        self.halo_rows = halo_rows_for(field, world, kernel)
        if world > 1 and self.rows < self.halo_rows:
            raise ValueError('Slabs of {} rows are thinner than the {} halo rows needed.'.format(
                self.rows, self.halo_rows))
        # decided on the whole grid: a lossy material confined to one slab must not make the slabs
        # pick different kernels (they would exchange halo rows at different cadences)
        lossy = field._device_model.startswith('acoustic') and is_lossy(field)
        self.engine = _engine.prepare(field, device=device, row0=self.row0, rows=self.rows,
                                      halo_rows=self.halo_rows, kernel=kernel, lossy=lossy)
        if world > 1:
            if unique_id is None:
                unique_id = _broadcast_unique_id(rank)
            self.engine.comm_init(unique_id, rank, world)
            self._open_peers()

    def _open_peers(self):
        """Maps the neighbour slabs' state buffers (CUDA IPC) so that halo rows are copied straight
        into the neighbours' memory over NVLink by a small kernel after every launch, ordered by flags
        in peer memory. Best effort: without peer access the rows travel by ncclSend/ncclRecv."""
        import os
        import torch.distributed as dist
        if os.environ.get('FDS_NO_PEER'):
            return
        handles = [None] * self.world
        dist.all_gather_object(handles, self.engine.peer_export())
        parts = partition_rows(self.field.y.samples, self.world)
        try:
            if self.rank > 0:
                self.engine.peer_import(0, handles[self.rank - 1], parts[self.rank - 1][1])
            if self.rank < self.world - 1:
                self.engine.peer_import(1, handles[self.rank + 1], parts[self.rank + 1][1])
        except _engine.EngineError:
            pass

    @property
    def cells(self):
        nx = self.field.x.samples
        return slice(self.row0 * nx, (self.row0 + self.rows) * nx)

    def upload_state(self):
        state = self.field.__dict__['_engine_state']
        for c, name in enumerate(self.field._device_components):
            values = _engine._host_values(getattr(self.field, name), self.field.num_points)
            own = values[self.cells]
            if values.ctypes.data == np.asarray(getattr(self.field, name).values).ctypes.data:
                state.pin(self.engine.lib, c, own)       # page-lock the rows this rank owns
            self.engine.upload_state(c, own)

    def download_state(self):
        for c, name in enumerate(self.field._device_components):
            component = getattr(self.field, name)
            values = np.ascontiguousarray(component.values, dtype=np.float64)
            self.engine.download_state(c, out=values[self.cells])
            component.values = values

    def simulate(self, num_steps, gather_probes=True):
        """Advances the slab ``num_steps`` steps. Probe signals are summed over ranks (every probe
        point is owned by exactly one slab) so that each rank ends up with complete ``signals``."""
        import time
        clock = time.perf_counter
        field, engine = self.field, self.engine
        first_step = field.step
        t0 = clock()
        n_slots, layout = _engine.upload_run_tables(field, engine, first_step, num_steps)
        t1 = clock()
        self.upload_state()
        t2 = clock()
        records = engine.step(first_step, num_steps, n_slots)
        t3 = clock()
        self.download_state()
        t4 = clock()
        field.__dict__['_last_run_profile'] = {
            'prepare_and_tables_s': t1 - t0, 'upload_state_s': t2 - t1, 'step_s': t3 - t2,
            'download_state_s': t4 - t3}
        if n_slots:
            if gather_probes and self.world > 1:
                import torch
                import torch.distributed as dist
                # gloo/NCCL agnostic: go through a tensor on the backend's device. The records are
                # summed as BIT PATTERNS (int64): every slot is zero bits on all ranks but its owner,
                # so the sum is the owner's sample bit for bit -- a float sum would turn -0.0 into +0.0
                backend = dist.get_backend()
                tensor = torch.from_numpy(np.ascontiguousarray(records).view(np.int64).copy())
                if backend == 'nccl':
                    tensor = tensor.cuda()
                dist.all_reduce(tensor)
                records = tensor.cpu().numpy().view(np.float64)
            _engine._append_signals(layout, records)
        field.__dict__['_last_run_profile']['probe_gather_s'] = clock() - t4
        field.step += num_steps

    def gather(self):
        """All-gathers the owned rows so that every rank holds the complete ``values`` arrays."""
        if self.world == 1:
            return
        import torch
        import torch.distributed as dist
        nx = self.field.x.samples
        parts = partition_rows(self.field.y.samples, self.world)
        for name in self.field._device_components:
            component = getattr(self.field, name)
            values = np.ascontiguousarray(component.values, dtype=np.float64)
            for src, (row0, rows) in enumerate(parts):
                block = torch.from_numpy(values[row0 * nx:(row0 + rows) * nx].copy())
                if dist.get_backend() == 'nccl':
                    block = block.cuda()
                dist.broadcast(block, src=src)
                values[row0 * nx:(row0 + rows) * nx] = block.cpu().numpy()
            component.values = values


class LocalSlabs:
    """All slabs of ``field`` in THIS process: one device context per GPU, wired to its neighbours
    directly (``fds_slab_init`` / ``fds_peer_connect``: peer access, no IPC, no NCCL, no
    ``torch.distributed``), and one host thread per slab for the duration of a call -- the slabs' step
    calls wait for each other on the device, so they must be in flight together. This is what
    ``field.simulate()`` runs on when the field has a ``devices`` attribute (or ``FDS_DEVICES`` is set):
    the reference's call, all GPUs of the box, host ``values`` and ``signals`` complete afterwards."""

    def __init__(self, field, devices, kernel=None):
        if not hasattr(field, 'y'):
            raise ValueError('Only 2-D fields are partitioned; 1-D problems stay on one GPU.')
        self.field = field
        self.devices = tuple(devices)
        self.kernel = getattr(field, 'device_kernel', 0) if kernel is None else kernel
        self.states = [_engine._State() for _ in self.devices]
        self.engines = []

    def close(self):
        for state in self.states:
            if state.engine is not None:
                state.engine.close()
                state.engine = None
        self.engines = []

    def _prepare(self):
        field, world = self.field, len(self.devices)
        if not field.matrices_assembled:
            field.assemble_matrices()
        parts = partition_rows(field.y.samples, world)
        halo = halo_rows_for(field, world, self.kernel)
        if min(rows for _, rows in parts) < halo:
            raise ValueError('Slabs of {} rows are thinner than the {} halo rows needed.'.format(
                min(rows for _, rows in parts), halo))
        lossy = field._device_model.startswith('acoustic') and is_lossy(field)
        engines = [_engine.prepare(field, device=device, row0=row0, rows=rows, halo_rows=halo,
                                   kernel=self.kernel, lossy=lossy, state=state)
                   for device, (row0, rows), state in zip(self.devices, parts, self.states)]
        if len(engines) != len(self.engines) or \
                any(a is not b for a, b in zip(engines, self.engines)):
            # at least one context is new: wire all of them afresh
            if self.engines and any(a is b for a, b in zip(engines, self.engines)):
                self.close()
                return self._prepare()
            for rank, engine in enumerate(engines):
                engine.slab_init(rank, world)
            for rank, engine in enumerate(engines):
                if rank > 0:
                    engine.peer_connect(0, engines[rank - 1])
                if rank < world - 1:
                    engine.peer_connect(1, engines[rank + 1])
            self.engines = engines
        return parts

    def simulate(self, num_steps, progress_logger=None):
        """``num_steps`` x ``sim_step`` on all slabs (does not advance ``field.step``: the caller
        does, as for the single-GPU path)."""
        import threading
        field = self.field
        parts = self._prepare()
        nx = field.x.samples
        first_step = field.step
        arrays = []
        for name in field._device_components:
            component = getattr(field, name)
            values = _engine._host_values(component, field.num_points)
            own = component.values
            in_place = isinstance(own, np.ndarray) and own.ctypes.data == values.ctypes.data and \
                own.nbytes == values.nbytes and values.flags.writeable
            if not in_place:      # a list, another dtype, a read-only view: work on a copy
                values = np.array(values, dtype=np.float64)
                component.values = values
            arrays.append(values)
        layouts = [_engine.upload_run_tables(field, engine, first_step, num_steps)
                   for engine in self.engines]
        n_slots, layout = layouts[0]
        uploaded = threading.Barrier(len(self.engines))
        records, errors = [None] * len(self.engines), []

        def work(rank):
            engine, state = self.engines[rank], self.states[rank]
            row0, rows = parts[rank]
            cells = slice(row0 * nx, (row0 + rows) * nx)
            try:
                try:
                    for c, values in enumerate(arrays):
                        state.pin(engine.lib, c, values[cells])
                        engine.upload_state(c, values[cells])
                    # nothing may be allocated once a neighbour is stepping (it may already be waiting
                    # for this slab on the device, and with peer access an allocation synchronises
                    # with the peers): do it all now
                    engine.step_prepare(first_step, num_steps)
                finally:
                    uploaded.wait()          # a slab pulls its halos from the neighbours' uploads
                records[rank] = engine.step(first_step, num_steps, n_slots)
                for c, values in enumerate(arrays):
                    engine.download_state(c, out=values[cells])
            except Exception as exc:      # noqa: BLE001  (re-raised on the calling thread)
                errors.append(exc)
                uploaded.abort()

        threads = [threading.Thread(target=work, args=(rank,)) for rank in range(len(self.engines))]
        for thread in threads:
            thread.start()
        for thread in threads:
            thread.join()
        if errors:
            raise errors[0]
        if n_slots:
            # every probe point is owned by exactly one slab: its record comes from that slab (picked,
            # not summed -- a sum would turn a recorded -0.0 into +0.0)
            cells = np.concatenate([np.asarray(output.region.indices, dtype=np.int64).reshape(-1)
                                    for output, _, _ in layout])
            first_cells = np.array([row0 * nx for row0, _ in parts], dtype=np.int64)
            owner = np.searchsorted(first_cells, cells, side='right') - 1
            merged = np.empty_like(records[0])
            for rank, block in enumerate(records):
                mine = owner == rank
                merged[:, mine] = block[:, mine]
            _engine._append_signals(layout, merged)
        if progress_logger is not None:
            for s in range(first_step, first_step + num_steps):
                progress_logger.log(s)
