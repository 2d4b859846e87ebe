"""Host-side logic of the multi-GPU path, on CPU: slab partition, per-slab boundary/probe windows,
and -- with two gloo ranks -- the halo-exchange protocol (which rows, how many, how often), emulated
with the CPU restatement and NaN-poisoned non-halo rows (tests/slab_emulation.py)."""

import os
import socket
import sys

import numpy as np
import pytest

import pyfds_b200 as fds
import scenarios
from conftest import bits
from oracle import restate
from pyfds_b200 import _bake, parallel


def test_partition_rows_is_balanced_and_contiguous():
    for ny, world in [(4096, 8), (37, 4), (5, 5), (100, 3)]:
        parts = parallel.partition_rows(ny, world)
        assert parts[0][0] == 0 and sum(r for _, r in parts) == ny
        for (a0, a), (b0, _) in zip(parts, parts[1:]):
            assert a0 + a == b0
        assert max(r for _, r in parts) - min(r for _, r in parts) <= 1


def test_halo_rows():
    lossless, _ = scenarios.acoustic2d_wide(fds)
    lossy, _ = scenarios.acoustic2d_lossy(fds)
    thermal, _ = scenarios.thermal2d(fds)
    assert parallel.halo_rows_for(lossless, 1) == 0
    assert parallel.halo_rows_for(lossless, 4) == parallel.STREAM_STEPS     # streaming kernel
    assert parallel.halo_rows_for(lossless, 4, kernel=1) == 1
    assert parallel.halo_rows_for(lossy, 2) == 2                           # viscous 5-point operator
    assert parallel.halo_rows_for(thermal, 2) == 1
    # wide grids: every 2-D model on a streaming kernel, halo = reach x steps per launch
    lossy_wide, _ = scenarios.acoustic2d_lossy_wide(fds)
    axi_lossy, _ = scenarios.acoustic3daxi_lossy_wide(fds)
    axi_lossless, _ = scenarios.acoustic3daxi_lossless_wide(fds)
    thermal_axi, _ = scenarios.thermal3daxi_wide(fds)
    assert parallel.halo_rows_for(lossy_wide, 2) == 2 * parallel.STREAMV_STEPS
    assert parallel.halo_rows_for(axi_lossy, 8) == 2 * parallel.STREAMV_STEPS
    assert parallel.halo_rows_for(axi_lossless, 2) == parallel.STREAM_STEPS
    assert parallel.halo_rows_for(thermal_axi, 2) == parallel.STREAM_STEPS
    assert parallel.halo_rows_for(lossy_wide, 2, kernel=1) == 2


def test_slab_tables_cover_the_global_tables():
    """Boundary cells of all slabs (owned rows) = boundary cells of the whole grid; every probe slot
    is owned by exactly one slab."""
    field, steps = scenarios.acoustic2d_boundaries(fds)
    nx, ny = field.x.samples, field.y.samples
    whole = _bake.boundary_table(field.pressure.boundaries, 0, steps, 0, nx * ny, [])
    seen_cells, seen_slots = [], []
    total_slots = None
    for row0, rows in parallel.partition_rows(ny, 3):
        part = _bake.boundary_table(field.pressure.boundaries, 0, steps, row0 * nx,
                                    (row0 + rows) * nx, [])
        seen_cells.append(part.cells + row0 * nx)
        _, slots, total_slots = _bake.probe_table(field.pressure.outputs, 0, row0 * nx,
                                                  (row0 + rows) * nx)
        seen_slots.append(slots)
    assert np.array_equal(np.concatenate(seen_cells), whole.cells)
    assert sorted(np.concatenate(seen_slots)) == list(range(total_slots))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, steps_per_exchange, queue):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import slab_emulation
        field, steps = scenarios.SCENARIOS[name](fds)
        row0, rows, owned, _ = slab_emulation.run_slab(field, steps, rank, world, steps_per_exchange)
        queue.put((rank, row0, rows, {k: v for k, v in owned.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('name,steps_per_exchange', [
    ('acoustic2d_lossless', 1), ('acoustic2d_wide', 4), ('acoustic2d_lossy', 1),
    ('acoustic2d_boundaries', 2), ('thermal2d', 1), ('acoustic2d_lossy_wide', 2),
    ('acoustic3daxi_lossless_wide', 4), ('thermal3daxi_wide', 4), ('acoustic2d_signal_lines', 4)])
def test_two_rank_halo_protocol_matches_single_domain(name, steps_per_exchange):
    _check_halo_protocol(name, steps_per_exchange, 2)


# interior slabs (two neighbours) from 3 ranks on
@pytest.mark.parametrize('world', [3, 4])
@pytest.mark.parametrize('name,steps_per_exchange', [
    ('acoustic2d_wide', 4), ('acoustic2d_lossy_wide', 2), ('thermal2d', 1)])
def test_interior_slab_halo_protocol_matches_single_domain(name, steps_per_exchange, world):
    _check_halo_protocol(name, steps_per_exchange, world)


def _check_halo_protocol(name, steps_per_exchange, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, steps_per_exchange, queue))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    field, steps = scenarios.SCENARIOS[name](fds)
    reference = restate.stepper_for(field).run(steps)
    nx, ny = field.x.samples, field.y.samples
    for rank, row0, rows, owned in results:
        for comp, block in owned.items():
            expect = reference.values(comp).reshape(ny, nx)[row0:row0 + rows]
            assert np.array_equal(bits(block), bits(expect)), (name, rank, comp)
