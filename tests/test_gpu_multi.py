"""Multi-GPU slab runs (-m gpu, needs >= 2 GPUs; skipped on a single-GPU box): the slabs of 2 ranks,
stepped with NCCL halo exchange inside libfdsb200.so, must reproduce the single-domain reference
bit for bit -- fields and probe signals."""

import os
import socket

import numpy as np
import pytest

import pyfds_b200 as fds
import scenarios
from conftest import bits
from oracle import restate

pytestmark = pytest.mark.gpu


def _gpu_count():
    from pyfds_b200 import _engine
    try:
        return _engine.load_library().fds_device_count()
    except RuntimeError:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _build(name):
    if name == 'big_lossless':
        return scenarios._acoustic2d(fds, lossy=False, nx=512, ny=300, steps=37, seed=41)
    if name == 'big_lossy':
        return scenarios._acoustic2d(fds, lossy=True, nx=256, ny=130, steps=21, seed=42)
    if name == 'big_axi_lossy':      # viscous streaming kernel, 2 steps per launch, 4 halo rows
        return scenarios._acoustic2d(fds, lossy=True, nx=256, ny=140, steps=19, seed=43,
                                     klass='Acoustic3DAxi')
    if name == 'big_axi_lossless':
        return scenarios._acoustic2d(fds, lossy=False, nx=384, ny=90, steps=17, seed=44,
                                     klass='Acoustic3DAxi')
    return scenarios.SCENARIOS[name](fds)


def _worker(rank, world, port, name, kernel, queue):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from pyfds_b200 import parallel
        field, steps = _build(name)
        field.device_kernel = kernel
        run = parallel.SlabRun(field, rank, world, device=rank)
        first = steps // 3
        run.simulate(first)
        run.simulate(steps - first)
        run.gather()
        if rank == 0:
            queue.put(scenarios.collect(field))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('name,kernel', [
    ('big_lossless', 0), ('big_lossless', 1), ('big_lossy', 0), ('acoustic2d_wide', 0),
    ('acoustic2d_boundaries', 0), ('thermal2d', 0), ('acoustic3daxi_lossy', 0),
    ('acoustic_flow2d', 0), ('acoustic_flow2d_wide', 0), ('big_axi_lossy', 0),
    ('big_axi_lossless', 0)])
def test_two_slabs_equal_single_domain(library, name, kernel):
    if _gpu_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, kernel, queue))
             for r in range(world)]
    for p in procs:
        p.start()
    got = queue.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    field, steps = _build(name)
    stepper = restate.stepper_for(field).run(steps)
    expected = scenarios.collect_stepper(stepper)
    assert sorted(got) == sorted(expected)
    for key in expected:
        assert np.array_equal(bits(np.asarray(got[key])), bits(np.asarray(expected[key]))), (name, key)
