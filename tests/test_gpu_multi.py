"""Multi-GPU slab runs (-m gpu, needs >= 2 GPUs; skipped on a single-GPU box): the slabs of 2, 3, 4
and 8 ranks -- from 3 on there are interior slabs with two neighbours -- stepped with the halo exchange
inside libfdsb200.so (rows pushed into the neighbours' memory by the step kernel itself; the kernel
pair and the NCCL send/recv path as variants) must reproduce the single-domain CPU restatement bit for
bit: fields and probe signals."""

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
    if name == 'flow_monotone':
        # every row its own shift period, none shared between slabs: all slabs must still end their
        # launches at the same steps (the schedule comes from the periods of the whole grid)
        return scenarios._acoustic_flow2d(fds, 256, 96, 45, seed=45,
                                          periods=tuple(range(3, 3 + 96)))
    if name == 'lossy_in_one_slab':
        return _lossy_corner()
    if name == 'seams':
        return _seams()
    return scenarios.SCENARIOS[name](fds)


def _lossy_corner():
    """A lossy material confined to the topmost rows: every slab must run the viscous kernel (and its
    halo cadence) although only the last one holds a lossy cell."""
    nx, ny, steps = 256, 128, 22
    fld = fds.Acoustic2D(t_delta=1e-7, t_samples=steps, x_delta=1e-3, x_samples=nx, y_delta=1e-3,
                         y_samples=ny, material=fds.AcousticMaterial(1500, 1000))
    fld.add_material_region(fld.get_rect_region((40e-3, (ny - 9) * 1e-3, 100e-3, 6e-3)),
                            fds.AcousticMaterial(1400, 950, absorption_coef=7.7))
    scenarios._randomise(fld, ('pressure', 'velocity_x', 'velocity_y'), seed=46)
    fld.pressure.add_output(fld.get_point_region((50e-3, (ny - 5) * 1e-3)))
    fld.pressure.add_output(fld.get_point_region((50e-3, 5e-3)))
    return fld, steps


def _seams(world=8):
    """Sources, walls and probes exactly on the first and last rows of the slabs of 2, 4 and 8 ranks."""
    nx, ny, steps = 384, 256, 30
    fld = fds.Acoustic2D(t_delta=1e-7, t_samples=steps, x_delta=1e-3, x_samples=nx, y_delta=1e-3,
                         y_samples=ny, material=fds.AcousticMaterial(1500, 1000))
    fld.add_material_region(fld.get_rect_region((90e-3, 20e-3, 120e-3, 200e-3)),
                            fds.AcousticMaterial(1200, 900))
    scenarios._randomise(fld, ('pressure', 'velocity_x', 'velocity_y'), seed=47)
    k = np.arange(steps)
    for g in range(1, world):
        for row in (g * ny // world - 1, g * ny // world):
            x = 30 + 37 * g
            fld.pressure.add_boundary(fld.get_point_region((x * 1e-3, row * 1e-3)),
                                      value=np.cos(0.3 * k + g), additive=True)
            fld.velocity_y.add_boundary(fld.get_line_region(((x + 10) * 1e-3, row * 1e-3,
                                                             (x + 25) * 1e-3, row * 1e-3)))
            fld.pressure.add_output(fld.get_point_region((x * 1e-3, row * 1e-3)))
            fld.velocity_y.add_output(fld.get_line_region(((x + 8) * 1e-3, row * 1e-3,
                                                           (x + 12) * 1e-3, row * 1e-3)))
    fld.velocity_x.add_boundary(fld.get_line_region((0, 0, 0, (ny - 1) * 1e-3)))
    return fld, steps


def _worker(rank, world, port, name, kernel, env, queue):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    os.environ.update(env)
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


def _run_slabs(name, kernel, world, env=None):
    if _gpu_count() < world:
        pytest.skip('needs {} GPUs'.format(world))
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, kernel, env or {}, queue))
             for r in range(world)]
    for p in procs:
        p.start()
    try:
        got = queue.get(timeout=300)
    finally:
        for p in procs:
            p.join(timeout=120)
            if p.is_alive():
                p.terminate()
    for p in procs:
        assert p.exitcode == 0

    field, steps = _build(name)
    stepper = restate.stepper_for(field).run(steps)
    expected = scenarios.collect_stepper(stepper)
    assert sorted(got) == sorted(expected)
    for key in expected:
        assert np.array_equal(bits(np.asarray(got[key])), bits(np.asarray(expected[key]))), \
            (name, world, key)


@pytest.mark.parametrize('name,kernel', [
    ('big_lossless', 0), ('big_lossless', 1), ('big_lossy', 0), ('acoustic2d_wide', 0),
    ('acoustic2d_boundaries', 0), ('thermal2d', 0), ('acoustic3daxi_lossy', 0),
    ('acoustic_flow2d', 0), ('acoustic_flow2d_wide', 0), ('big_axi_lossy', 0),
    ('big_axi_lossless', 0), ('flow_monotone', 0), ('lossy_in_one_slab', 0), ('seams', 0),
    ('thermal2d_wide', 0)])
def test_two_slabs_equal_single_domain(library, name, kernel):
    _run_slabs(name, kernel, 2)


# Interior slabs have two neighbours: both flags, both edge bands, pushes in both directions.
@pytest.mark.parametrize('world', [3, 4, 8])
@pytest.mark.parametrize('name,kernel', [
    ('big_lossless', 0), ('big_lossless', 1), ('big_lossy', 0), ('seams', 0), ('flow_monotone', 0),
    ('big_axi_lossy', 0), ('lossy_in_one_slab', 0), ('thermal2d_wide', 0)])
def test_many_slabs_equal_single_domain(library, name, kernel, world):
    _run_slabs(name, kernel, world)


# The other two transports of the same rows: a waiting and a pushing kernel either side of the sweep
# (no in-kernel exchange), and NCCL send/recv (no peer memory at all); and sweeps that do not overlap.
@pytest.mark.parametrize('env', [{'FDS_HALO_KERNELS': '1'}, {'FDS_NO_PEER': '1'},
                                 {'FDS_NO_OVERLAP': '1'}])
@pytest.mark.parametrize('name', ['big_lossless', 'big_lossy', 'flow_monotone'])
def test_other_halo_transports(library, name, env):
    _run_slabs(name, 0, 2, env)
    if _gpu_count() >= 4:
        _run_slabs(name, 0, 4, env)


# ---- the same slabs in ONE process: field.simulate() with a `devices` attribute ------------------------
# (contexts wired with fds_slab_init / fds_peer_connect, one host thread per slab; no torch.distributed)

@pytest.mark.parametrize('world', [2, 3, 4, 8])
@pytest.mark.parametrize('name', ['big_lossless', 'big_lossy', 'seams', 'flow_monotone',
                                  'thermal2d_wide', 'big_axi_lossy', 'lossy_in_one_slab'])
def test_simulate_on_several_devices_of_one_process_equals_single_domain(library, name, world):
    if _gpu_count() < world:
        pytest.skip('needs {} GPUs'.format(world))
    field, steps = _build(name)
    field.devices = list(range(world))
    first = steps // 3
    field.simulate(first)                       # the reference's call; all slabs behind it
    field.simulate(steps - first)
    slabs = field.__dict__.get('_local_slabs')
    from pyfds_b200 import _engine
    if field.y.samples // world >= _engine.MIN_SLAB_ROWS:
        assert slabs is not None and len(slabs.engines) == world, 'the run did not use the devices'
    else:       # slabs would be thinner than the kernels' edge bands: one GPU does the whole field
        assert slabs is None
    got = scenarios.collect(field)
    if slabs is not None:
        slabs.close()

    fresh, _ = _build(name)
    expected = scenarios.collect_stepper(restate.stepper_for(fresh).run(steps))
    assert sorted(got) == sorted(expected)
    for key in expected:
        assert np.array_equal(bits(np.asarray(got[key])), bits(np.asarray(expected[key]))), \
            (name, world, key)
