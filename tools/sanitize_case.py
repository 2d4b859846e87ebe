#!/usr/bin/env python
"""Small cases for compute-sanitizer: a 256 x 118 grid through the streaming kernels
(mbarrier rings, bulk copies, per-task flags) and, with `peer2`, two slabs on two GPUs exchanging halo
rows through peer memory. Every case is compared with the CPU restatement, so a sanitizer run is also
a parity run.

    compute-sanitizer --tool memcheck  python tools/sanitize_case.py stream2d streamv axi thermal line1d
    compute-sanitizer --tool racecheck python tools/sanitize_case.py stream2d streamv
    compute-sanitizer --tool memcheck --target-processes all python tools/sanitize_case.py peer2
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)

import pyfds_b200 as fds  # noqa: E402
import scenarios  # noqa: E402
from oracle import restate  # noqa: E402


def build(case):
    if case == 'stream2d':
        return scenarios._acoustic2d(fds, lossy=False, nx=256, ny=118, steps=14, seed=5)
    if case == 'streamv':
        return scenarios._acoustic2d(fds, lossy=True, nx=256, ny=118, steps=9, seed=6)
    if case == 'axi':
        return scenarios._acoustic2d(fds, lossy=True, nx=256, ny=118, steps=9, seed=7,
                                     klass='Acoustic3DAxi')
    if case == 'thermal':
        return scenarios._thermal2d(fds, 'Thermal2D', 256, 118, 14, seed=8)
    if case == 'line1d':
        return scenarios.acoustic1d_lossy(fds)
    if case == 'pipeline':       # fds_simulate cut into row bands (FDS_PIPELINE_FORCE is set below)
        return scenarios._acoustic2d(fds, lossy=False, nx=256, ny=420, steps=18, seed=9)
    raise SystemExit('unknown case ' + case)


def single(case):
    if case == 'pipeline':
        os.environ['FDS_PIPELINE_FORCE'] = '1'
        os.environ['FDS_PIPELINE_BANDS'] = '3'
    if case == 'coupled':
        return coupled()
    field, steps = build(case)
    field.simulate(steps // 2)
    field.simulate(steps - steps // 2)
    reference, _ = build(case)
    # the oracle runs on a fresh copy of the scenario
    ok = equal_fields(field, reference, steps)
    engine = field.__dict__['_engine_state'].engine
    print(case, 'kernel', engine.last_launch_info()[2], 'bitwise equal to the CPU restatement:', ok,
          flush=True)
    return ok


def coupled():
    """A ThermoAcoustic1D group on the device (pair kernel, chained launches) against its golden."""
    group, steps = scenarios.thermoacoustic1d_stepping(fds)
    group.simulate(steps)
    got = scenarios.collect_group(group)
    golden = np.load(os.path.join(ROOT, 'tests', 'golden', 'coupled_thermoacoustic1d_stepping.npz'))
    ok = all(np.array_equal(np.ascontiguousarray(got[k], dtype=np.float64).view(np.int64),
                            np.ascontiguousarray(golden[k], dtype=np.float64).view(np.int64))
             for k in golden.files if k != 'versions')
    print('coupled session', getattr(group, '_last_session', None),
          'bitwise equal to the reference golden:', ok, flush=True)
    return ok


def equal_fields(field, fresh, steps):
    stepper = restate.stepper_for(fresh).run(steps)
    got, expected = scenarios.collect(field), scenarios.collect_stepper(stepper)
    return all(np.array_equal(np.ascontiguousarray(got[k], dtype=np.float64).view(np.int64),
                              np.ascontiguousarray(expected[k], dtype=np.float64).view(np.int64))
               for k in expected)


def _peer_worker(rank, world, port, queue):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from pyfds_b200 import parallel
        field, steps = build('stream2d')
        run = parallel.SlabRun(field, rank, world, device=rank)
        run.simulate(steps // 2)
        run.simulate(steps - steps // 2)
        run.gather()
        if rank == 0:
            fresh, _ = build('stream2d')
            queue.put(equal_fields(field, fresh, steps))
    finally:
        dist.destroy_process_group()


def peer2():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    ok = queue.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
    print('peer2 (two slabs, in-kernel halo exchange) bitwise equal to the CPU restatement:', ok,
          flush=True)
    return ok


if __name__ == '__main__':
    results = [peer2() if case == 'peer2' else single(case) for case in sys.argv[1:]]
    sys.exit(0 if all(results) else 1)
