#!/usr/bin/env python
"""Benchmark of the time-stepping hot path: Acoustic2D Gcell-updates/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size 4096]

* A *step* is one time step (``Acoustic2D.sim_step``, pyfds/acoustics.py:111-128) of the whole grid.
* N = 1 workload = BASELINE.json configs[1]: Acoustic2D 4096 x 4096 fp64, two material regions,
  additive point source, rigid line at x = 0, 4 Output probes. State (3 x 134 MB, two buffers) is far
  larger than the 126 MB L2, so successive steps cannot be served from cache.
* N > 1 (torchrun, one rank per GPU): weak scaling -- every rank owns a 4096-row y-slab of a
  4096 x (4096 N) grid; after every launch the outermost rows go straight into the neighbours' halo
  rows over NVLink peer memory (NCCL send/recv if there is no peer access). ``--strong`` cuts ONE
  size x size grid into slabs instead (BASELINE.json config 5: ``--size 32768 --strong``).
* ``value``: K steps with the state resident in HBM, timed with CUDA events on the engine's stream
  inside a barrier + device-synchronise bracket, max over ranks.
* ``e2e``: the same K steps through the public API (``field.simulate(K)``): host arrays in, host
  arrays and probe signals out, copies inside the timed region.
* ``roofline``: algorithmic bytes (48 B per cell update) per launch / measured launch
  time against the measured HBM copy peak in MEASURED_PEAKS.json.
* ``cpu_baseline`` / ``--impl reference``: the reference's CPU algorithm (scipy DIA mat-vec leapfrog,
  restated in oracle/restate.py with scipy's own dia_matvec doing the arithmetic) timed on this box.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

BYTES_PER_CELL_UPDATE = 48        # p, vx, vy: fp64 read + write
METRIC = 'Acoustic2D Gcell-updates/s'


def build_field(fds, nx, ny, t_samples, wall=True):
    """BASELINE.json configs[1] (SURVEY.md 8d 'C2 inputs'), scaled with the grid."""
    fld = fds.Acoustic2D(t_delta=1e-7, t_samples=t_samples, x_delta=1e-3, x_samples=nx,
                         y_delta=1e-3, y_samples=ny, material=fds.AcousticMaterial(1500, 1000))
    qx, qy = nx // 4, ny // 4
    fld.add_material_region(fld.get_rect_region((qx * 1e-3, qy * 1e-3, qx * 1e-3, qy * 1e-3)),
                            fds.AcousticMaterial(1200, 900))
    k = np.arange(t_samples)
    signal = np.sin(0.1 * k) * np.exp(-((k - 200) / 60) ** 2)
    fld.pressure.add_boundary(fld.get_point_region(((nx // 2) * 1e-3, (ny // 2) * 1e-3)),
                              value=signal, additive=True)
    if wall:
        fld.velocity_x.add_boundary(fld.get_line_region((0, 0, 0, (ny - 1) * 1e-3)))
    for m in range(1, 5):
        fld.pressure.add_output(fld.get_point_region(((m * nx // 8) * 1e-3, (m * ny // 8) * 1e-3)))
    return fld


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device = device
        self.lines = []          # (arrival time, csv line)
        self.proc = None

    def start(self, wait=8.0):
        """Launches nvidia-smi in loop mode and returns once its first sample has arrived (start-up
        takes longer than a short timed region, the more GPUs the longer)."""
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '25'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
            return
        deadline = time.perf_counter() + wait
        while not self.lines and time.perf_counter() < deadline:
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Clocks and throttle reasons of the samples that arrived in [t0, t1] (a sample describes
        the instant it was taken, just before its arrival); all samples if none fell inside."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.05)
        self.proc.terminate()
        rows = list(self.lines)
        inside = [line for at, line in rows if t0 is not None and t0 <= at <= t1 + 0.03]
        chosen = inside or [line for _, line in rows]
        sm, sm_max, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in chosen:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(sm_max) if sm_max else None,
                'reasons': sorted(reasons), 'samples': len(sm),
                'samples_in_timed_region': len(inside)}


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except (KeyError, ValueError):
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(kernel, nx, rows):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of the same kernel
    on the same per-GPU grid, if there is one (null otherwise: a capture of another grid says nothing
    about this one)."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(path):
        try:
            entry = json.load(open(path)).get(kernel)
        except ValueError:
            return None
        if isinstance(entry, dict):
            return entry.get('{}x{}'.format(nx, rows))
    return None


def cpu_reference_rate(nx, ny, steps):
    """Times the reference's CPU algorithm (scipy dia_matvec leapfrog) on the host: `steps` steps of
    the config-2 scenario on an nx x ny grid, after assembly. Returns (Gcell-updates/s, seconds)."""
    import pyfds_b200 as fds
    from oracle import restate
    field = build_field(fds, nx, ny, max(steps + 1, 8))
    rng = np.random.default_rng(0)
    for name in ('pressure', 'velocity_x', 'velocity_y'):
        getattr(field, name).values = 1e-3 * rng.standard_normal(nx * ny)
    stepper = restate.stepper_for(field, backend='scipy')
    stepper.run(1)                                   # touch all pages once
    t0 = time.perf_counter()
    stepper.run(steps)
    seconds = time.perf_counter() - t0
    return nx * ny * steps / seconds / 1e9, seconds


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on this box's host cores, rank 0 only."""
    if rank != 0:
        return
    # bounded sample per step: one time step of the config-2 scenario on a sub-grid chosen so that
    # warmup + K steps finish within a few minutes (the scipy path is single-threaded, ~55 ns/cell)
    total = args.steps + args.warmup
    size = 2048
    while size > 256 and total * size * size * 55e-9 > 150:
        size //= 2
    import pyfds_b200 as fds
    from oracle import restate
    field = build_field(fds, size, size, total + 1)
    rng = np.random.default_rng(0)
    for name in ('pressure', 'velocity_x', 'velocity_y'):
        getattr(field, name).values = 1e-3 * rng.standard_normal(size * size)
    stepper = restate.stepper_for(field, backend='scipy')
    stepper.run(args.warmup)
    t0 = time.perf_counter()
    stepper.run(args.steps)
    seconds = time.perf_counter() - t0
    value = size * size * args.steps / seconds / 1e9
    sample = ('Acoustic2D config-2 scenario on a {0}x{0} sub-grid, {1} steps after assembly; scipy '
              'dia_matvec + NumPy, single-threaded by construction').format(size, args.steps)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'Gcell-updates/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': seconds / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'Acoustic2D 4096x4096 fp64, 2 material regions, point source, '
                               '4 Output probes (sampled on a {0}x{0} sub-grid)'.format(size)},
        'cpu_baseline': {'value': value, 'unit': 'Gcell-updates/s', 'cores': 1, 'kind': 'port',
                         'sample': sample, 'host_cores': os.cpu_count()},
        'e2e': {'value': value, 'unit': 'Gcell-updates/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=2000)
    parser.add_argument('--warmup', type=int, default=100)
    parser.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    parser.add_argument('--size', type=int, default=4096, help='grid is size x size per GPU')
    parser.add_argument('--kernel', type=int, default=0, help='0 auto, 1 one-step, 2 streaming')
    parser.add_argument('--no-cpu-baseline', action='store_true')
    parser.add_argument('--no-e2e', action='store_true')
    parser.add_argument('--no-wall', action='store_true', help='experiment: drop the x=0 rigid line')
    parser.add_argument('--strong', action='store_true',
                        help='strong scaling: ONE size x size grid cut into y-slabs over the ranks '
                             '(BASELINE.json config 5 with --size 32768); default is weak scaling, '
                             'size x size per GPU')
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import pyfds_b200 as fds
    from pyfds_b200 import _engine, parallel

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    nx, rows = args.size, args.size
    ny = rows * world
    if args.strong:
        if args.size % world:
            raise SystemExit('--strong needs a size that is a multiple of the number of GPUs')
        rows, ny = args.size // world, args.size
    total_steps = args.warmup + args.steps
    field = build_field(fds, nx, ny, total_steps + 1, wall=not args.no_wall)
    field.assemble_matrices()

    # ---- device-resident run --------------------------------------------------------------------
    if world > 1:
        run = parallel.SlabRun(field, rank, world, device=local_rank, kernel=args.kernel)
        engine = run.engine
    else:
        engine = _engine.prepare(field, device=local_rank, kernel=args.kernel)
        run = None
    n_slots, _ = _engine.upload_run_tables(field, engine, 0, total_steps)
    rng = np.random.default_rng(1 + rank)
    for c in range(3):
        engine.upload_state(c, 1e-3 * rng.standard_normal(engine.owned))

    def barrier():
        engine.sync()
        if dist is not None:
            dist.barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    engine.step_async(0, args.warmup)
    barrier()
    t0 = time.perf_counter()
    engine.step_async(args.warmup, args.steps)
    engine.sync()
    t1 = time.perf_counter()
    wall = t1 - t0
    device_ms = engine.last_step_ms()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    barrier()
    launches, steps_per_launch, kernel = engine.last_launch_info()

    if dist is not None:
        import torch
        t = torch.tensor([device_ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        device_ms = float(t.item())
        # launches of the step kernel on all ranks (the halo wait / push kernels of the peer-memory
        # path, two more per launch and rank, are not counted)
        n = torch.tensor([launches], dtype=torch.int64, device='cuda')
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        total_launches = int(n.item())
    else:
        total_launches = launches

    cells = nx * ny
    per_gpu_cells = nx * rows
    value = cells * args.steps / (device_ms * 1e-3) / 1e9
    peak, peak_source = measured_peak()
    # time steps advance steps_per_launch at a time (a slab without peer access launches its edge
    # bands and its interior separately: `launches` counts those, not this)
    step_launches = -(-args.steps // max(steps_per_launch, 1))
    launch_ms = device_ms / max(step_launches, 1)
    achieved = BYTES_PER_CELL_UPDATE * per_gpu_cells * steps_per_launch / (launch_ms * 1e-3) / 1e9

    # ---- end to end through the public API (host arrays in and out) -----------------------------
    e2e = None
    if not args.no_e2e:
        api_field = build_field(fds, nx, ny, total_steps + 1, wall=not args.no_wall)
        for name in ('pressure', 'velocity_x', 'velocity_y'):
            values = getattr(api_field, name).values
            if world == 1:
                values[:] = 1e-3 * rng.standard_normal(cells)
            else:   # only the rows this rank owns are ever read
                values[rank * per_gpu_cells:(rank + 1) * per_gpu_cells] = \
                    1e-3 * rng.standard_normal(per_gpu_cells)
        api_field.device_kernel = args.kernel
        if world == 1:
            runner = api_field
        else:
            # drop the device-resident context first: one communicator / IPC mapping set per rank
            del run, engine
            runner = parallel.SlabRun(api_field, rank, world, device=local_rank, kernel=args.kernel)
        runner.simulate(args.warmup)             # includes assembly, context creation, first copies
        first_call = dict(api_field.__dict__.get('_last_run_profile') or {})
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        runner.simulate(args.steps)
        if dist is not None:
            dist.barrier()
        seconds = time.perf_counter() - t0
        state_bytes = 3 * per_gpu_cells * 8
        e2e = {'value': cells * args.steps / seconds / 1e9, 'unit': 'Gcell-updates/s',
               'h2d_bytes_per_step': world * state_bytes / args.steps,
               'd2h_bytes_per_step': world * (state_bytes + 4 * 8 * args.steps) / args.steps,
               'seconds': seconds, 'phases': api_field.__dict__.get('_last_run_profile'),
               'first_call_phases': first_call,
               'what': '{}.simulate({}) from host numpy arrays (page-locked on first use): upload of '
                       'p/vx/vy rows, boundary and probe tables, {} steps, download of p/vx/vy and '
                       'probe signals'.format('field' if world == 1 else 'SlabRun', args.steps,
                                              args.steps)}

    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        rate, seconds = cpu_reference_rate(2048, 2048, 20)
        cpu = {'value': rate, 'unit': 'Gcell-updates/s', 'cores': 1, 'kind': 'port',
               'sample': 'config-2 scenario on a 2048x2048 sub-grid, 20 steps after assembly '
                         '({:.1f} s); scipy dia_matvec + NumPy as in pyfds/acoustics.py:111-128, '
                         'single-threaded by construction'.format(seconds),
               'host_cores': os.cpu_count()}

    if rank == 0:
        traffic = ncu_traffic(kernel, nx, rows)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'Gcell-updates/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': device_ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if args.strong else 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {
                'workload': 'Acoustic2D {}x{} fp64 per GPU, 2 material regions, additive point '
                            'source, rigid line x=0, 4 Output probes'.format(nx, rows),
                'grid': [nx, ny], 'parallelism': 'y-slabs x{}'.format(world),
                'l2_policy': 'working set {:.0f} MB per GPU >> 126 MB L2'.format(
                    6 * per_gpu_cells * 8 / 1e6),
                'kernel': kernel, 'steps_per_launch': steps_per_launch},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic,
                         'peak_source': peak_source,
                         'bytes_per_cell_update': BYTES_PER_CELL_UPDATE,
                         'launch_ms': launch_ms,
                         # what the HBM actually carried: the committed ncu byte count of one launch
                         # over the launch time measured in this run (temporal blocking puts the
                         # algorithmic figure above the peak; this one cannot exceed it)
                         'dram_gbs': traffic / (launch_ms * 1e6) if traffic else None,
                         'dram_frac': traffic / (launch_ms * 1e6) / peak if traffic else None},
            'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': total_launches, 'clocks': clocks,
            'wall_seconds': wall,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
