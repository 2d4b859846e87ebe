#!/usr/bin/env python
"""Benchmark of the time-stepping hot path: Acoustic2D Gcell-updates/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* A *step* is one time step (``Acoustic2D.sim_step``, pyfds/acoustics.py:111-128) of the whole grid.
* ``value`` -- N = 1: BASELINE.json configs[1], Acoustic2D 4096 x 4096 fp64, two material regions,
  additive point source, rigid line at x = 0, 4 Output probes. N > 1 (torchrun, one rank per GPU):
  weak scaling with the same 4096 x 4096 cells per GPU, the grid being 4096 x (4096 N) cut into
  y-slabs; halo rows travel over NVLink peer memory from inside the step kernel. The K steps are timed
  ``--repeats`` (default 5) times back to back with CUDA events on the engine's stream, each repeat
  inside a barrier + device-synchronise bracket, max over ranks; ``value`` is the MEDIAN repeat
  (min / max are in ``repeats``). State is far larger than the 126 MB L2.
* ``target`` (N = 1) -- the north-star size, Acoustic2D 16384 x 16384, measured the same way in the
  same run, with its own roofline record.
* ``config5`` -- BASELINE.json configs[4]: ``weak`` = 32768 x 4096 cells per GPU (a 32768 x 4096 N
  grid), ``strong`` = ONE 32768 x 32768 grid cut into N slabs (on one GPU as well: 51.5 GB).
* ``other_models`` (N = 1) -- the remaining single-GPU entries of BASELINE.json ``configs``, device
  resident, through ``benchmarks/configs.py``: configs[2] (Acoustic3DAxi 8192 x 4096 with lossy
  regions), configs[3] (Thermal2D 8192 x 8192), and the lossy twin of configs[1]. Parity-test
  cases, not bench lines: they are reported so that their kernels' rates are driver-run numbers.
* ``parity`` (N > 1) -- a 4096 x (512 N) grid stepped as N slabs and, on every rank, as one grid on
  that rank's GPU; true if every rank's rows and all probe records are bitwise equal.
* ``e2e`` -- the same K steps through the public API (``field.simulate(K)``): host arrays in, host
  arrays and probe signals out, copies inside the timed region.
* ``roofline`` -- algorithmic bytes (48 B per cell update) per launch / measured launch
  time against the measured HBM copy peak in MEASURED_PEAKS.json (``frac``), next to what the HBM
  actually carried (``dram_frac``: the committed ncu byte count of one launch over the launch time
  measured here -- with K steps per launch the algorithmic figure exceeds the peak by design).
* ``cpu_baseline`` / ``--impl reference`` -- the reference itself (the unmodified pyfds package staged
  in the git-ignored ``baseline/_ref`` by ``__graft_entry__.build()``) stepping the same 4096 x 4096
  scenario on this box's host cores; the CPU restatement in ``oracle/`` only if that copy is absent.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

BYTES_PER_CELL_UPDATE = 48        # p, vx, vy: fp64 read + write
METRIC = 'Acoustic2D Gcell-updates/s'
REFERENCE_DIR = os.path.join(ROOT, 'baseline', '_ref')


def build_field(fds, nx, ny, t_samples, wall=True, seam_rows=None):
    """BASELINE.json configs[1] (SURVEY.md 8d 'C2 inputs'), scaled with the grid. ``fds`` is the
    package under test: ``pyfds_b200`` or the reference ``pyfds`` (same public API).

    ``seam_rows`` (parity check): additionally an additive source, a probe line and a wall segment on
    the given rows -- the first and last rows of interior slabs."""
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
    for n, row in enumerate(seam_rows or ()):
        x = (nx // 3 + 17 * n) % nx
        fld.pressure.add_boundary(fld.get_point_region((x * 1e-3, row * 1e-3)),
                                  value=np.cos(0.05 * k + n), additive=True)
        fld.velocity_y.add_boundary(fld.get_line_region(((x + 40) * 1e-3, row * 1e-3,
                                                         (x + 90) * 1e-3, row * 1e-3)))
        fld.velocity_y.add_output(fld.get_line_region(((x + 30) * 1e-3, row * 1e-3,
                                                       (x + 36) * 1e-3, row * 1e-3)))
        fld.pressure.add_output(fld.get_point_region((x * 1e-3, row * 1e-3)))
    return fld


def workload_name(nx, rows):
    return ('Acoustic2D {}x{} fp64 per GPU, 2 material regions, additive point source, rigid line '
            'x=0, 4 Output probes'.format(nx, rows))


def config_record(nx, rows, ny, world, strong=False):
    """Identical for the product arm and the reference arm."""
    return {'workload': workload_name(nx, rows), 'grid': [nx, ny],
            'parallelism': 'y-slabs x{}'.format(world),
            'l2_policy': 'working set {:.0f} MB per GPU >> 126 MB L2, no flush needed'.format(
                6 * nx * rows * 8 / 1e6)}


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi for the whole run; ``window`` summarises
    the samples that arrived during one timed region."""

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
                 '--format=csv,noheader,nounits', '-lms', '20'],
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

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def window(self, t0, t1):
        """Clocks and throttle reasons of the samples that arrived in [t0, t1] (a sample describes
        the instant it was taken, just before its arrival); the nearest samples if none fell inside."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.03)
        rows = list(self.lines)
        inside = [line for at, line in rows if t0 <= at <= t1 + 0.03]
        chosen = inside
        if not chosen and rows:
            nearest = sorted(rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))[:2]
            chosen = [line for _, line in nearest]
        sm, sm_max, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in chosen:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(sm_max) if sm_max else None,
                'power_w': max(power) if power else None,
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
    about this one). A constant read from profiles/traffic.json, NOT a counter of this run."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(path):
        try:
            entry = json.load(open(path)).get(kernel)
        except ValueError:
            return None
        if isinstance(entry, dict):
            return entry.get('{}x{}'.format(nx, rows))
    return None


# ---------------------------------------------------------------------------------------------------
# the reference on the host
# ---------------------------------------------------------------------------------------------------

def import_reference():
    """The unmodified reference package from baseline/_ref (staged by __graft_entry__.build()); None
    if it is not there. ``import pyfds`` pulls in matplotlib through pyfds/gfx.py, which nothing on the
    time-stepping path uses: empty stand-in modules are registered first."""
    if not os.path.isdir(os.path.join(REFERENCE_DIR, 'pyfds')):
        return None
    for name in ('matplotlib', 'matplotlib.patches', 'matplotlib.pyplot', 'matplotlib.animation'):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import warnings
    warnings.simplefilter('ignore', DeprecationWarning)
    import pyfds
    if not os.path.abspath(pyfds.__file__).startswith(os.path.abspath(REFERENCE_DIR)):
        return None
    return pyfds


def cpu_reference_rate(nx, ny, warmup, steps):
    """Times the reference's CPU path on this host: ``steps`` steps of the config-2 scenario on an
    nx x ny grid after assembly and ``warmup`` steps, random initial state.
    Returns ``(Gcell-updates/s, seconds, kind, setup seconds)``."""
    rng = np.random.default_rng(0)
    t_setup = time.perf_counter()
    pyfds = import_reference()
    if pyfds is not None:
        kind = 'reference'
        field = build_field(pyfds, nx, ny, warmup + steps + 1)
        for name in ('pressure', 'velocity_x', 'velocity_y'):
            getattr(field, name).values = 1e-3 * rng.standard_normal(nx * ny)
        field.assemble_matrices()
        run = field.simulate
    else:
        kind = 'port'
        import pyfds_b200 as fds
        from oracle import restate
        field = build_field(fds, nx, ny, warmup + steps + 1)
        for name in ('pressure', 'velocity_x', 'velocity_y'):
            getattr(field, name).values = 1e-3 * rng.standard_normal(nx * ny)
        run = restate.stepper_for(field, backend='scipy').run
    run(max(warmup, 1))
    t0 = time.perf_counter()
    run(steps)
    seconds = time.perf_counter() - t0
    return nx * ny * steps / seconds / 1e9, seconds, kind, t0 - t_setup


def reference_sample_text(kind, nx, ny, steps, seconds, setup):
    what = ('the unmodified reference package (baseline/_ref/pyfds, Acoustic2D.simulate: scipy '
            'dia_matrix.dot + NumPy, pyfds/acoustics.py:111-128)' if kind == 'reference' else
            'the CPU restatement oracle/restate.py with scipy\'s own dia_matvec (baseline/_ref absent)')
    return ('config-2 scenario on the full {}x{} grid, {} steps in {:.1f} s after {:.1f} s of '
            'construction + assembly + warm-up; {}; single-threaded by construction'.format(
                nx, ny, steps, seconds, setup, what))


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on this box's host cores, rank 0 only. Same config
    as the product arm; one step = one time step of one GPU's 4096 x 4096 share of it."""
    if rank != 0:
        return
    nx = rows = args.size
    value, seconds, kind, setup = cpu_reference_rate(nx, rows, args.warmup, args.steps)
    sample = reference_sample_text(kind, nx, rows, args.steps, seconds, setup)
    if world > 1:
        sample += '; the sample is ONE slab ({}x{}) of the {}-slab grid'.format(nx, rows, world)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'Gcell-updates/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': seconds / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config_record(nx, rows, rows * world, world),
        'cpu_baseline': {'value': value, 'unit': 'Gcell-updates/s', 'cores': 1, 'kind': kind,
                         'sample': sample, 'host_cores': os.cpu_count()},
        'e2e': {'value': value, 'unit': 'Gcell-updates/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# the product arm
# ---------------------------------------------------------------------------------------------------

class Job:
    """Rank, world, the torch.distributed handle (None on one GPU) and the clock sampler."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.dist = None
        self.torch = None
        self.sampler = None

    def init(self):
        if self.world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank))
            self.dist, self.torch = dist, torch
        if self.rank == 0:
            self.sampler = ClockSampler(self.local_rank)
            self.sampler.start()

    def barrier(self, engine=None):
        if engine is not None:
            engine.sync()
        if self.dist is not None:
            self.dist.barrier()

    def reduce(self, value, op='max', dtype='float64'):
        if self.dist is None:
            return value
        t = self.torch.tensor([value], dtype=getattr(self.torch, dtype), device='cuda')
        self.dist.all_reduce(t, op={'max': self.dist.ReduceOp.MAX, 'min': self.dist.ReduceOp.MIN,
                                    'sum': self.dist.ReduceOp.SUM}[op])
        return t.item()

    def close(self):
        if self.sampler is not None:
            self.sampler.stop()
        if self.dist is not None:
            self.dist.destroy_process_group()


def random_state(rng, n):
    """n doubles of noise at 1e-3: a 4 Mi-sample block repeated (generating 10^9 normal deviates on the
    host would take longer than everything else the benchmark does)."""
    block = 1e-3 * rng.standard_normal(min(n, 1 << 22))
    return block if block.shape[0] == n else np.resize(block, n)


def measure_resident(job, nx, rows, ny, strong=False, repeats=None):
    """K steps with the state resident in HBM, ``repeats`` times; returns the record of the median."""
    import pyfds_b200 as fds
    from pyfds_b200 import _engine, parallel
    args = job.args
    repeats = repeats or args.repeats
    total_steps = args.warmup + repeats * args.steps
    field = build_field(fds, nx, ny, total_steps + 1, wall=not args.no_wall)
    field.assemble_matrices()
    if job.world > 1:
        run = parallel.SlabRun(field, job.rank, job.world, device=job.local_rank, kernel=args.kernel)
        engine = run.engine
    else:
        run = None
        engine = _engine.prepare(field, device=job.local_rank, kernel=args.kernel)
    try:
        _engine.upload_run_tables(field, engine, 0, total_steps)
        rng = np.random.default_rng(1 + job.rank)
        for c in range(3):
            engine.upload_state(c, random_state(rng, engine.owned))
        engine.step_async(0, args.warmup)
        times, t_first, t_last, launches = [], None, None, 0
        for r in range(repeats):
            job.barrier(engine)
            t0 = time.perf_counter()
            engine.step_async(args.warmup + r * args.steps, args.steps)
            engine.sync()
            t1 = time.perf_counter()
            t_first = t0 if t_first is None else t_first
            t_last = t1
            times.append(job.reduce(engine.last_step_ms(), 'max'))
            job.barrier()
        launches, steps_per_launch, kernel = engine.last_launch_info()
        total_launches = int(job.reduce(launches, 'sum', 'int64'))
    finally:
        # one communicator / IPC mapping set per rank: drop this context before the next record
        state = field.__dict__.get('_engine_state')
        if state is not None and state.engine is not None:
            state.engine.close()
            state.engine = None
        del run, engine
    device_ms = float(np.median(times))
    cells = nx * ny
    per_gpu_cells = nx * (rows if not strong else -(-ny // job.world))
    peak, peak_source = measured_peak()
    step_launches = -(-args.steps // max(steps_per_launch, 1))
    launch_ms = device_ms / max(step_launches, 1)
    achieved = BYTES_PER_CELL_UPDATE * per_gpu_cells * steps_per_launch / (launch_ms * 1e-3) / 1e9
    traffic = ncu_traffic(kernel, nx, per_gpu_cells // nx)
    record = {
        'value': cells * args.steps / (device_ms * 1e-3) / 1e9, 'unit': 'Gcell-updates/s',
        'ms_per_step': device_ms / args.steps,
        'grid': [nx, ny], 'rows_per_gpu': per_gpu_cells // nx,
        'repeats': {'n': repeats, 'statistic': 'median', 'ms_per_step_min': min(times) / args.steps,
                    'ms_per_step_max': max(times) / args.steps,
                    'value_min': cells * args.steps / (max(times) * 1e-3) / 1e9,
                    'value_max': cells * args.steps / (min(times) * 1e-3) / 1e9},
        'kernel': kernel, 'steps_per_launch': steps_per_launch, 'gpu_launches': total_launches,
        'roofline': {
            'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
            'frac': achieved / peak, 'traffic': traffic,
            'traffic_source': 'committed ncu capture of this kernel on this grid '
                              '(profiles/traffic.json), not a counter of this run' if traffic
                              else None,
            'peak_source': peak_source, 'bytes_per_cell_update': BYTES_PER_CELL_UPDATE,
            'launch_ms': launch_ms,
            'dram_gbs': traffic / (launch_ms * 1e6) if traffic else None,
            'dram_frac': traffic / (launch_ms * 1e6) / peak if traffic else None,
            'note': 'frac counts 48 B per cell update for each of the {} steps a launch advances '
                    '(SURVEY.md 8d); the state crosses the HBM once per launch, dram_frac is the '
                    'physical fraction'.format(steps_per_launch)},
    }
    if job.rank == 0 and job.sampler is not None:
        record['clocks'] = job.sampler.window(t_first, t_last)
    return record


def measure_e2e(job, nx, rows, ny):
    """The same K steps through the public API: host arrays in, host arrays and probe signals out."""
    import pyfds_b200 as fds
    from pyfds_b200 import parallel
    args = job.args
    total_steps = args.warmup + 2 * args.steps
    rng = np.random.default_rng(11 + job.rank)
    cells, per_gpu_cells = nx * ny, nx * rows
    field = build_field(fds, nx, ny, total_steps + 1, wall=not args.no_wall)
    for name in ('pressure', 'velocity_x', 'velocity_y'):
        values = getattr(field, name).values
        if job.world == 1:
            values[:] = 1e-3 * rng.standard_normal(cells)
        else:   # only the rows this rank owns are ever read
            values[job.rank * per_gpu_cells:(job.rank + 1) * per_gpu_cells] = \
                1e-3 * rng.standard_normal(per_gpu_cells)
    field.device_kernel = args.kernel
    if job.world == 1:
        runner = field
    else:
        runner = parallel.SlabRun(field, job.rank, job.world, device=job.local_rank,
                                  kernel=args.kernel)
    try:
        runner.simulate(args.warmup)             # includes assembly, context creation, first copies
        first_call = dict(field.__dict__.get('_last_run_profile') or {})
        # one more untimed call of the shape that is timed: the engine keeps per-call-shape state
        # (task tables of the row bands a K-step call is cut into), as any repeated simulate(K) would
        runner.simulate(args.steps)
        job.barrier()
        t0 = time.perf_counter()
        runner.simulate(args.steps)
        job.barrier()
        seconds = time.perf_counter() - t0
    finally:
        state = field.__dict__.get('_engine_state')
        if state is not None and state.engine is not None:
            state.engine.close()
            state.engine = None
    state_bytes = 3 * per_gpu_cells * 8
    return {'value': cells * args.steps / seconds / 1e9, 'unit': 'Gcell-updates/s',
            'h2d_bytes_per_step': job.world * state_bytes / args.steps,
            'd2h_bytes_per_step': job.world * (state_bytes + 4 * 8 * args.steps) / args.steps,
            'seconds': seconds, 'phases': field.__dict__.get('_last_run_profile'),
            'first_call_phases': first_call,
            'what': '{}.simulate({}) from host numpy arrays (page-locked on first use): upload of '
                    'p/vx/vy rows, boundary and probe tables, {} steps, download of p/vx/vy and '
                    'probe signals; warm-up = one call of {} steps and one of {} steps'.format(
                        'field' if job.world == 1 else 'SlabRun', args.steps, args.steps,
                        args.warmup, args.steps)}


OTHER_MODELS = {'config3': 3, 'config4': 4, 'config6_lossy_twin_of_config2': 6}
OTHER_MODEL_KEYS = ('model', 'grid', 'steps', 'ms_per_step', 'gcell_updates_per_s', 'kernel',
                    'steps_per_launch', 'bytes_per_cell_update', 'algorithmic_gbs')


def measure_other_model(number, steps=200, warmup=20):
    """One of the other BASELINE.json configurations, device resident (CUDA-event time of `steps`
    steps after `warmup`, random initial state): benchmarks/configs.py::run."""
    bench_dir = os.path.join(ROOT, 'benchmarks')
    if bench_dir not in sys.path:
        sys.path.insert(0, bench_dir)
    import configs
    line = configs.run(number, steps, warmup)
    return {key: line[key] for key in OTHER_MODEL_KEYS}


def check_parity(job, nx=4096, rows=512, steps=26):
    """Multi-GPU against single-GPU, bit for bit: a 4096 x (512 N) grid with sources, walls and probes
    on slab seams is stepped as N slabs; every rank also steps the whole grid on its own GPU and
    compares the rows it owns and all probe records."""
    import pyfds_b200 as fds
    from pyfds_b200 import parallel
    world, rank = job.world, job.rank
    ny = rows * world
    seams = sorted({r for g in range(1, world) for r in (g * rows - 1, g * rows)})[:12]

    def make():
        field = build_field(fds, nx, ny, steps + 1, seam_rows=seams)
        rng = np.random.default_rng(77)
        for name in ('pressure', 'velocity_x', 'velocity_y'):
            getattr(field, name).values = 1e-3 * rng.standard_normal(nx * ny)
        field.device_kernel = job.args.kernel
        return field

    def drop(field):
        state = field.__dict__.get('_engine_state')
        if state is not None and state.engine is not None:
            state.engine.close()
            state.engine = None

    whole = make()
    whole.device_index = job.local_rank
    whole.simulate(10)
    whole.simulate(steps - 10)
    drop(whole)
    slabbed = make()
    run = parallel.SlabRun(slabbed, rank, world, device=job.local_rank, kernel=job.args.kernel)
    run.simulate(10)
    run.simulate(steps - 10)
    own = run.cells
    drop(slabbed)
    ok = True
    for name in ('pressure', 'velocity_x', 'velocity_y'):
        a = np.ascontiguousarray(getattr(whole, name).values[own]).view(np.int64)
        b = np.ascontiguousarray(getattr(slabbed, name).values[own]).view(np.int64)
        ok = ok and bool(np.array_equal(a, b))
        for out_a, out_b in zip(getattr(whole, name).outputs, getattr(slabbed, name).outputs):
            sa = np.asarray(out_a.signals, dtype=np.float64).view(np.int64)
            sb = np.asarray(out_b.signals, dtype=np.float64).view(np.int64)
            ok = ok and sa.shape == sb.shape and bool(np.array_equal(sa, sb))
    all_ok = bool(job.reduce(1 if ok else 0, 'min', 'int64'))
    return {'ok': all_ok, 'grid': [nx, ny], 'slabs': world, 'steps': steps,
            'seam_rows_with_source_wall_probe': seams,
            'what': 'fields (rows owned by each rank) and all probe signals of the {}-slab run '
                    'bitwise equal to the single-GPU run of the same grid, checked on every rank'
                    .format(world)}


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=2000)
    parser.add_argument('--warmup', type=int, default=100)
    parser.add_argument('--repeats', type=int, default=5, help='timed repeats of the K steps')
    parser.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    parser.add_argument('--size', type=int, default=4096, help='grid is size x size per GPU')
    parser.add_argument('--kernel', type=int, default=0, help='0 auto, 1 one-step, 2 streaming')
    parser.add_argument('--no-cpu-baseline', action='store_true')
    parser.add_argument('--no-e2e', action='store_true')
    parser.add_argument('--no-target', action='store_true', help='skip the 16384^2 record')
    parser.add_argument('--no-config5', action='store_true', help='skip the 32768-wide records')
    parser.add_argument('--no-parity', action='store_true', help='skip the multi-GPU parity check')
    parser.add_argument('--no-other-models', action='store_true',
                        help='skip configs 3, 4 and 6 (benchmarks/configs.py)')
    parser.add_argument('--only', default='',
                        help='comma list: main,target,weak,strong,parity,e2e,other')
    parser.add_argument('--no-wall', action='store_true', help='experiment: drop the x=0 rigid line')
    parser.add_argument('--strong', action='store_true',
                        help='make the headline value the strong-scaling one: ONE size x size grid '
                             'cut into y-slabs over the ranks')
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3)
    args.repeats = max(args.repeats, 1)

    job = Job(args)
    if args.impl == 'reference':
        run_reference(args, job.rank, job.world)
        return
    job.init()
    rank, world = job.rank, job.world
    only = set(filter(None, args.only.split(',')))

    def wanted(name, default=True):
        return name in only if only else default

    nx, rows = args.size, args.size
    ny = rows * world
    if args.strong:
        if args.size % world:
            raise SystemExit('--strong needs a size that is a multiple of the number of GPUs')
        rows, ny = args.size // world, args.size

    t_wall = time.perf_counter()
    main_record = measure_resident(job, nx, rows, ny, strong=args.strong)

    def sub_record(fn, *a, **kw):
        """Secondary records must not take the headline down with them."""
        try:
            return fn(*a, **kw)
        except Exception as exc:      # noqa: BLE001  (reported in the line)
            return {'error': '{}: {}'.format(type(exc).__name__, exc)}

    e2e = None
    if wanted('e2e', not args.no_e2e):
        e2e = sub_record(measure_e2e, job, nx, rows, ny)
    target = None
    if world == 1 and wanted('target', not args.no_target):
        target = sub_record(measure_resident, job, 16384, 16384, 16384)
        if 'error' not in target:
            target['workload'] = workload_name(16384, 16384) + ' (north-star target size)'
    config5 = None
    if wanted('weak', not args.no_config5) or wanted('strong', not args.no_config5):
        config5 = {'what': 'BASELINE.json configs[4], Acoustic2D 32768-wide y-slabs: weak = 32768 x '
                           '4096 cells per GPU, strong = one 32768 x 32768 grid over the GPUs'}
        if wanted('weak', not args.no_config5):
            config5['weak'] = sub_record(measure_resident, job, 32768, 4096, 4096 * world,
                                         repeats=3)
        if wanted('strong', not args.no_config5) and 32768 % world == 0:
            config5['strong'] = sub_record(measure_resident, job, 32768, 32768 // world, 32768,
                                           strong=True, repeats=3)
    parity = None
    if world > 1 and wanted('parity', not args.no_parity):
        parity = sub_record(check_parity, job)
    other_models = None
    if world == 1 and rank == 0 and wanted('other', not args.no_other_models):
        other_models = {name: sub_record(measure_other_model, number)
                        for name, number in OTHER_MODELS.items()}

    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        rate, seconds, kind, setup = cpu_reference_rate(4096, 4096, 1, 5)
        cpu = {'value': rate, 'unit': 'Gcell-updates/s', 'cores': 1, 'kind': kind,
               'sample': reference_sample_text(kind, 4096, 4096, 5, seconds, setup),
               'host_cores': os.cpu_count()}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': main_record['value'], 'unit': 'Gcell-updates/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': main_record['ms_per_step'],
            'higher_is_better': True, 'scaling': 'strong' if args.strong else 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_record(nx, rows, ny, world, args.strong),
            'engine': {'kernel': main_record['kernel'],
                       'steps_per_launch': main_record['steps_per_launch'],
                       'halo_exchange': None if world == 1 else
                       'edge rows stored into the neighbour slabs over NVLink peer memory by the '
                       'step kernel itself, edge tasks first; per-task flags, no kernel or host '
                       'call between sweeps'},
            'repeats': main_record['repeats'],
            'roofline': main_record['roofline'],
            'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': main_record['gpu_launches'],
            'clocks': main_record.get('clocks'),
            'target': target, 'config5': config5, 'other_models': other_models, 'parity': parity,
            'wall_seconds': time.perf_counter() - t_wall,
        }
        if parity is not None:
            line['parity_ok'] = bool(parity.get('ok', False))
        print(json.dumps(line), flush=True)
    job.close()


if __name__ == '__main__':
    main()
