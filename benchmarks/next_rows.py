#!/usr/bin/env python
"""Measures the rows either side of the hot path on one GPU, one JSON line each:

* ``flow``      AcousticFlow2D (pyfds/acoustic_flow.py) -- device-resident Gcell-updates/s with the row
                shift on the device, next to the same grid without flow;
* ``frames``    FrameStream (the frames of pyfds/gfx.py:72-86) -- frames/s and Gcell-updates/s of a
                run that delivers a decimated frame every `steps_per_frame` steps, next to the plain
                `simulate` of the same number of steps;
* ``coupled``   ThermoAcoustic1D (pyfds/coupled_fields.py) -- steps/s of `simulate` with the device
                session against the per-step seam (full host coherence around every step).

    python benchmarks/next_rows.py [--size 4096] [--steps 400]
"""

import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import pyfds_b200 as fds  # noqa: E402
from pyfds_b200 import _engine  # noqa: E402


def device_rate(field, steps, warmup=40):
    field.assemble_matrices()
    engine = _engine.prepare(field)
    _engine.upload_run_tables(field, engine, 0, warmup + steps)
    rng = np.random.default_rng(1)
    for c in range(3):
        engine.upload_state(c, 1e-3 * rng.standard_normal(engine.owned))
    engine.step_async(0, warmup)
    engine.sync()
    engine.step_async(warmup, steps)
    engine.sync()
    ms = engine.last_step_ms()
    launches, spl, kernel = engine.last_launch_info()
    return field.num_points * steps / (ms * 1e-3) / 1e9, launches, kernel, engine


def flow(size, steps):
    import bench
    plain = bench.build_field(fds, size, size, steps + 41)
    base, base_launches, kernel, engine = device_rate(plain, steps)
    engine.close()
    del plain, engine

    # 10 m/s in the lower half, 25 m/s in the upper half: rows move every 1000 / 400 steps
    speeds = np.where(np.arange(size) < size // 2, 10.0, 25.0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        moving = fds.AcousticFlow2D(speeds, t_delta=1e-7, t_samples=steps + 41, x_delta=1e-3,
                                    x_samples=size, y_delta=1e-3, y_samples=size,
                                    material=fds.AcousticMaterial(1500, 1000))
    template = bench.build_field(fds, size, size, steps + 41)
    moving.material_regions = template.material_regions
    for name in ('pressure', 'velocity_x', 'velocity_y'):
        getattr(moving, name).boundaries = getattr(template, name).boundaries
        getattr(moving, name).outputs = getattr(template, name).outputs
    rate, launches, kernel, engine = device_rate(moving, steps)
    print(json.dumps({
        'row': 'flow', 'workload': 'AcousticFlow2D {0}x{0}, config-2 layout, rows moving every '
                                   '1000 / 400 steps'.format(size),
        'value': rate, 'unit': 'Gcell-updates/s', 'same_grid_without_flow': base,
        'steps': steps, 'launches': launches, 'launches_without_flow': base_launches,
        'flow_shift_passes': engine.last_flow_shifts(), 'kernel': kernel}), flush=True)


def frames(size, steps, steps_per_frame=20, decimate=(4, 4)):
    import bench
    num_frames = steps // steps_per_frame
    plain = bench.build_field(fds, size, size, 2 * steps + 1)
    plain.simulate(steps)                          # context, page locking, first copies
    t0 = time.perf_counter()
    plain.simulate(steps)
    plain_s = time.perf_counter() - t0
    del plain

    field = bench.build_field(fds, size, size, 2 * steps + 1)
    for _ in fds.FrameStream(field, 'pressure', steps_per_frame, num_frames, decimate=decimate):
        pass
    t0 = time.perf_counter()
    count, checksum = 0, 0.0
    for _, frame in fds.FrameStream(field, 'pressure', steps_per_frame, num_frames,
                                    decimate=decimate):
        count += 1
        checksum += float(frame[0, 0])
    seconds = time.perf_counter() - t0
    print(json.dumps({
        'row': 'frames', 'workload': 'Acoustic2D {0}x{0} config 2, pressure frame every {1} steps, '
                                     'decimated {2}x{3}, host arrays in and out'.format(
                                         size, steps_per_frame, *decimate),
        'value': count / seconds, 'unit': 'frames/s',
        'gcell_updates_per_s': size * size * count * steps_per_frame / seconds / 1e9,
        'plain_simulate_gcell_updates_per_s': size * size * steps / plain_s / 1e9,
        'frames': count, 'frame_shape': list(frame.shape),
        'frame_bytes': int(frame.nbytes)}), flush=True)


def coupled(nx=10000, steps=20000):
    def build():
        f = fds.ThermoAcoustic1D(x_samples=nx, x_delta=1e-3, t_samples=steps + 50, t_delta=1e-7,
                                 thermal_material=fds.ThermalMaterial(900, 2700, 200),
                                 acoustic_material=fds.AcousticMaterial(700, 0.01,
                                                                        shear_viscosity=1e-3))
        f.fields[0].pressure.add_boundary(f.fields[0].get_point_region(1000 * 1e-3),
                                          value=np.sin(0.05 * np.arange(steps + 50)), additive=True)
        f.fields[1].temperature.add_output(f.fields[1].get_point_region(1001 * 1e-3))
        return f
    rates, sessions = {}, {}
    for mode in ('session', 'per_step'):
        f = build()
        f.device_session = mode == 'session'
        f.simulate(50)
        n = steps if mode == 'session' else max(50, steps // 10)
        t0 = time.perf_counter()
        f.simulate(n)
        rates[mode] = n / (time.perf_counter() - t0)
        sessions[mode] = getattr(f, '_last_session', None)
    print(json.dumps({
        'row': 'coupled', 'workload': 'ThermoAcoustic1D {} cells, viscous loss -> temperature every '
                                      'step'.format(nx),
        'value': rates['session'], 'unit': 'steps/s', 'per_step_seam_steps_per_s': rates['per_step'],
        'session': sessions['session'], 'steps': steps}), flush=True)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--size', type=int, default=4096)
    parser.add_argument('--steps', type=int, default=400)
    parser.add_argument('--rows', default='flow,frames,coupled')
    args = parser.parse_args()
    for row in args.rows.split(','):
        {'flow': lambda: flow(args.size, args.steps),
         'frames': lambda: frames(args.size, args.steps),
         'coupled': coupled}[row]()


if __name__ == '__main__':
    main()
