#!/usr/bin/env python
"""Runs the five BASELINE.json configurations on one GPU (config 5: the single-GPU slab size) and
prints one JSON line per configuration: device-resident Gcell-updates/s, kernel, algorithmic GB/s.
Inputs follow SURVEY.md 8(d). Parity of the same code paths is covered by tests/test_gpu_parity.py;
here every config is additionally checked against the CPU restatement on a reduced grid (--check).

    python benchmarks/configs.py [--configs 1,2,3,4,5] [--steps 100] [--check]
"""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import pyfds_b200 as fds  # noqa: E402
from pyfds_b200 import _engine  # noqa: E402


def pulse(n, centre=200, width=60):
    k = np.arange(n)
    return np.sin(0.1 * k) * np.exp(-((k - centre) / width) ** 2)


def config1(t_samples=20000, nx=10000):
    """Acoustic1D linear pulse, 10k cells, 20k steps (doc/ex_acoustics.rst:14-37 scaled)."""
    fld = fds.Acoustic1D(t_delta=1e-7, t_samples=t_samples, x_delta=1e-3, x_samples=nx,
                         material=fds.AcousticMaterial(700, 0.01, shear_viscosity=1e-3))
    t = np.arange(t_samples) * 1e-7 - 0.3e-3
    # scipy.signal.gausspulse(t, 10e3, 0.7) written out: exp(-a t^2) cos(2 pi fc t)
    fc, bw, bwr = 10e3, 0.7, -6
    ref = 10.0 ** (bwr / 20.0)
    a = -(np.pi * fc * bw) ** 2 / (4.0 * np.log(ref))
    signal = np.exp(-a * t * t) * np.cos(2 * np.pi * fc * t)
    fld.velocity.add_boundary(fld.get_point_region(0))
    fld.pressure.add_boundary(fld.get_point_region((nx - 1) * 1e-3))
    fld.pressure.add_boundary(fld.get_point_region((nx // 10) * 1e-3), value=signal, additive=True)
    fld.pressure.add_output(fld.get_point_region((nx // 5) * 1e-3))
    return fld, 32


def config2(nx=4096, ny=4096, t_samples=1000):
    import bench
    return bench.build_field(fds, nx, ny, t_samples), 48


def config3(nx=8192, ny=4096, t_samples=1000, lossy=True):
    """Acoustic3DAxi with lossy sponge regions, Dirichlet lines, line source, line probe."""
    main = fds.AcousticMaterial(1500, 1000, shear_viscosity=1e-3 if lossy else 0)
    sponge = fds.AcousticMaterial(1500, 1000, absorption_coef=500) if lossy \
        else fds.AcousticMaterial(1400, 1100)
    fld = fds.Acoustic3DAxi(t_delta=1e-7, t_samples=t_samples, x_delta=1e-3, x_samples=nx,
                            y_delta=1e-3, y_samples=ny, material=main)
    w = max(2, nx // 128)
    X, Y = (nx - 1) * 1e-3, (ny - 1) * 1e-3
    fld.add_material_region(fld.get_rect_region(((nx - w) * 1e-3, 0, (w - 1) * 1e-3, Y)), sponge)
    fld.add_material_region(fld.get_rect_region((0, 0, X, (w - 1) * 1e-3)), sponge)
    fld.add_material_region(fld.get_rect_region((0, (ny - w) * 1e-3, X, (w - 1) * 1e-3)), sponge)
    for line in ((X, 0, X, Y), (0, 0, X, 0), (0, Y, X, Y)):
        fld.pressure.add_boundary(fld.get_line_region(line))
    fld.velocity_x.add_boundary(fld.get_line_region((0, 0, 0, Y)))
    fld.pressure.add_boundary(fld.get_line_region((0, (ny // 8) * 1e-3, (nx // 40) * 1e-3,
                                                   (ny // 8) * 1e-3)),
                              value=pulse(t_samples), additive=True)
    fld.pressure.add_output(fld.get_line_region((0, (ny // 4) * 1e-3, (nx // 40) * 1e-3,
                                                 (ny // 4) * 1e-3)))
    return fld, 48


def config4(nx=8192, ny=8192, t_samples=1000, klass='Thermal2D'):
    """Thermal2D explicit diffusion with mixed Dirichlet / Neumann boundaries."""
    fld = getattr(fds, klass)(t_delta=1e-3, t_samples=t_samples, x_delta=1e-3, x_samples=nx,
                              y_delta=1e-3, y_samples=ny,
                              material=fds.ThermalMaterial(900, 2700, 200))
    fld.add_material_region(
        fld.get_rect_region(((nx // 3) * 1e-3, (ny // 3) * 1e-3, (nx // 4) * 1e-3,
                             (ny // 4) * 1e-3)), fds.ThermalMaterial(450, 7800, (50, 30)))
    X, Y = (nx - 1) * 1e-3, (ny - 1) * 1e-3
    fld.temperature.add_boundary(fld.get_line_region((0, 0, 0, Y)), value=100)
    fld.temperature.add_boundary(fld.get_line_region((X, 0, X, Y)), value=0)
    fld.heat_flux_y.add_boundary(fld.get_line_region((0, 0, X, 0)), value=0)
    fld.heat_flux_y.add_boundary(fld.get_line_region((0, Y, X, Y)), value=0)
    for m in range(1, 5):
        fld.temperature.add_output(
            fld.get_point_region(((m * nx // 5) * 1e-3, (m * ny // 5) * 1e-3)))
    return fld, 16


def config5(nx=32768, ny=4096, t_samples=1000):
    """Config 5 per-GPU slab of the weak-scaling series (32768 x 4096 rows), run as a whole grid."""
    import bench
    return bench.build_field(fds, nx, ny, t_samples), 48


def config6(nx=4096, ny=4096, t_samples=1000):
    """The lossy twin of config 2 (SURVEY.md 8d 'C2 inputs'): shear_viscosity = 1e-3 in the main
    material, absorption_coef = 7.7 in the region; same source, wall and probes."""
    fld = fds.Acoustic2D(t_delta=1e-7, t_samples=t_samples, x_delta=1e-3, x_samples=nx,
                         y_delta=1e-3, y_samples=ny,
                         material=fds.AcousticMaterial(1500, 1000, shear_viscosity=1e-3))
    qx, qy = nx // 4, ny // 4
    fld.add_material_region(fld.get_rect_region((qx * 1e-3, qy * 1e-3, qx * 1e-3, qy * 1e-3)),
                            fds.AcousticMaterial(1200, 900, absorption_coef=7.7))
    k = np.arange(t_samples)
    fld.pressure.add_boundary(fld.get_point_region(((nx // 2) * 1e-3, (ny // 2) * 1e-3)),
                              value=np.sin(0.1 * k) * np.exp(-((k - 200) / 60) ** 2), additive=True)
    fld.velocity_x.add_boundary(fld.get_line_region((0, 0, 0, (ny - 1) * 1e-3)))
    for m in range(1, 5):
        fld.pressure.add_output(fld.get_point_region(((m * nx // 8) * 1e-3, (m * ny // 8) * 1e-3)))
    return fld, 48


def config7(**kwargs):
    """Config 3 without losses: the axisymmetric model on the lossless streaming kernel."""
    return config3(lossy=False, **kwargs)


def config8(**kwargs):
    """Config 4 as Thermal3DAxi."""
    return config4(klass='Thermal3DAxi', **kwargs)


CONFIGS = {1: config1, 2: config2, 3: config3, 4: config4, 5: config5, 6: config6, 7: config7,
           8: config8}
SMALL = {1: dict(t_samples=600, nx=3000), 2: dict(nx=256, ny=192, t_samples=60),
         3: dict(nx=256, ny=160, t_samples=50), 4: dict(nx=192, ny=160, t_samples=80),
         5: dict(nx=512, ny=96, t_samples=40), 6: dict(nx=256, ny=192, t_samples=60),
         7: dict(nx=256, ny=160, t_samples=50), 8: dict(nx=192, ny=160, t_samples=80)}


def check(number):
    """Reduced-size run of the same scenario against the CPU restatement, bit for bit."""
    import scenarios
    from oracle import restate
    field, _ = CONFIGS[number](**SMALL[number])
    steps = SMALL[number]['t_samples']
    rng = np.random.default_rng(number)
    for name in field._device_components:
        scale = 20.0 if name == 'temperature' else 1e-3
        getattr(field, name).values = scale * rng.standard_normal(field.num_points)
    stepper = restate.stepper_for(field).run(steps)
    field.simulate(steps)
    got, expected = scenarios.collect(field), scenarios.collect_stepper(stepper)
    for key in expected:
        a = np.ascontiguousarray(got[key], dtype=np.float64).view(np.int64)
        b = np.ascontiguousarray(expected[key], dtype=np.float64).view(np.int64)
        if a.shape != b.shape or not np.array_equal(a, b):
            return False
    return True


def run(number, steps, warmup):
    field, bytes_per_update = CONFIGS[number]()
    if number == 1:
        steps, warmup = 20000, 0
    field.assemble_matrices()
    t0 = time.perf_counter()
    engine = _engine.prepare(field)
    _engine.upload_run_tables(field, engine, 0, steps + warmup)
    rng = np.random.default_rng(number)
    if number != 1:
        for c, name in enumerate(field._device_components):
            scale = 20.0 if name == 'temperature' else 1e-3
            engine.upload_state(c, scale * rng.standard_normal(engine.owned))
    setup = time.perf_counter() - t0
    if warmup:
        engine.step_async(0, warmup)
        engine.sync()
    t0 = time.perf_counter()
    engine.step_async(warmup, steps)
    engine.sync()
    wall = time.perf_counter() - t0
    ms = engine.last_step_ms()
    launches, spl, kernel = engine.last_launch_info()
    cells = field.num_points
    rate = cells * steps / (ms * 1e-3) / 1e9
    return {'config': number, 'model': type(field).__name__,
            'grid': [field.x.samples, field.y.samples if hasattr(field, 'y') else 1],
            'steps': steps, 'device_ms': ms, 'ms_per_step': ms / steps, 'wall_s': wall,
            'gcell_updates_per_s': rate, 'kernel': kernel, 'launches': launches,
            'steps_per_launch': spl, 'algorithmic_gbs': rate * bytes_per_update,
            'bytes_per_cell_update': bytes_per_update, 'setup_s': setup,
            'device_gb': engine.device_bytes() / 1e9}


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--configs', default='1,2,3,4,5')
    parser.add_argument('--steps', type=int, default=100)
    parser.add_argument('--warmup', type=int, default=10)
    parser.add_argument('--check', action='store_true')
    args = parser.parse_args()
    for number in [int(k) for k in args.configs.split(',')]:
        line = run(number, args.steps, args.warmup)
        if args.check:
            line['bitwise_equal_to_cpu_restatement_on_reduced_grid'] = check(number)
        print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
