#!/usr/bin/env python
"""How much the rows that miss the branch-free bodies of the streaming kernels cost: a grid of vertical
material stripes (every strip of every kernel contains interfaces along y) and a grid with a source
column (a signal on every row of one column), each on the streaming kernel and on the one-step / tile
kernels. One JSON line per run."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import pyfds_b200 as fds                      # noqa: E402
from pyfds_b200 import _engine               # noqa: E402


def build(pattern, lossy, n, steps, klass='Acoustic2D'):
    mm = 1e-3
    f = getattr(fds, klass)(t_delta=1e-7, t_samples=steps, x_delta=mm, x_samples=n, y_delta=mm,
                            y_samples=n,
                            material=fds.AcousticMaterial(1500, 1000,
                                                          shear_viscosity=1e-3 if lossy else 0))
    second = fds.AcousticMaterial(1200, 900, absorption_coef=7.7 if lossy else None)
    top = (n - 1) * mm

    def rect(x0, x1, y0, y1):   # by index: sums of coordinates miss the reference's snap radius
        from pyfds_b200 import regions
        return regions.RectRegion._from_descriptor(('rect', x0, min(x1, n - 1), y0, min(y1, n - 1), n),
                                                   (x0 * mm, y0 * mm, 0, 0))
    if pattern == 'stripes':          # 20-cell stripes every 40 cells
        for x in range(0, n, 40):
            f.add_material_region(rect(x, x + 19, 0, n - 1), second)
    elif pattern == 'checker':        # 20 x 20 blocks: the map changes every 20 rows as well
        for y in range(0, n, 40):
            for x in range(0, n, 40):
                f.add_material_region(rect(x, x + 19, y, y + 19), second)
    elif pattern == 'source_columns':  # a signal on every row of 8 columns
        k = np.arange(steps)
        for x in range(n // 16, n, n // 8):
            f.pressure.add_boundary(f.get_line_region((x * mm, 0, x * mm, top)),
                                    value=np.sin(0.1 * k), additive=True)
    return f


def rate(f, steps, kernel):
    f.device_kernel = kernel
    f.assemble_matrices()
    engine = _engine.prepare(f)
    _engine.upload_run_tables(f, engine, 0, steps)
    rng = np.random.default_rng(1)
    for c in range(3):
        engine.upload_state(c, 1e-3 * rng.standard_normal(engine.owned))
    engine.step_async(0, steps // 4)
    engine.sync()
    engine.step_async(0, steps)
    engine.sync()
    ms = engine.last_step_ms()
    launches, spl, name = engine.last_launch_info()
    engine.close()
    return f.num_points * steps / (ms * 1e-3) / 1e9, name, spl


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--size', type=int, default=2048)
    parser.add_argument('--steps', type=int, default=40)
    parser.add_argument('--patterns', default='plain,stripes,checker,source_columns')
    args = parser.parse_args()
    for lossy in (False, True):
        for pattern in args.patterns.split(','):
            for kernel in (0, 3):
                value, name, spl = rate(build(pattern, lossy, args.size, args.steps), args.steps,
                                        kernel)
                print(json.dumps({'pattern': pattern, 'lossy': lossy, 'grid': args.size,
                                  'kernel': name, 'steps_per_launch': spl,
                                  'gcell_updates_per_s': round(value, 2)}), flush=True)


if __name__ == '__main__':
    main()
