#!/usr/bin/env python
"""Tuning aid for the 1-D kernel: config 1 with parts of its boundary/probe set removed, to see what
the slowest warp of a launch is busy with. Prints ms per step for each variant."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'benchmarks')]
import pyfds_b200 as fds                      # noqa: E402
from pyfds_b200 import _engine               # noqa: E402
import configs                               # noqa: E402


def rate(field, steps=20000):
    field.assemble_matrices()
    engine = _engine.prepare(field)
    _engine.upload_run_tables(field, engine, 0, steps)
    engine.step_async(0, steps)
    engine.sync()
    engine.step_async(0, steps)
    engine.sync()
    ms = engine.last_step_ms()
    engine.close()
    return ms / steps * 1e3


for name in ('full', 'no_probe', 'no_source', 'no_walls', 'plain'):
    field, _ = configs.config1()
    if name in ('no_probe', 'plain'):
        field.pressure.outputs = []
    if name in ('no_source', 'plain'):
        field.pressure.boundaries = [b for b in field.pressure.boundaries if np.ndim(b.value) == 0]
    if name in ('no_walls', 'plain'):
        field.pressure.boundaries = [b for b in field.pressure.boundaries if np.ndim(b.value) != 0]
        field.velocity.boundaries = []
    print('{:10s} {:.3f} us/step'.format(name, rate(field)), flush=True)
