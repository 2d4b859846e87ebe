#!/bin/bash
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "several_devices" > gpurun_out/r2c15_pytest.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/r2c15_pytest.log
python - <<PY
import sys, time
sys.path.insert(0, '.')
import numpy as np, bench, pyfds_b200 as fds
for devices in (None, list(range($N))):
    f = bench.build_field(fds, 4096, 4096 * $N, 500)
    rng = np.random.default_rng(0)
    for name in ('pressure', 'velocity_x', 'velocity_y'):
        getattr(f, name).values[:] = 1e-3 * rng.standard_normal(f.num_points)
    if devices: f.devices = devices
    f.simulate(8)
    for steps in (20, 400):
        t0 = time.perf_counter(); f.simulate(steps); dt = time.perf_counter() - t0
        print('devices', devices, 'simulate(%d)' % steps, round(dt * 1e3, 1), 'ms', round(f.num_points * steps / dt / 1e9, 1), 'Gcell-updates/s end to end', flush=True)
PY
