#!/bin/bash
# round 2, final 8-GPU validation: one-process slabs at 3/4/8, SPMD subset, the scaling series, drop-in timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "several_devices" > gpurun_out/r2c16_pytest_local.log 2>&1
echo "local rc=$?"; tail -3 gpurun_out/r2c16_pytest_local.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "many_slabs and ((8 and (big_lossless-0 or seams or flow_monotone)) or (4 and (big_lossy or thermal2d_wide or lossy_in_one_slab)))" > gpurun_out/r2c16_pytest_spmd.log 2>&1
echo "spmd rc=$?"; tail -3 gpurun_out/r2c16_pytest_spmd.log
run() {
  local label="$1" g="$2"; shift 2
  if [ "$g" = 1 ]; then
    timeout 600 python bench.py --gpus 1 "$@" 2>gpurun_out/r2c16_$label.err | tail -1 > gpurun_out/r2c16_$label.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g "$@" 2>gpurun_out/r2c16_$label.err | tail -1 > gpurun_out/r2c16_$label.json
  fi
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c16_$label.json'))
    c5=d.get('config5') or {}
    e=d.get('e2e') or {}
    print('$label', 'value', round(d['value'],1), 'launch_ms', round(d['roofline']['launch_ms'],4), 'parity', d.get('parity_ok'),
          'weak5', round((c5.get('weak') or {}).get('value',0),1), 'strong5', round((c5.get('strong') or {}).get('value',0),1),
          'e2e', round(e.get('value',0),1), 'wall', round(d['wall_seconds'],1))
except Exception as ex:
    print('$label', 'failed', ex)
PY
}
run n1_driver 1 --steps 20 --warmup 5 --no-cpu-baseline
run n2_driver 2 --steps 20 --warmup 5
run n4_driver 4 --steps 20 --warmup 5
run n8_driver 8 --steps 20 --warmup 5
python - <<PY
import sys, time
sys.path.insert(0, '.')
import numpy as np, bench, pyfds_b200 as fds
for n in (1, 8):
    f = bench.build_field(fds, 4096, 4096 * 8, 500)
    rng = np.random.default_rng(0)
    block = 1e-3 * rng.standard_normal(1 << 22)
    for name in ('pressure', 'velocity_x', 'velocity_y'):
        getattr(f, name).values[:] = np.resize(block, f.num_points)
    if n > 1: f.devices = list(range(n))
    f.simulate(8)
    for steps in (20, 400):
        t0 = time.perf_counter(); f.simulate(steps); dt = time.perf_counter() - t0
        print('one process,', n, 'GPU(s): simulate(%d)' % steps, round(dt * 1e3, 1), 'ms', round(f.num_points * steps / dt / 1e9, 1), 'Gcell-updates/s end to end (4096 x 32768)', flush=True)
PY
