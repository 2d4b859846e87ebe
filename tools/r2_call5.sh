#!/bin/bash
# round 2, GPU call 5 (8 GPUs): parity at 3 / 4 / 8 slabs, then the 1-2-4-8 scaling series as the driver runs it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c5_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "many_slabs and ((4- or -4]) or (8 and (big_lossless-0 or big_lossy or seams or flow_monotone)) or (3 and (big_lossless-0 or seams or thermal2d_wide)))" > gpurun_out/r2c5_pytest_multi.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2c5_pytest_multi.log
run() {  # label, gpus, extra args...
  local label="$1" g="$2"; shift 2
  if [ "$g" = 1 ]; then
    timeout 600 python bench.py --gpus 1 "$@" 2>gpurun_out/r2c5_$label.err | tail -1 > gpurun_out/r2c5_$label.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g "$@" 2>gpurun_out/r2c5_$label.err | tail -1 > gpurun_out/r2c5_$label.json
  fi
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c5_$label.json'))
    c5=d.get('config5') or {}
    e=d.get('e2e') or {}
    print('$label', 'value', round(d['value'],1), 'launch_ms', round(d['roofline']['launch_ms'],4), 'parity', d.get('parity_ok'),
          'weak5', round((c5.get('weak') or {}).get('value',0),1), 'strong5', round((c5.get('strong') or {}).get('value',0),1),
          'e2e', round(e.get('value',0),1), e.get('phases'), 'wall', round(d['wall_seconds'],1))
    for k in ('weak','strong'):
        if 'error' in (c5.get(k) or {}): print(k, c5[k]['error'])
    if d.get('parity') and 'error' in d['parity']: print('parity', d['parity']['error'])
    if 'error' in e: print('e2e', e['error'])
except Exception as ex:
    print('$label', 'failed', ex)
PY
}
run n1_driver 1 --steps 20 --warmup 5 --no-cpu-baseline
run n2_driver 2 --steps 20 --warmup 5
run n4_driver 4 --steps 20 --warmup 5
run n8_driver 8 --steps 20 --warmup 5
run n1_long 1 --steps 400 --warmup 40 --only main --no-cpu-baseline
run n8_long 8 --steps 400 --warmup 40 --only main
run n4_long 4 --steps 400 --warmup 40 --only main
tail -3 gpurun_out/r2c5_n8_driver.err
