#!/bin/bash
# round 2, GPU call 3 (N GPUs, default 2): multi-GPU parity tests and the scaling lines.
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c3_topo.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2c3_pytest_multi_n$N.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2c3_pytest_multi_n$N.log
run() {  # label, gpus, extra args..., env through ENVV
  local label="$1" g="$2"; shift 2
  if [ "$g" = 1 ]; then
    env $ENVV timeout 600 python bench.py --gpus 1 "$@" 2>gpurun_out/r2c3_$label.err | tail -1 > gpurun_out/r2c3_$label.json
  else
    env $ENVV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g "$@" 2>gpurun_out/r2c3_$label.err | tail -1 > gpurun_out/r2c3_$label.json
  fi
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c3_$label.json'))
    c5=d.get('config5') or {}
    print('$label', 'value', round(d['value'],1), 'launch_ms', round(d['roofline']['launch_ms'],4), 'parity', d.get('parity_ok'),
          'weak5', round((c5.get('weak') or {}).get('value',0),1), 'strong5', round((c5.get('strong') or {}).get('value',0),1),
          'e2e', round((d.get('e2e') or {}).get('value',0),1), 'wall', round(d['wall_seconds'],1))
    for k in ('weak','strong'):
        if 'error' in (c5.get(k) or {}): print(k, c5[k]['error'])
    if d.get('parity') and 'error' in d['parity']: print('parity', d['parity']['error'])
    if d.get('e2e') and 'error' in d['e2e']: print('e2e', d['e2e']['error'])
except Exception as e:
    print('$label', 'failed', e)
PY
}
ENVV="X=1" run n1_long 1 --steps 400 --warmup 40 --only main --no-cpu-baseline
ENVV="X=1" run n${N}_long $N --steps 400 --warmup 40 --only main
ENVV="FDS_HALO_KERNELS=1" run n${N}_long_halokernels $N --steps 400 --warmup 40 --only main
ENVV="FDS_NO_OVERLAP=1" run n${N}_long_nooverlap $N --steps 400 --warmup 40 --only main
ENVV="X=1" run n1_driver 1 --steps 20 --warmup 5 --no-cpu-baseline
ENVV="X=1" run n${N}_driver $N --steps 20 --warmup 5
tail -3 gpurun_out/r2c3_n${N}_driver.err
