#!/bin/bash
# round 2, GPU call 4 (N GPUs): short edge tasks -- parity subset and scaling lines.
N=${1:-2}
SEL=${2:-"two_slabs"}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q -k "$SEL" > gpurun_out/r2c4_pytest_multi_n$N.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2c4_pytest_multi_n$N.log
run() {  # label, gpus, extra args..., env through ENVV
  local label="$1" g="$2"; shift 2
  if [ "$g" = 1 ]; then
    env $ENVV timeout 600 python bench.py --gpus 1 "$@" 2>gpurun_out/r2c4_$label.err | tail -1 > gpurun_out/r2c4_$label.json
  else
    env $ENVV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g "$@" 2>gpurun_out/r2c4_$label.err | tail -1 > gpurun_out/r2c4_$label.json
  fi
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2c4_$label.json'))
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
ENVV="X=1" run n1_long 1 --steps 400 --warmup 40 --only main --no-cpu-baseline
g=2
while [ $g -le $N ]; do
  ENVV="X=1" run n${g}_long $g --steps 400 --warmup 40 --only main
  ENVV="X=1" run n${g}_driver $g --steps 20 --warmup 5
  g=$((g*2))
done
ENVV="X=1" run n1_driver 1 --steps 20 --warmup 5 --no-cpu-baseline
