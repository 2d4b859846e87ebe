#!/bin/bash
# round 2, call 18: the viscous streaming kernel with 2 or 3 resident CTAs per SM (FDS_SV_CTAS) and as
# the generic or the main-material instantiation (FDS_SV_MAINONLY), configs 3 (axisymmetric) and 6
# (plain lossy), one library; parity of every combination on the viscous / axisymmetric tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2_c18_ab.jsonl; : > $out
for combo in "2 0" "3 0" "2 1" "3 1"; do
  set -- $combo
  for cfg in 3 6; do
    FDS_SV_CTAS=$1 FDS_SV_MAINONLY=$2 timeout 120 python benchmarks/configs.py --configs $cfg --steps 200 --warmup 20 2>>gpurun_out/r2_c18.err \
      | python -c "import sys,json; d=json.loads(sys.stdin.readline()); d['sv_ctas']=$1; d['sv_main_only']=$2; print(json.dumps(d))" >> $out
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2_c18_ab.jsonl'):
    d=json.loads(l); print('config', d['config'], 'ctas', d['sv_ctas'], 'main_only', d['sv_main_only'], round(d['gcell_updates_per_s'],1))
PY
for combo in "2 0" "2 1" "3 1"; do
  set -- $combo
  echo "parity ctas=$1 main_only=$2"
  FDS_SV_CTAS=$1 FDS_SV_MAINONLY=$2 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -m gpu \
    -k "division or viscous or axisymmetric or axi or reduced_grid" 2>&1 | tail -4 | tee gpurun_out/r2_c18_pytest_$1_$2.log
done
