#!/bin/bash
# round 2, call 17: A/B of the lossy axisymmetric streaming kernel (BASELINE config 3) across builds of
# the library -- default (quotient sequence + per-column coefficients in shared memory), _fd (quotient
# sequence only), _c2 (default at 2 CTAs per SM), _base (the IEEE division per cell, as before) -- and
# parity of the new code paths.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2_c17_ab.jsonl; : > $out
for rep in 1 2; do
  for v in "" _fd _c2 _base; do
    lib=$PWD/pyfds_b200/libfdsb200$v.so
    [ -f "$lib" ] || continue
    FDS_LIBRARY_PATH=$lib timeout 120 python benchmarks/configs.py --configs 3 --steps 200 --warmup 20 2>>gpurun_out/r2_c17.err \
      | python -c "import sys,json; d=json.loads(sys.stdin.readline()); d['library']='libfdsb200$v.so'; print(json.dumps(d))" >> $out
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2_c17_ab.jsonl'):
    d=json.loads(l); print(d['library'], round(d['gcell_updates_per_s'],1), d['kernel'])
PY
# the lossy 2-D twin and the lossless axisymmetric kernel must not have moved
timeout 120 python benchmarks/configs.py --configs 6,7 --steps 200 --warmup 20 2>>gpurun_out/r2_c17.err | tee gpurun_out/r2_c17_c67.jsonl | cut -c1-220
# parity of the division paths, both builds that have them
for v in "" _fd; do
  FDS_LIBRARY_PATH=$PWD/pyfds_b200/libfdsb200$v.so timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
    -k "division or (viscous and Acoustic3DAxi and True) or axisymmetric" 2>&1 | tail -4 | tee gpurun_out/r2_c17_pytest_div$v.log
done
# the whole GPU suite on the default build
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2_c17_pytest_gpu.log
