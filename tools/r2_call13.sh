#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_coupled_golden.py -x -q -m gpu -k "1d or config1 or coupl or thermo or reduced" > gpurun_out/r2c13_pytest.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2c13_pytest.log
for env in "X=1" "FDS_NO_OVERLAP=1"; do env $env timeout 200 python benchmarks/configs.py --configs 1 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$env', d['ms_per_step']*1000, 'us/step', d['launches'])"; done
python - <<'PY'
import logging, sys
sys.path.insert(0, '.')
logging.basicConfig(level=logging.INFO)
import bench, pyfds_b200 as fds
f = bench.build_field(fds, 1024, 1024, 50)
f.simulate(40)
PY
