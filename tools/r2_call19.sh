#!/bin/bash
# round 2, call 19: final evidence of the build as committed -- the GPU suite, configs 3 / 6 with the
# default choice of resident CTAs, and an ncu capture of the config-3 kernel (2 CTAs per SM).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 115 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2g_pytest_gpu.log
timeout 40 python benchmarks/configs.py --configs 3,6 --steps 200 --warmup 20 > gpurun_out/r2g_configs.jsonl 2>gpurun_out/r2g_cfg.err
python - <<'PY'
import json
for l in open('gpurun_out/r2g_configs.jsonl'):
    c=json.loads(l); print(c['config'], c['kernel'], round(c['gcell_updates_per_s'],1))
PY
timeout 60 ncu --set full --clock-control none --import-source on -c 1 -f -k regex:streamv_kernel -s 3 \
    -o gpurun_out/r2g_streamv_c3 python benchmarks/configs.py --configs 3 --steps 20 --warmup 10 \
    > gpurun_out/r2g_ncu_c3.log 2>&1
tail -2 gpurun_out/r2g_ncu_c3.log | cut -c1-300
timeout 20 python tools/ncu_summary.py gpurun_out/r2g_streamv_c3.ncu-rep -o gpurun_out/prof_r2g_streamv_c3.md > /dev/null
timeout 20 ncu -i gpurun_out/r2g_streamv_c3.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r2g_streamv_c3_raw.csv
rm -f gpurun_out/r2g_streamv_c3.ncu-rep
head -45 gpurun_out/prof_r2g_streamv_c3.md | tail -38
