#!/bin/bash
# last regression on the final commit (1 GPU): build check, smoke, GPU suite, both bench arms.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2fc_bench.err | tail -1 > gpurun_out/r2fc_bench.json; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2fc_bench.json'))
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'target', round(d['target']['value'],1), d['target']['roofline']['dram_frac'], 'weak5', round(d['config5']['weak']['value'],1), 'strong5', round(d['config5']['strong']['value'],1), 'cpu', round(d['cpu_baseline']['value'],4), d['cpu_baseline']['kind'], 'wall', round(d['wall_seconds'],1), d['clocks'])
PY
