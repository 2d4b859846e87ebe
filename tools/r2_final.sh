#!/bin/bash
# round 2, final evidence on ONE B200: the GPU suite, the bench lines as the driver runs them, the other
# configurations and next rows, the launch list and ncu captures of the headline kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_default.json 2>gpurun_out/r2f_def.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_bench_reference.json 2>gpurun_out/r2f_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 2000 --warmup 100 --only main,e2e --no-cpu-baseline > gpurun_out/r2f_bench_2000steps.json 2>gpurun_out/r2f_2000.err
timeout 300 python benchmarks/configs.py --configs 1,2,3,4,5,6,7,8 --check > gpurun_out/r2f_configs.jsonl 2>gpurun_out/r2f_cfg.err
timeout 300 python benchmarks/next_rows.py > gpurun_out/r2f_next_rows.jsonl 2>gpurun_out/r2f_next.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/r2f_launches.csv python bench.py --steps 20 --warmup 5 --only main,e2e --no-cpu-baseline \
    > gpurun_out/r2f_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -c 1 -f -k regex:stream2d_kernel -s 4 \
    -o gpurun_out/r2f_stream2d_16384 python bench.py --size 16384 --steps 12 --warmup 4 --only main --no-cpu-baseline --repeats 1 \
    > gpurun_out/r2f_ncu16k.log 2>&1
python tools/ncu_summary.py gpurun_out/r2f_stream2d_16384.ncu-rep -o gpurun_out/r2f_stream2d_16384.md > /dev/null
rm -f gpurun_out/r2f_stream2d_16384.ncu-rep
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_default.json'))
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'target', round(d['target']['value'],1), 'weak5', round(d['config5']['weak']['value'],1), 'strong5', round(d['config5']['strong']['value'],1), 'cpu', d['cpu_baseline']['value'], 'wall', round(d['wall_seconds'],1))
d=json.load(open('gpurun_out/r2f_bench_2000steps.json')); print('2000 steps', round(d['value'],1), d['repeats'], 'e2e', round(d['e2e']['value'],1), d['clocks'])
for l in open('gpurun_out/r2f_configs.jsonl'):
    c=json.loads(l); print(c['config'], c['kernel'], round(c['gcell_updates_per_s'],1), c.get('bitwise_equal_to_cpu_restatement_on_reduced_grid'))
for l in open('gpurun_out/r2f_next_rows.jsonl'):
    c=json.loads(l); print(c['row'], round(c['value'],1), c['unit'])
PY
