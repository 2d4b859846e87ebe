#!/bin/bash
# round 2, GPU call 2 (1 GPU): the full default bench line, 16384^2 and the other configs with and
# without overlapped sweeps, launch list and one full ncu capture of the headline kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c2_bench_default.json 2> gpurun_out/r2c2_bench_default.err
echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"; tail -c 400 gpurun_out/r2c2_bench_default.err
for env in "X=1" "FDS_NO_OVERLAP=1"; do
  env $env timeout 300 python bench.py --steps 200 --warmup 20 --size 16384 --only main --no-cpu-baseline 2>>gpurun_out/r2c2.err | tail -1 > gpurun_out/r2c2_16384_$env.json
  python -c "import json,sys; d=json.load(open('gpurun_out/r2c2_16384_$env.json')); print('16384', '$env', round(d['value'],1), d['repeats']['value_min'], d['repeats']['value_max'], d['clocks'])"
  env $env timeout 300 python benchmarks/configs.py --configs 1,2,3,4,5,6,7,8 > gpurun_out/r2c2_configs_$env.jsonl 2>>gpurun_out/r2c2.err
  python -c "
import json
for l in open('gpurun_out/r2c2_configs_$env.jsonl'):
    d=json.loads(l); print('$env', d.get('config'), d.get('kernel'), round(d.get('gcell_updates_per_s',0),1))
"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/r2c2_launches.csv python bench.py --steps 40 --warmup 8 --only main,e2e --no-cpu-baseline \
    > gpurun_out/r2c2_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -c 1 -f -k regex:stream2d_kernel -s 6 \
    -o gpurun_out/r2c2_stream2d_4096 python bench.py --steps 40 --warmup 8 --only main --no-cpu-baseline \
    > gpurun_out/r2c2_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2c2_stream2d_4096.ncu-rep -o gpurun_out/r2c2_stream2d_4096.md > /dev/null
ls -la gpurun_out/r2c2_*
