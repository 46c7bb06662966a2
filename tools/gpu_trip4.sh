#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest all gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_all.log
for w in "mlp --steps 5" "mlp --steps 5 --precision fp32 --no-cpu --no-e2e" "lstm --steps 3" "lstm --steps 2 --graph 0 --no-cpu --no-e2e"; do
  name=$(echo $w | tr ' ' '_' | tr -d '-')
  echo "== bench $w"; timeout 1200 python bench.py --workload $w --warmup 3 > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "rc=$?"; tail -3 gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    print('$name', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],3), 'roofline', d['roofline']['achieved'], d['roofline']['frac'], 'hbm', (d.get('roofline_hbm') or {}).get('frac'), 'dev_ms', d.get('device_ms'), 'e2e', (d.get('e2e') or {}).get('value'))
except Exception as e: print('$name FAILED', e)
PY
done
echo "== launch list mlp bf16"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_mlp_bf16_v2.csv python bench.py --workload mlp --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_mlp.log 2>&1; echo "rc=$?"
echo "== launch list lstm"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches_lstm.csv python bench.py --workload lstm --steps 1 --warmup 3 --graph 0 --no-e2e --no-cpu > gpurun_out/launches_lstm.log 2>&1; echo "rc=$?"
