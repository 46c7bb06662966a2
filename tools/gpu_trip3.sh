#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest all gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_all.log
for w in "mlp --steps 5" "elemwise --steps 20" "readme --steps 200" "logreg --steps 5" "lstm --steps 3"; do
  name=$(echo $w | cut -d' ' -f1)
  echo "== bench $w"; timeout 1200 python bench.py --workload $w --warmup 3 > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "rc=$?"; tail -3 gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    print('$name', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],3), 'roofline', d['roofline']['achieved'], d['roofline']['frac'], 'hbm', (d.get('roofline_hbm') or {}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'launches', d['gpu_launches'], d['config'].get('executor'))
except Exception as e: print('$name FAILED', e)
PY
done
echo "== lstm eager fp32 (no graph) for comparison"; timeout 900 python bench.py --workload lstm --steps 2 --warmup 3 --graph 0 --no-e2e --no-cpu > gpurun_out/bench_lstm_eager.json 2> gpurun_out/bench_lstm_eager.err; echo "rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_lstm_eager.json'));print(d['ms_per_step'],d['device_ms'])"
echo "== ncu gemm bf16 persistent"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 5 -c 2 -f -o gpurun_out/prof_gemm_bf16_v2 python bench.py --workload mlp --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_gemm2.log 2>&1; echo "rc=$?"
