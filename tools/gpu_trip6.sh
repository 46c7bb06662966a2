#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_all.log
run() { name=$1; shift
  echo "== bench $name: $*"; env "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "rc=$?"; tail -2 gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    print('$name', 'ms/step', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), 'hbm', (d.get('roofline_hbm') or {}).get('frac'), 'dev_ms', d.get('device_ms'))
except Exception as e: print('$name FAILED', e)
PY
}
run mlp_u_default timeout 600 python bench.py --workload mlp --steps 5 --warmup 3 --no-cpu --no-e2e
run mlp_u1 AB_EW_UNROLL=1 timeout 600 python bench.py --workload mlp --steps 5 --warmup 3 --no-cpu --no-e2e
run mlp_u2 AB_EW_UNROLL=2 timeout 600 python bench.py --workload mlp --steps 5 --warmup 3 --no-cpu --no-e2e
echo "== ncu elemwise kernels in mlp"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ab_ew|ab_red" -s 24 -c 12 -f -o gpurun_out/prof_mlp_ew python bench.py --workload mlp --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_mlp_ew.log 2>&1; echo "rc=$?"
