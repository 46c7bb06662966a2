#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest blas + parity (2-CTA default)"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_all.log
echo "== pytest blas (1-CTA forced)"; AB_GEMM_1CTA=1 timeout 900 python -m pytest tests/test_gpu_blas.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_1cta.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_1cta.log
run() { # name, env..., args
  name=$1; shift
  echo "== bench $name: $*"; env "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "rc=$?"; tail -2 gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    print('$name', 'ms/step', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), 'dev_ms', d.get('device_ms'))
except Exception as e: print('$name FAILED', e)
PY
}
run mlp_2cta_mn timeout 600 python bench.py --workload mlp --steps 5 --warmup 3 --no-cpu --no-e2e
run mlp_2cta_nomn AB_GEMM_NO_MN=1 timeout 600 python bench.py --workload mlp --steps 5 --warmup 3 --no-cpu --no-e2e
run mlp_1cta_mn AB_GEMM_1CTA=1 timeout 600 python bench.py --workload mlp --steps 5 --warmup 3 --no-cpu --no-e2e
run mlp_1cta_nomn AB_GEMM_1CTA=1 AB_GEMM_NO_MN=1 timeout 600 python bench.py --workload mlp --steps 5 --warmup 3 --no-cpu --no-e2e
run mlp_fp32_2cta timeout 600 python bench.py --workload mlp --steps 3 --warmup 3 --precision fp32 --no-cpu --no-e2e
run lstm_2cta timeout 600 python bench.py --workload lstm --steps 3 --warmup 3 --no-cpu --no-e2e
echo "== ncu gemm 2cta"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 5 -c 2 -f -o gpurun_out/prof_gemm_bf16_v3 python bench.py --workload mlp --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_gemm3.log 2>&1; echo "rc=$?"
