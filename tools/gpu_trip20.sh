#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== fusion tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "mlp or cfg3 or blas_gemm_alpha_beta" 2>&1 | tail -15 | cut -c1-300
echo "== default bench"; timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; tail -2 gpurun_out/bench_default.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['ms_per_step'], d['device_ms'], d['roofline']['frac'], d['roofline_hbm']['frac'], d['e2e']['ms_per_step'], d['clocks'], d['gpu_launches'])"
echo "== unfused"; AB_NO_GEMM_FUSE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['device_ms'])"
echo "== mlp fp32 fused"; timeout 900 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['device_ms'])"
