#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_blas.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "gemm or mlp or cfg3 or blas or dot22" 2>&1 | tail -6 | cut -c1-300
echo "== default bench"; timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['device_ms'], d['roofline']['frac'], d['roofline_hbm']['frac'], d['clocks'])"
echo "== unfused"; AB_NO_GEMM_FUSE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['device_ms'])"
