#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_all.log | cut -c1-300
for seg in 4 8; do
echo "== mlp fp32 seg=$seg"; AB_GEMM_SEG_KB=$seg timeout 900 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['device_ms'], d['clocks'])"
done
