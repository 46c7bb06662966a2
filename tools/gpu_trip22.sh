#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 | cut -c1-300
echo "== default bench"; timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['ms_per_step'], d['device_ms'], d['roofline']['frac'], d['roofline_hbm']['frac'], d['e2e']['ms_per_step'], d['clocks'], d['gpu_launches'], d['cpu_baseline']['value'])"
echo "== ncu fused gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ab_gemm_ep -c 3 -f -o gpurun_out/prof_gemm_fused python bench.py --workload mlp --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_gemm_fused.log 2>&1; echo "rc=$?"
