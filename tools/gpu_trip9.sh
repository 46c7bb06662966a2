#!/bin/bash
# segmented-K accumulation: K sweep with/without segments, BLAS + parity tests, benches
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== k sweep (no segments)"; AB_GEMM_SEG_KB=0 timeout 300 python tools/gemm_k_sweep.py > gpurun_out/ksweep_noseg.json 2> gpurun_out/ksweep_noseg.err; echo "rc=$?"; cat gpurun_out/ksweep_noseg.json | tr -d '\n ' | cut -c1-1500; echo
echo "== k sweep (segments)"; timeout 300 python tools/gemm_k_sweep.py > gpurun_out/ksweep_seg.json 2> gpurun_out/ksweep_seg.err; echo "rc=$?"; cat gpurun_out/ksweep_seg.json | tr -d '\n ' | cut -c1-1500; echo
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_all.log
echo "== default bench"; timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; cut -c1-900 gpurun_out/bench_default.json
echo "== mlp fp32"; timeout 900 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_mlp_fp32.json 2> gpurun_out/bench_mlp_fp32.err; echo "rc=$?"; cut -c1-700 gpurun_out/bench_mlp_fp32.json
echo "== lstm fast"; timeout 900 python bench.py --workload lstm --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_lstm_fast.json 2> gpurun_out/bench_lstm_fast.err; echo "rc=$?"; cut -c1-900 gpurun_out/bench_lstm_fast.json; tail -3 gpurun_out/bench_lstm_fast.err
echo "== lstm general"; AB_SCAN_NO_FAST=1 timeout 900 python bench.py --workload lstm --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_lstm_general.json 2> gpurun_out/bench_lstm_general.err; echo "rc=$?"; cut -c1-700 gpurun_out/bench_lstm_general.json
