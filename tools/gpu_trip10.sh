#!/bin/bash
# ncu: persistent LSTM kernel and the bf16 2-CTA GEMM after the 8-warp epilogue change
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== ncu lstm_scan_kernel (T=16)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_scan -c 1 -f -o gpurun_out/prof_lstm_scan python bench.py --workload lstm --steps-t 16 --steps 1 --warmup 1 --graph 0 --no-e2e --no-cpu > gpurun_out/ncu_lstm.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_lstm.log
echo "== ncu gemm 2cta bf16"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 5 -c 2 -f -o gpurun_out/prof_gemm_bf16_v4 python bench.py --workload mlp --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_gemm.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_gemm.log
echo "== lstm T sweep (fast path, eager)"; for T in 8 32; do timeout 600 python bench.py --workload lstm --steps-t $T --steps 5 --warmup 3 --graph 0 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('T=$T ms', d['ms_per_step'], 'launches', d['gpu_launches'])"; done
