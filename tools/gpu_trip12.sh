#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== lstm tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lstm or scan" -p no:cacheprovider 2>&1 | tail -4
echo "== lstm trace 2cta"; AB_LSTM_TRACE=1 timeout 600 python bench.py --workload lstm --steps-t 32 --steps 1 --warmup 1 --graph 0 --no-cpu --no-e2e 2>&1 | grep -A8 "ab_lstm_scan trace" | tail -9 | cut -c1-300
echo "== lstm trace 1cta"; AB_LSTM_1CTA=1 AB_LSTM_TRACE=1 timeout 600 python bench.py --workload lstm --steps-t 32 --steps 1 --warmup 1 --graph 0 --no-cpu --no-e2e 2>&1 | grep -A8 "ab_lstm_scan trace" | tail -9 | cut -c1-300
echo "== lstm bench 2cta"; timeout 900 python bench.py --workload lstm --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | cut -c1-300
