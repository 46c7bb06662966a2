#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for T in 32 128; do
echo "== lstm eager T=$T"; timeout 900 python bench.py --workload lstm --steps-t $T --steps 5 --warmup 3 --graph 0 --no-cpu --no-e2e 2>/dev/null > gpurun_out/bench_lstm_T$T.json; python -c "
import json; d=json.load(open('gpurun_out/bench_lstm_T$T.json')); print(d['ms_per_step'], d['device_ms'], d['clocks'], d['gpu_launches'])"
done
echo "== lstm graph T=128"; timeout 900 python bench.py --workload lstm --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null > gpurun_out/bench_lstm_graph.json; python -c "
import json; d=json.load(open('gpurun_out/bench_lstm_graph.json')); print(d['ms_per_step'], d['device_ms'], d['clocks'], d['gpu_launches'])"
echo "== lstm trace T=128"; AB_LSTM_TRACE=1 timeout 600 python bench.py --workload lstm --steps 1 --warmup 1 --graph 0 --no-cpu --no-e2e 2>&1 | grep -A8 "ab_lstm_scan trace" | tail -9 | cut -c1-200
