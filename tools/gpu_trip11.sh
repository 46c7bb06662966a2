#!/bin/bash
# 2-CTA LSTM kernel, pinned upload/download path: tests + benches
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_all.log
echo "== lstm fast 2cta"; timeout 900 python bench.py --workload lstm --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_lstm_2cta.json 2> gpurun_out/bench_lstm_2cta.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_lstm_2cta.json; tail -3 gpurun_out/bench_lstm_2cta.err
echo "== lstm fast 1cta"; AB_LSTM_1CTA=1 timeout 900 python bench.py --workload lstm --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_lstm_1cta.json 2> gpurun_out/bench_lstm_1cta.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_lstm_1cta.json
echo "== default bench"; timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['ms_per_step'], d['e2e'], d['cpu_baseline'])"
echo "== elemwise bench"; timeout 900 python bench.py --workload elemwise --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_elemwise.json 2> gpurun_out/bench_elemwise.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_elemwise.json')); print(d['ms_per_step'], d['e2e'])"
