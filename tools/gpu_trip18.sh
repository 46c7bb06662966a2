#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_all.log | cut -c1-300
echo "== logreg bench"; timeout 900 python bench.py --workload logreg --steps 10 --warmup 3 > gpurun_out/bench_logreg_fused.json 2> gpurun_out/bench_logreg_fused.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_logreg_fused.json')); print(d['ms_per_step'], d['roofline'], d['e2e'], d['cpu_baseline'], d['gpu_launches'])"
echo "== ncu rowfused"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ab_rowfused -c 1 -f -o gpurun_out/prof_rowfused python bench.py --workload logreg --steps 1 --warmup 1 --graph 0 --no-e2e --no-cpu > gpurun_out/ncu_rowfused.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_rowfused.log | cut -c1-200
echo "== launch list logreg"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_logreg_fused.csv python bench.py --workload logreg --steps 1 --warmup 1 --graph 0 --no-cpu --no-e2e > /dev/null 2>&1; echo "rc=$?"
