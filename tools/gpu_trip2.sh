#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest all gpu"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench elemwise"; timeout 900 python bench.py --workload elemwise --steps 20 --warmup 3 > gpurun_out/bench_elemwise.json 2> gpurun_out/bench_elemwise.err; echo "rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_elemwise.json'));print(d['ms_per_step'],d['roofline']['achieved'],d['roofline']['frac'],d['e2e']['value'])"
echo "== ncu elemwise"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ab_ew_flat_vec -s 3 -c 1 -f -o gpurun_out/prof_ew2 python bench.py --workload elemwise --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_ew2.log 2>&1; echo "rc=$?"
echo "== launch list mlp bf16"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_mlp_bf16.csv python bench.py --workload mlp --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_mlp.log 2>&1; echo "rc=$?"
echo "== ncu gemm bf16"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 5 -c 2 -f -o gpurun_out/prof_gemm_bf16 python bench.py --workload mlp --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_gemm.log 2>&1; echo "rc=$?"
echo "== bench mlp fp32"; timeout 900 python bench.py --workload mlp --steps 5 --warmup 3 --precision fp32 --no-cpu > gpurun_out/bench_mlp_fp32.json 2> gpurun_out/bench_mlp_fp32.err; echo "rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_mlp_fp32.json'));print(d['ms_per_step'],d['roofline'],d['device_ms'])"
echo "== bench mlp tf32"; timeout 900 python bench.py --workload mlp --steps 5 --warmup 3 --precision tf32 --no-cpu --no-e2e > gpurun_out/bench_mlp_tf32.json 2> gpurun_out/bench_mlp_tf32.err; echo "rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_mlp_tf32.json'));print(d['ms_per_step'],d['roofline'],d['device_ms'])"
