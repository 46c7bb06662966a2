#!/bin/bash
# One GPU-box session: parity tests, benches, ncu captures.  Logs go to gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest non-blas" ; timeout 900 python -m pytest tests -m gpu -q -k "not blas and not cfg3 and not cfg4 and not cfg1 and not cfg5 and not scan" -p no:cacheprovider > gpurun_out/pytest_a.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_a.log
echo "== pytest gemv/scan/cfg1/cfg5" ; timeout 600 python -m pytest tests -m gpu -q -k "cfg1 or cfg5 or gemv" -p no:cacheprovider > gpurun_out/pytest_b.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_b.log
echo "== pytest gemm family" ; timeout 600 python -m pytest tests -m gpu -q -k "blas_dot22 or blas_gemm or cfg3 or cfg4 or scan" -p no:cacheprovider > gpurun_out/pytest_c.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_c.log
echo "== bench elemwise"; timeout 900 python bench.py --workload elemwise --steps 20 --warmup 3 > gpurun_out/bench_elemwise.json 2> gpurun_out/bench_elemwise.err; echo "rc=$?"; cat gpurun_out/bench_elemwise.json
echo "== ncu elemwise"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ab_ew_flat_vec -s 3 -c 2 -f -o gpurun_out/prof_ew python bench.py --workload elemwise --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_ew.log 2>&1; echo "rc=$?"
echo "== bench mlp small"; timeout 600 python bench.py --workload mlp --batch 8192 --hidden 1024 --steps 5 --warmup 3 --precision fp32 --no-cpu > gpurun_out/bench_mlp_small_fp32.json 2> gpurun_out/bench_mlp_small.err; echo "rc=$?"; cat gpurun_out/bench_mlp_small_fp32.json
echo "== bench mlp full bf16"; timeout 900 python bench.py --workload mlp --steps 5 --warmup 3 --precision bf16 > gpurun_out/bench_mlp_bf16.json 2> gpurun_out/bench_mlp_bf16.err; echo "rc=$?"; cat gpurun_out/bench_mlp_bf16.json; tail -5 gpurun_out/bench_mlp_bf16.err
