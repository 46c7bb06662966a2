#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
SEL='test_gemm_split_k and 512 or test_mlp_gemm_epilogue_fusion or test_logreg_row_region_fusion or careduce_big_1d or cfg3_mlp or blas_gemm_alpha_beta or test_lstm_medium or test_pack_cache or scan_grad_rnn or blas_edge_shapes or views_negative'
echo "== memcheck"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_blas.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/sanitizer_memcheck2.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/sanitizer_memcheck2.log | cut -c1-200
echo "== racecheck"; timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "test_logreg_row_region_fusion and 333 or careduce_big_1d" > gpurun_out/sanitizer_racecheck2.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/sanitizer_racecheck2.log | cut -c1-200
