#!/bin/bash
# smoke + compute-sanitizer (memcheck, racecheck) over representative GPU tests
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
SEL='test_gemm_fp32_faithful and (130-70-96 or 257-513-100 or 128-256-33) or test_gemv or test_ger or test_gemm_reduced or cfg3_mlp or cfg4_lstm or cfg5_logreg or cfg2_fused or softmax_classifier or indexing_embedding or careduce_sum_axes or scan_grad_rnn or test_lstm_medium'
echo "== memcheck"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_blas.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/sanitizer_memcheck.log | cut -c1-200
echo "== racecheck"; timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_blas.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "careduce_sum_axes or softmax_classifier or test_gemv or cfg5_logreg" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/sanitizer_racecheck.log | cut -c1-200
