#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== logreg tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "logreg" 2>&1 | tail -15 | cut -c1-300
for R in 1 2 4; do
echo "== bench logreg fused rows/iter=$R"; AB_ROWFUSE_ROWS=$R timeout 900 python bench.py --workload logreg --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline'], d['gpu_launches'])"
done
echo "== bench logreg node-by-node"; AB_NO_ROWFUSE=1 timeout 900 python bench.py --workload logreg --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline'], d['gpu_launches'])"
