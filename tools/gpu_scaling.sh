#!/bin/bash
# usage: gpu_scaling.sh N   (run under gpurun --gpus N): benches only, short
set -u
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for w in mlp logreg; do
  echo "== bench $w N=$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_${w}_n$N.json 2> gpurun_out/bench_${w}_n$N.err; echo "rc=$?"; tail -2 gpurun_out/bench_${w}_n$N.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${w}_n$N.json').read().strip().splitlines()[-1])
    print('$w', 'N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],3), d['config']['parallelism'])
except Exception as e: print('$w FAILED', e)
PY
done
