#!/bin/bash
# usage: gpu_trip_multi.sh N   (run under gpurun --gpus N)
set -u
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi -L > gpurun_out/gpus_$N.txt
for w in mlp logreg; do
  echo "== bench $w N=$N"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${w}_n$N.json 2> gpurun_out/bench_${w}_n$N.err; echo "rc=$?"; tail -3 gpurun_out/bench_${w}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${w}_n$N.json').read().strip().splitlines()[-1])
    print('$w', 'N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],3), d['config']['parallelism'])
except Exception as e: print('$w FAILED', e)
PY
done
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_ref_n$N.json
