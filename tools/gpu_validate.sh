#!/bin/bash
# Round validation on one B200 (run under gpurun): smoke, the full -m gpu suite, the default
# bench in the driver's form, the reference (CPU) arm, and the other BASELINE workloads.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_all.log | cut -c1-300
echo "== default bench (driver form)"; timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_default.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
for w in elemwise lstm logreg readme; do
  echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$w.json"))
    print("$w", round(d["ms_per_step"], 4), "ms", d["roofline"]["unit"], round(d["roofline"]["achieved"], 1), "frac", round(d["roofline"]["frac"], 3), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print("$w FAILED", e)
PY
done
