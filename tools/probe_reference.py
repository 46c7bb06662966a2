#!/usr/bin/env python
"""How fast is the reference's own C-linker (Mode("cvm","fast_run")) on this host?

Times the cfg3 graph at a few batch sizes with both OpenMP settings so that bench.py's
reference arm can be sized (which sample fits the driver's window).  Checker-side tool:
uses the travelling copy of the reference (oracle/_ref)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count()))

import numpy as np  # noqa: E402

from oracle import ref  # noqa: E402

assert ref.activate(), "no reference available"
from aesara_b200 import graphs as G  # noqa: E402
from aesara_b200.compat.bootstrap import load_aesara  # noqa: E402

aesara = load_aesara()
from aesara.compile.mode import Mode  # noqa: E402

try:
    from threadpoolctl import threadpool_info, threadpool_limits

    threadpool_limits(limits=os.cpu_count())
    pools = [(p.get("internal_api"), p.get("num_threads")) for p in threadpool_info()]
except Exception:
    pools = None

H = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sizes = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1024, 8192]
res = {"cores": os.cpu_count(), "pools": pools, "H": H, "runs": []}
for openmp in (False, True):
    with aesara.config.change_flags(openmp=openmp):
        i, o = G.cfg3_mlp()
        t0 = time.perf_counter()
        f = aesara.function(i, o, mode=Mode("cvm", "fast_run"))
        t_compile = time.perf_counter() - t0
    f.trust_input = True
    for B in sizes:
        vals = G.cfg3_inputs(B, H)
        f(*vals)
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            f(*vals)
            ts.append(time.perf_counter() - t0)
        res["runs"].append({"openmp": openmp, "B": B, "s_per_eval": min(ts), "compile_s": t_compile})
        print(res["runs"][-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe_reference.json"), "w") as fh:
    json.dump(res, fh, indent=1)
