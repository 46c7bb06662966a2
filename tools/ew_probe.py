#!/usr/bin/env python
"""HBM bandwidth of an Elemwise over transposed views: a.T * b + 1 at 8192 x 8192 float32
(3 x 256 MiB of traffic), tiled kernel vs index-arithmetic kernel."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from aesara_b200.runtime import kernels as K
from aesara_b200.runtime import lib
from aesara_b200.runtime.device import DeviceArray

lib.check(lib.load().ab_init(0))
torch.cuda.set_device(0)
n = 8192
a = DeviceArray.from_torch(torch.randn(n, n, device="cuda"))
b = DeviceArray.from_torch(torch.randn(n, n, device="cuda"))
out = DeviceArray.empty((n, n), "float32")
expr = {"inputs": ["float32", "float32"], "out_dtypes": ["float32"], "outputs": ["t1"], "name": "aT_b_1",
        "stmts": [{"op": "mul", "args": ["i0", "i1"], "dtype": "float32", "in_dtypes": ["float32", "float32"]},
                  {"op": "add", "args": ["t0", {"const": 1.0, "dtype": "float32"}], "dtype": "float32",
                   "in_dtypes": ["float32", "float32"]}]}
kern = K.ElemwiseKernel.get(expr)
res = {}
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6482.4
VARIANTS = [("ab_ew_tile", None, None), ("ab_ew_nd", "1", None)] + [
    (f"ab_ew_tile_band{g}", None, str(g)) for g in (1, 2, 4, 16, 32)]
for name, env, band in VARIANTS:
    if env:
        os.environ["AB_EW_NO_TILE"] = env
    if band:
        os.environ["AB_EW_TILE_BAND"] = band
        kern = K.ElemwiseKernel(expr)      # a fresh module with that tile order
        os.environ.pop("AB_EW_TILE_BAND")
    ins = [a.dimshuffle([1, 0]), b]
    for _ in range(3):
        kern.launch((n, n), ins, [out])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        kern.launch((n, n), ins, [out])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gbs = 3 * n * n * 4 / ms / 1e6
    res[name] = {"ms": ms, "gbs": gbs, "frac_of_measured_hbm": gbs / peak}
    print(f"a.T*b+1 {n}x{n} f32  {name:10s} {ms:7.3f} ms  {gbs:7.0f} GB/s  = {gbs/peak:.2f} of the measured copy bandwidth")
    os.environ.pop("AB_EW_NO_TILE", None)
# yardsticks on the same box: a library transposing Elemwise (torch) and the plain copy, and a
# shape that is not a power of two (column stride 32 800 B instead of 32 768 B)
def _time(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ta, tb = a.owner, b.owner
to = torch.empty(n, n, device="cuda")
ms = _time(lambda: torch.add(torch.mul(ta.t(), tb, out=to), 1.0, out=to))
res["torch_mul_t_then_add (2 passes, 5 x 256 MiB)"] = {"ms": ms, "gbs": 5 * n * n * 4 / ms / 1e6}
ms = _time(lambda: to.copy_(ta.t()))
res["torch_copy_transposed (2 x 256 MiB)"] = {"ms": ms, "gbs": 2 * n * n * 4 / ms / 1e6, "frac_of_measured_hbm": 2 * n * n * 4 / ms / 1e6 / peak}
ms = _time(lambda: to.copy_(ta))
res["torch_copy (2 x 256 MiB)"] = {"ms": ms, "gbs": 2 * n * n * 4 / ms / 1e6, "frac_of_measured_hbm": 2 * n * n * 4 / ms / 1e6 / peak}
m = 8200
a2 = DeviceArray.from_torch(torch.randn(m, m, device="cuda"))
b2 = DeviceArray.from_torch(torch.randn(m, m, device="cuda"))
out2 = DeviceArray.empty((m, m), "float32")
kern = K.ElemwiseKernel.get(expr)
ins2 = [a2.dimshuffle([1, 0]), b2]
ms = _time(lambda: kern.launch((m, m), ins2, [out2]))
res["ab_ew_tile_8200"] = {"ms": ms, "gbs": 3 * m * m * 4 / ms / 1e6, "frac_of_measured_hbm": 3 * m * m * 4 / ms / 1e6 / peak}
for k in list(res)[-4:]:
    print(k, res[k])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ew_probe.json"), "w"), indent=1)
