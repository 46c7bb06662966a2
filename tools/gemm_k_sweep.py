"""3xTF32 GEMM error against float64 as K grows (run on the GPU box).

    AB_GEMM_SEG_KB=0 python tools/gemm_k_sweep.py   # whole K loop in the TMEM accumulator
    python tools/gemm_k_sweep.py                    # 128-element segments (default)
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aesara_b200.runtime import kernels as K, lib  # noqa: E402
from aesara_b200.runtime.device import DeviceArray  # noqa: E402

lib.check(lib.load().ab_init(0))
torch.cuda.set_device(0)
rows = []
for k in (128, 512, 2048, 4096, 16384, 65536):
    rng = np.random.default_rng(k)
    a = rng.standard_normal((256, k)).astype("float32")
    b = rng.standard_normal((k, 256)).astype("float32")
    C = DeviceArray.from_numpy(np.zeros((256, 256), "float32"))
    K.gemm(C, 1.0, DeviceArray.from_numpy(a), DeviceArray.from_numpy(b), 0.0, precision=0)
    want = a.astype(np.float64) @ b.astype(np.float64)
    got = C.to_numpy()
    nw = lambda g: float(np.max(np.abs(g - want)) / np.max(np.abs(want)))  # noqa: E731
    rel_rms = float(np.sqrt(np.mean((got - want) ** 2)) / np.sqrt(np.mean(want ** 2)))
    bias = float(np.mean((got - want) * np.sign(want)) / np.sqrt(np.mean(want ** 2)))
    rows.append({"K": k, "normwise_dev": nw(got), "normwise_cpu_f32": nw(a @ b), "rel_rms_dev": rel_rms,
                 "signed_bias_dev": bias})
print(json.dumps({"seg_kb": os.environ.get("AB_GEMM_SEG_KB", "default(4)"), "rows": rows}, indent=1))
