#!/usr/bin/env python
"""The weight-gradient product of cfg3 (4096 x 4096 x 65536, bf16) in three operand layouts, two
launches each, for an ncu capture: what differs between MN-major and K-major operands."""
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
B, H = 65536, 4096
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(B, H, device="cuda", generator=g)
D = torch.randn(B, H, device="cuda", generator=g)
dX, dD = DeviceArray.from_torch(X), DeviceArray.from_torch(D)
dXt, dDt = DeviceArray.from_torch(X.t().contiguous()), DeviceArray.from_torch(D.t().contiguous())
C = DeviceArray.empty((H, H), "float32")
for name, (A, Bm) in {
    "mn_mn": (dX.dimshuffle([1, 0]), dD),
    "k_k": (dXt, dDt.dimshuffle([1, 0])),
    "k_mn": (dXt, dD),
}.items():
    cache = K.PackCache()
    for _ in range(2):
        K.gemm(C, 1.0, A, Bm, 0.0, precision=2, cache=cache)
    torch.cuda.synchronize()
    print(name, "done", flush=True)
