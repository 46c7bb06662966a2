#!/usr/bin/env python
"""Where does the host-in / host-out evaluation of cfg3 spend its time?  (diagnostic)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from aesara_b200.runtime import lib
from aesara_b200.runtime.device import DeviceArray

lib.check(lib.load().ab_init(0))
torch.cuda.set_device(0)
B, H = 65536, 4096
X = torch.empty(B, H, dtype=torch.float32, pin_memory=True)
X.normal_()
xa = X.numpy()
sl = xa[8192:16384]
hb = torch.from_numpy(sl.reshape(-1).view(np.uint8))
print("slice pinned:", hb.is_pinned(), "full pinned:", torch.from_numpy(xa.reshape(-1).view(np.uint8)).is_pinned(), flush=True)
cur = torch.cuda.current_stream()
cs = torch.cuda.Stream()
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    keep = []
    for c in range(8):
        d, tok = DeviceArray.upload(xa[c * 8192:(c + 1) * 8192], cs, cur)
        keep.append((d, tok))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"8 chunk uploads of 128 MiB: enqueue {1e3*(t1-t0):.1f} ms, done {1e3*(t2-t0):.1f} ms -> {1.07/(t2-t0):.1f} GB/s", flush=True)
    del keep
torch.cuda.synchronize()
t0 = time.perf_counter()
d, tok = DeviceArray.upload(xa, cs, cur)
torch.cuda.synchronize()
print(f"one upload of 1 GiB: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)

# the chunked executor with host timestamps
from oracle import ref
ref.activate()
from aesara_b200.compat.bootstrap import load_aesara
aesara = load_aesara()
import aesara_b200.linker as L
from aesara_b200 import graphs as G

def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
    t.numpy()[...] = a
    return t

vals = G.cfg3_inputs(B, H)
pins = [pinned(v) for v in vals]
args = [p.numpy() for p in pins]
for chunks in (0, 8):
    i, o = G.cfg3_mlp()
    f = aesara.function(i, o, mode=L.mode(precision="bf16", host_chunks=chunks))
    f(*args)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        r = f(*args)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(f"host_chunks={chunks}: {1e3*min(ts):.1f} ms per call (chunks run: {getattr(f.vm.executor, 'chunks_run', 1)})", flush=True)
    if chunks:
        os.environ["AB_CHUNK_TRACE"] = "1"
        f(*args)
        del os.environ["AB_CHUNK_TRACE"]
        import cProfile
        import pstats

        pr = cProfile.Profile()
        pr.enable()
        f(*args)
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
