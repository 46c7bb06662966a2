"""Host<->device copy rates on the box (pinned / pageable), to size the e2e pipeline."""
import json
import time

import numpy as np
import torch

torch.cuda.init()
n = 1 << 30
res = {}
hp = torch.empty(n, dtype=torch.uint8, pin_memory=True)
hp.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
pg = np.ones(n, dtype=np.uint8)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


res["h2d_pinned_GBps"] = n / timed(lambda: d.copy_(hp, non_blocking=True)) / 1e9
res["d2h_pinned_GBps"] = n / timed(lambda: hp.copy_(d, non_blocking=True)) / 1e9
tp = torch.from_numpy(pg)
res["h2d_pageable_GBps"] = n / timed(lambda: d.copy_(tp)) / 1e9
res["d2h_pageable_GBps"] = n / timed(lambda: tp.copy_(d)) / 1e9
res["d2h_cpu_call_GBps"] = n / timed(lambda: d.cpu()) / 1e9
# both directions at once on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
hp2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")


def both():
    with torch.cuda.stream(s1):
        d.copy_(hp, non_blocking=True)
    with torch.cuda.stream(s2):
        hp2.copy_(d2, non_blocking=True)


res["bidir_each_GBps"] = n / timed(both) / 1e9
t0 = time.perf_counter()
x = torch.empty(n, dtype=torch.uint8, pin_memory=True)
res["pin_alloc_1GiB_ms"] = (time.perf_counter() - t0) * 1e3
del x
t0 = time.perf_counter()
x = torch.empty(n, dtype=torch.uint8, pin_memory=True)
res["pin_alloc_1GiB_again_ms"] = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
hp2.copy_(hp)
res["host_memcpy_pinned_GBps"] = n / (time.perf_counter() - t0) / 1e9
print(json.dumps(res, indent=1))
