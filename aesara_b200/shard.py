"""Leading-axis sharding across GPUs.

The reference has no parallelism of any kind; the one thing
that shards naturally is a graph whose leading axis is an independent batch
map and whose outputs are either per-row (concatenate) or batch reductions
(partial + combine).  One process per GPU runs the *same* lowered program on
its row block; the only exchange step is a single NCCL all-gather of the packed
outputs followed by a local combine executed by this backend's own CAReduce /
Elemwise kernels.

``row_block`` / ``pack_layout`` are pure host logic (covered by the gloo
world_size-2 tests on CPU); ``OutputCombiner`` is the device path.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np


def row_block(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block ``[start, stop)`` of ``rank``; blocks differ by at most one row."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_rows, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


@dataclass
class PackLayout:
    shapes: List[Tuple[int, ...]]
    offsets: List[int]
    total: int


def pack_layout(shapes: Sequence[Tuple[int, ...]], align: int = 4) -> PackLayout:
    """Offsets (in elements) of each output inside the flat exchange buffer."""
    offs, cur = [], 0
    for s in shapes:
        offs.append(cur)
        n = int(np.prod(s)) if len(s) else 1
        cur += (n + align - 1) // align * align
    return PackLayout([tuple(int(x) for x in s) for s in shapes], offs, cur)


def combine_weights(mode: str, rows_per_rank: Sequence[int]) -> List[float]:
    """Weights turning per-shard results into the global result:
    ``sum`` -> 1 each, ``mean`` -> n_r / N (a mean of shard means weighted by rows)."""
    if mode == "sum":
        return [1.0] * len(rows_per_rank)
    if mode == "mean":
        tot = float(sum(rows_per_rank))
        return [r / tot for r in rows_per_rank]
    raise ValueError(mode)


ALLREDUCE_MIN_ELEMS = 1 << 18  # payloads at least this big are pre-weighted and all-reduced


def exchange_plan(total_elems: int, world: int) -> str:
    """Which collective combines the packed outputs.

    ``allgather``  one ``all_gather_into_tensor`` + a local weighted sum (the north star's
                   single all-gather; right for the (D+2)-float payload of cfg5 — latency bound);
    ``allreduce``  the local buffer is multiplied by this rank's weight and summed in place
                   by one ``all_reduce``: per GPU 2(w-1)/w of the payload crosses NVLink instead
                   of (w-1) payloads, and nothing is re-read locally (cfg3: 134 MB of weight
                   gradients per rank — 0.24 GB instead of 0.94 GB received at 8 GPUs)."""
    return "allreduce" if world > 1 and total_elems >= ALLREDUCE_MIN_ELEMS else "allgather"


class OutputCombiner:
    """Combine float32 outputs that are batch reductions across the ranks."""

    def __init__(self, world: int, mode: str = "mean", rows_per_rank=None, rank=None, plan=None):
        import torch.distributed as dist

        self.world = world
        self.dist = dist
        self.mode = mode
        self.rows = list(rows_per_rank) if rows_per_rank is not None else [1] * world
        self.rank = rank
        self.plan = plan  # None: exchange_plan() decides from the payload size
        self.layout = None
        self._kern = None

    def __call__(self, outs):
        import torch

        from .runtime import kernels as K
        from .runtime.device import DeviceArray

        outs = [o if isinstance(o, DeviceArray) else DeviceArray.from_numpy(np.asarray(o, "float32"))
                for o in outs]
        if self.layout is None:
            for o in outs:
                if o.dtype != np.float32:
                    raise TypeError("OutputCombiner handles float32 outputs")
            self.layout = pack_layout([o.shape for o in outs])
            w = combine_weights(self.mode, self.rows)
            self._weights = DeviceArray.from_numpy(np.asarray(w, "float32").reshape(self.world, 1))
            dt = "float32"
            # weighted sum over ranks: Elemwise mul (broadcast weights) then CAReduce add over axis 0
            self._mul = K.ElemwiseKernel.get({
                "inputs": [dt, dt], "out_dtypes": [dt], "outputs": ["t0"], "name": "shard_weight",
                "stmts": [{"op": "mul", "args": ["i0", "i1"], "dtype": dt, "in_dtypes": [dt, dt]}]})
            self._sum = K.CAReduceKernel.get("add", dt, "float64", dt)
        L = self.layout
        flat = DeviceArray.empty((L.total,), "float32")
        for o, off in zip(outs, L.offsets):
            n = o.size
            K.copy_into(flat.view(o.shape, _c(o.shape), off), o)
            del n
        t_flat = flat.owner.view(torch.float32)[: L.total]
        plan = self.plan or exchange_plan(L.total, self.world)
        if plan == "allreduce":
            rank = self.dist.get_rank() if self.rank is None else self.rank
            mine = self._weights.index((slice(rank, rank + 1),)).reshape_view((1,))
            self._mul.launch(flat.shape, [flat, mine], [flat])
            self.dist.all_reduce(t_flat, op=self.dist.ReduceOp.SUM)
            return [flat.view(s, _c(s), off) for s, off in zip(L.shapes, L.offsets)]
        gathered = DeviceArray.empty((self.world, L.total), "float32")
        t_all = gathered.owner.view(torch.float32)[: self.world * L.total]
        self.dist.all_gather_into_tensor(t_all, t_flat)
        self._mul.launch(gathered.shape, [gathered, self._weights], [gathered])
        red = self._sum.launch(gathered, (0,))
        return [red.view(s, _c(s), off) for s, off in zip(L.shapes, L.offsets)]


def _c(shape):
    from .runtime.device import c_strides

    return c_strides(shape)
