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

import os
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


# ---------------------------------------------------------------------------------------------
# Sharding behind the linker: ``aesara_b200.mode(shard="rows")``
# ---------------------------------------------------------------------------------------------
SMALL_OUTPUT_BYTES = 1 << 16  # outputs below this are packed into one exchange at the end


class ShardedExecutor:
    """One rank of a row-sharded evaluation.

    ``plan`` (``shardplan.analyse`` / ``infer_sharded_inputs``) says which function inputs are
    this rank's row block and how each output combines — decided from the graph, SURVEY §8e's
    rule, not by the caller.  The wrapped executor runs the unchanged program on the local
    rows; batch reductions (``sum`` / ``mean`` outputs: cost, gradients) are all-reduced over
    NCCL, each rank's contribution weighted by its share of the rows for ``mean``.  A large
    reduction output is handed to NCCL on a side stream as soon as the node that produces it
    has been launched (``ProgramExecutor.output_hook``), so the exchange of an early gradient
    overlaps the rest of the backward pass; the small ones travel packed in one buffer at the
    end.  ``concat`` outputs stay sharded (this rank's rows) unless ``gather=True`` asks for the
    north star's all-gather."""

    def __init__(self, executor, plan, group=None, gather=False, overlap=True):
        import torch.distributed as dist

        self.ex = executor
        self.plan = plan
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.gather = gather
        self.overlap = overlap
        prog = executor.program
        consumed = {v for n in prog.nodes for v in n.inputs}
        # outputs nobody else reads may be reduced in place as soon as they exist
        self._early_ok = [prog.outputs[k] not in consumed and prog.outputs[k] not in prog.inputs
                          and prog.outputs.count(prog.outputs[k]) == 1 for k in range(len(prog.outputs))]
        self._side = None
        self._mul = None
        self.early_issued = 0      # collectives issued before the evaluation finished (last call)
        self.exchanges = 0         # collectives of the last call

    # forwarded attributes the VM / profiler / debugger consult
    def __getattr__(self, name):
        return getattr(self.ex, name)

    @property
    def time_nodes(self):
        return self.ex.time_nodes

    @time_nodes.setter
    def time_nodes(self, v):
        self.ex.time_nodes = v

    @property
    def trace(self):
        return self.ex.trace

    @trace.setter
    def trace(self, v):
        self.ex.trace = v

    def _weight(self, args):
        """Device scalar n_r / N (float32, shape [1]) for the ``mean`` outputs, computed on the
        device: a one-element all-reduce of the local row count, no host synchronisation."""
        import torch

        from .runtime.device import DeviceArray

        n_local = None
        for a, ax in zip(args, self.plan.sharded_inputs):
            if ax is not None:
                n_local = int(np.shape(a)[ax] if not isinstance(a, DeviceArray) else a.shape[ax])
                break
        if n_local is None:
            raise ValueError("sharded evaluation without a sharded input")
        t = torch.full((1,), float(n_local), dtype=torch.float64, device="cuda")
        tot = t.clone()
        self.dist.all_reduce(tot, op=self.dist.ReduceOp.SUM, group=self.group)
        w = (t / tot).to(torch.float32)
        return DeviceArray.from_torch(w)

    def _scale_inplace(self, arr, w):
        from .runtime import kernels as K

        if self._mul is None:
            dt = "float32"
            self._mul = K.ElemwiseKernel.get({
                "inputs": [dt, dt], "out_dtypes": [dt], "outputs": ["t0"], "name": "shard_weight",
                "stmts": [{"op": "mul", "args": ["i0", "i1"], "dtype": dt, "in_dtypes": [dt, dt]}]})
        wv = w.view((1,) * arr.ndim, (0,) * arr.ndim) if arr.ndim else w.view((), ())
        self._mul.launch(arr.shape, [arr, wv], [arr])

    def __call__(self, *args, output_subset=None):
        import torch

        from .runtime import kernels as K
        from .runtime.device import DeviceArray

        if self.world == 1:
            return self.ex(*args, output_subset=output_subset)
        modes = self.plan.outputs
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        side = self._side
        need_w = any(m[0] == "mean" for m in modes)
        side.wait_stream(cur)
        w = w_ready = None
        if need_w:
            with torch.cuda.stream(side):
                w = self._weight(args)
                w_ready = torch.cuda.Event()
                w_ready.record(side)
        works, done = [], set()
        self.early_issued = self.exchanges = 0

        def reduce_now(k, val):
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                side.wait_event(ev)
                if modes[k][0] == "mean":
                    self._scale_inplace(val, w)
                t = val.owner.view(torch.float32)[: val.size]
                works.append((self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group,
                                                   async_op=True), val))
            self.exchanges += 1

        def hook(k, val):
            if (modes[k][0] in ("sum", "mean") and self._early_ok[k] and isinstance(val, DeviceArray)
                    and val.dtype == np.float32 and val.is_c_contiguous()
                    and val.nbytes >= SMALL_OUTPUT_BYTES and (output_subset is None or k in output_subset)):
                reduce_now(k, val)
                done.add(k)
                self.early_issued += 1

        host_out, self.ex.host_outputs = self.ex.host_outputs, False
        self.ex.output_hook = hook if self.overlap else None
        try:
            outs = list(self.ex(*args, output_subset=output_subset))
        finally:
            self.ex.output_hook = None
            self.ex.host_outputs = host_out
        # the rest: large ones individually, small ones packed into one buffer
        small = []
        for k, (o, m) in enumerate(zip(outs, modes)):
            if o is None or k in done or m[0] not in ("sum", "mean"):
                continue
            if not isinstance(o, DeviceArray):
                o = DeviceArray.from_numpy(np.asarray(o))
            if o.dtype != np.float32:
                raise TypeError(f"output {k}: batch reductions are combined in float32, got {o.dtype}")
            if not o.is_c_contiguous() or not self._early_ok[k]:
                o = K.contiguous_copy(o)
            outs[k] = o
            if o.nbytes >= SMALL_OUTPUT_BYTES:
                reduce_now(k, o)
            else:
                small.append(k)
        if small:
            layout = pack_layout([outs[k].shape for k in small])
            flat = DeviceArray.empty((layout.total,), "float32")
            for k, off in zip(small, layout.offsets):
                o = outs[k]
                seg = flat.view(o.shape, _c(o.shape), off)
                K.copy_into(seg, o)
                if modes[k][0] == "mean":
                    if w_ready is not None:
                        cur.wait_event(w_ready)  # only the weight, not the exchanges queued behind it
                        w_ready = None
                    self._scale_inplace(seg, w)
                outs[k] = seg
            t = flat.owner.view(torch.float32)[: layout.total]
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            self.exchanges += 1
        if self.gather:
            for k, (o, m) in enumerate(zip(outs, modes)):
                if o is not None and m[0] == "concat":
                    outs[k] = self._all_gather_rows(o, m[1])
        for work, _val in works:
            work.wait()      # the current stream waits for the collective
        cur.wait_stream(side)
        if host_out:
            got = DeviceArray.download_all([o for o in outs if o is not None])
            it = iter(got)
            outs = [None if o is None else next(it) for o in outs]
        return outs

    def _all_gather_rows(self, o, axis):
        """Equal row blocks only (``all_gather_into_tensor``); the rows move to axis 0 and back."""
        import torch

        from .runtime import kernels as K
        from .runtime.device import DeviceArray

        if axis != 0:
            order = [axis] + [d for d in range(o.ndim) if d != axis]
            o = o.dimshuffle(order)
        o = o if o.is_c_contiguous() else K.contiguous_copy(o)
        out = DeviceArray.empty((self.world * o.shape[0],) + o.shape[1:], o.dtype)
        nb = o.nbytes
        self.dist.all_gather_into_tensor(out.owner[: self.world * nb], o.owner[:nb], group=self.group)
        self.exchanges += 1
        if axis != 0:
            inv = [0] * out.ndim
            for pos, d in enumerate([axis] + [d for d in range(out.ndim) if d != axis]):
                inv[d] = pos
            out = out.dimshuffle(inv)
        return out


# ---------------------------------------------------------------------------------------------
# Row-chunked evaluation of host inputs: the same proof, used over time instead of over GPUs
# ---------------------------------------------------------------------------------------------
class ChunkedHostExecutor:
    """Evaluate a batch-map program on host (page-locked) inputs in ``n_chunks`` row blocks.

    A user of the reference hands NumPy arrays to ``Function.__call__``; for BASELINE config 3
    that is 2.15 GB per call over PCIe (~43 ms at 53 GB/s) in front of 12 ms of kernels.  The
    shardability analysis that allows splitting the rows over GPUs (``shardplan``) equally
    allows splitting them over *time*: row block i+1 is uploaded on the copy stream while
    block i is evaluated, and the outputs are combined exactly as the plan says (``mean``
    outputs weighted by the block's share of the rows and accumulated, ``concat`` outputs
    written into their rows).  Device-resident arguments are evaluated in one piece."""

    def __init__(self, executor, plan, n_chunks=8, min_rows=4096):
        self.ex = executor
        self.plan = plan
        self.n_chunks = int(n_chunks)
        self.min_rows = int(min_rows)
        self._fma = None
        self._mul = None
        self.chunks_run = 0

    def __getattr__(self, name):
        return getattr(self.ex, name)

    @property
    def time_nodes(self):
        return self.ex.time_nodes

    @time_nodes.setter
    def time_nodes(self, v):
        self.ex.time_nodes = v

    @property
    def trace(self):
        return self.ex.trace

    @trace.setter
    def trace(self, v):
        self.ex.trace = v

    def _kernels(self):
        from .runtime import kernels as K

        if self._fma is None:
            dt = "float32"
            self._mul = K.ElemwiseKernel.get({
                "inputs": [dt, dt], "out_dtypes": [dt], "outputs": ["t0"], "name": "chunk_weight",
                "stmts": [{"op": "mul", "args": ["i0", "i1"], "dtype": dt, "in_dtypes": [dt, dt]}]})
            self._fma = K.ElemwiseKernel.get({
                "inputs": [dt, dt, dt], "out_dtypes": [dt], "outputs": ["t1"], "name": "chunk_accumulate",
                "stmts": [{"op": "mul", "args": ["i1", "i2"], "dtype": dt, "in_dtypes": [dt, dt]},
                          {"op": "add", "args": ["i0", "t0"], "dtype": dt, "in_dtypes": [dt, dt]}]})
        return self._mul, self._fma

    def __call__(self, *args, output_subset=None):
        from .runtime import kernels as K
        from .runtime.device import DeviceArray

        axes = self.plan.sharded_inputs
        rows = None
        chunkable = output_subset is None and self.ex.trace is None
        for a, ax in zip(args, axes):
            if ax is None:
                continue
            if isinstance(a, DeviceArray) or not isinstance(a, np.ndarray) or ax != 0 or not a.flags.c_contiguous:
                chunkable = False
                break
            if rows is None:
                rows = a.shape[0]
            elif a.shape[0] != rows:
                chunkable = False  # run-time broadcasting along the batch axis
        n = self.n_chunks
        if not chunkable or rows is None or rows < 2 * self.min_rows or n < 2:
            self.chunks_run = 1
            return self.ex(*args, output_subset=output_subset)
        n = max(2, min(n, rows // self.min_rows))
        modes = self.plan.outputs
        host_out, self.ex.host_outputs = self.ex.host_outputs, False
        acc = [None] * len(modes)
        mul, fma = self._kernels()
        # every block's weight (and a 1 for `sum` outputs) goes to the device in ONE small copy
        # before anything is queued: a pageable-memory copy is synchronous, and issued between
        # blocks it would wait for the previous block's kernels — no overlap at all
        blocks = [row_block(rows, n, c) for c in range(n)]
        # replicated host tensors (weights) cross the bus ONCE, not once per block
        import torch

        cur = torch.cuda.current_stream()
        args = list(args)
        keep_alive = []
        for k, (a, ax) in enumerate(zip(args, axes)):
            if ax is None and isinstance(a, np.ndarray) and a.size > 64:
                for ev, _ in self.ex._inflight:     # sources of the previous call's uploads
                    ev.synchronize()
                self.ex._inflight = []
                d, tok = DeviceArray.upload(a, self.ex._copy_stream_for(cur), cur)
                if tok is not None:
                    cur.wait_event(tok[0])
                    keep_alive.append(tok)
                args[k] = d
        wall = DeviceArray.from_numpy(np.asarray([(b0 - a0) / float(rows) for a0, b0 in blocks] + [1.0], "float32"))
        one = wall.index((slice(n, n + 1),))
        trace = [] if os.environ.get("AB_CHUNK_TRACE") else None
        if trace is not None:
            import time

            import torch

            t_origin = time.perf_counter()
            e_origin = torch.cuda.Event(enable_timing=True)
            e_origin.record()
        try:
            for c in range(n):
                a0, b0 = blocks[c]
                local = [a[a0:b0] if ax == 0 else a for a, ax in zip(args, axes)]
                if trace is not None:
                    t0 = time.perf_counter() - t_origin
                outs = self.ex(*local)
                if trace is not None:
                    e_done = torch.cuda.Event(enable_timing=True)
                    e_done.record()
                    ups = [ev for ev, _ in self.ex._inflight]
                    trace.append((c, t0, time.perf_counter() - t_origin, ups, e_done))
                w = wall.index((slice(c, c + 1),))
                for k, (o, m) in enumerate(zip(outs, modes)):
                    if m[0] in ("sum", "mean"):
                        o = o if isinstance(o, DeviceArray) else DeviceArray.from_numpy(np.asarray(o))
                        if o.dtype != np.float32:
                            raise TypeError(f"output {k}: batch reductions are combined in float32")
                        wv = (w if m[0] == "mean" else None)
                        if acc[k] is None:
                            if wv is None:
                                acc[k] = o if o.is_c_contiguous() else K.contiguous_copy(o)
                            else:
                                acc[k] = DeviceArray.empty(o.shape, "float32")
                                mul.launch(o.shape, [o, wv.view((1,) * o.ndim, (0,) * o.ndim)], [acc[k]])
                        else:
                            sc = wv if wv is not None else one
                            fma.launch(o.shape, [acc[k], o, sc.view((1,) * o.ndim, (0,) * o.ndim)], [acc[k]])
                    elif m[0] == "concat":
                        ax = m[1]
                        if acc[k] is None:
                            shape = list(o.shape)
                            shape[ax] = rows
                            acc[k] = DeviceArray.empty(shape, o.dtype)
                        sl = [slice(None)] * o.ndim
                        sl[ax] = slice(a0, b0)
                        K.copy_into(acc[k].index(tuple(sl)), o)
                    elif acc[k] is None:
                        acc[k] = o
            self.chunks_run = n
            for ev, _ in keep_alive:  # the page-locked sources stay alive until the copies ran
                ev.synchronize()
            if trace is not None:
                torch.cuda.synchronize()
                print("[chunk trace] block: host enters / host leaves the evaluation (ms) | device: compute done (ms)")
                for c, t0, t1, ups, e_done in trace:
                    print(f"  {c:2d}: {1e3 * t0:7.2f} / {1e3 * t1:7.2f} | {e_origin.elapsed_time(e_done):7.2f}")
        finally:
            self.ex.host_outputs = host_out
        if host_out:
            return DeviceArray.download_all(acc)
        return acc
