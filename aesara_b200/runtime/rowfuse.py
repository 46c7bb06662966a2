"""Row-region fusion: run ``Gemv(X, w)`` -> per-row ``Elemwise`` nodes -> ``Sum`` over the
rows / ``Gemv(X.T, r)`` as ONE pass over ``X`` (``codegen/rowfuse.py``).

The reference executes the logistic-regression gradient (BASELINE config 5, SURVEY.md
App. A.5: ``Gemv{inplace}``, three fused ``Elemwise``, two ``Sum``, a second ``Gemv`` on
``X.T``) node by node, so the [N, D] matrix is read twice.  The optimised graph is not
changed here: ``detect`` finds the region in the lowered program, the executor skips its
nodes and calls ``RowFusion.run`` at the position of the region's last node.  If the operands
at run time are not what the kernel handles (layout, dtype, D > 1024, a per-row vector of
the wrong length), the executor runs the region's nodes one by one as usual — still on the
device.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from ..codegen.rowfuse import THREADS, WARPS, rowfuse_source
from . import lib as _lib
from .device import DeviceArray, stream_handle

MAX_J = 8  # D <= 1024: the row slice, and the X.T @ r accumulator, live in registers


def _const_scalar(program, vid):
    v = program.vars[vid]
    if v.const is None or np.size(v.const) != 1:
        return None
    return float(np.asarray(v.const).reshape(()))


class RowFusion:
    def __init__(self, program, g1):
        self.program = program
        self.g1 = g1
        self.members = [g1]
        self.steps = []       # (node index, [ref per Elemwise input])
        self.ext = []         # external variable ids (per-row vectors or row-invariant scalars)
        self.sums = []        # (node index, ref)
        self.g2 = None        # (node index, ref)
        self._modules = {}

    # ------------------------------------------------------------------ analysis
    @staticmethod
    def detect(program, destroys=None):
        """All fusable regions of ``program`` (normally zero or one).  ``destroys[i]``: input
        positions node i rewrites in place (the executor's destroy map)."""
        if os.environ.get("AB_NO_ROWFUSE"):
            return []
        producer = {}
        for i, n in enumerate(program.nodes):
            for v in n.outputs:
                producer[v] = i
        consumers = {}
        for i, n in enumerate(program.nodes):
            for v in n.inputs:
                consumers.setdefault(v, []).append(i)
        found, taken = [], set()
        for i, n in enumerate(program.nodes):
            if n.op != "Gemv" or i in taken:
                continue
            f = RowFusion._grow(program, i, producer, consumers)
            if f is not None and not RowFusion._schedule_ok(program, f, consumers, destroys):
                f = None
            if f is not None and not (set(f.members) & taken):
                found.append(f)
                taken.update(f.members)
        return found

    @staticmethod
    def _schedule_ok(program, f, consumers, destroys):
        """The members are deferred to ``f.last``: no node in between may read a value the
        region produces (it would not exist yet), nor rewrite memory in place (X, w or an
        operand vector could change under the deferred kernel) — ADVICE r1."""
        members = set(f.members)
        outs = {v for m in f.members for v in program.nodes[m].outputs}
        for i in range(f.first + 1, f.last):
            if i in members:
                continue
            if any(v in outs for v in program.nodes[i].inputs):
                return False
            if destroys is not None and destroys[i]:
                return False
        return True

    @staticmethod
    def _is_fresh_gemv(program, node, producer):
        """``Gemv(AllocEmpty, alpha, A, x, 0)`` on float32 with constant alpha."""
        y, alpha, A, x, beta = node.inputs
        if program.vars[node.outputs[0]].dtype != "float32":
            return False
        py = producer.get(y)
        if py is None or program.nodes[py].op != "AllocEmpty":
            return False
        if _const_scalar(program, beta) != 0.0 or _const_scalar(program, alpha) is None:
            return False
        return True

    @staticmethod
    def _grow(program, g1, producer, consumers):
        nodes = program.nodes
        n1 = nodes[g1]
        if not RowFusion._is_fresh_gemv(program, n1, producer):
            return None
        X = n1.inputs[2]
        if program.vars[X].ndim != 2 or program.vars[X].dtype != "float32":
            return None
        px = producer.get(X)
        if px is not None and nodes[px].op == "DimShuffle":
            return None  # Gemv on a transposed view: the column pattern, not a row region
        f = RowFusion(program, g1)
        f.X, f.w = X, n1.inputs[3]
        f.alpha1 = _const_scalar(program, n1.inputs[1])
        rowvals = {n1.outputs[0]: ("z",)}
        ext_index = {}

        def ext_ref(v):
            if v not in ext_index:
                ext_index[v] = len(f.ext)
                f.ext.append(v)
            return ("ext", ext_index[v])

        for i in range(g1 + 1, len(nodes)):
            n = nodes[i]
            if not any(v in rowvals for v in n.inputs):
                continue
            if n.op == "Elemwise" and len(n.outputs) == 1 and program.vars[n.outputs[0]].ndim == 1:
                refs = []
                for v in n.inputs:
                    var = program.vars[v]
                    if v in rowvals:
                        refs.append(rowvals[v])
                    elif var.kind == "tensor" and var.ndim <= 1:
                        refs.append(ext_ref(v))
                    else:
                        return None
                rowvals[n.outputs[0]] = ("val", len(f.steps))
                f.steps.append((i, refs))
                f.members.append(i)
            elif (n.op == "CAReduce" and n.params.get("scalar_op") == "add"
                  and list(n.params.get("axis") or [0]) == [0] and n.params.get("acc_dtype") == "float64"):
                f.sums.append((i, rowvals[n.inputs[0]]))
                f.members.append(i)
            elif n.op == "Gemv" and f.g2 is None and n.inputs[3] in rowvals:
                A = n.inputs[2]
                pa = producer.get(A)
                if (pa is None or nodes[pa].op != "DimShuffle" or nodes[pa].inputs[0] != X
                        or list(nodes[pa].params.get("new_order", [])) != [1, 0]
                        or not RowFusion._is_fresh_gemv(program, n, producer)):
                    return None
                f.g2 = (i, rowvals[n.inputs[3]])
                f.alpha2 = _const_scalar(program, n.inputs[1])
                f.members.append(i)
            else:
                return None  # a per-row value escapes the region
        if f.g2 is None or not f.steps:
            return None
        if any(v in program.outputs for v in rowvals):
            return None
        # no external operand may depend on something the region computes
        region_out = {nodes[i].outputs[0] for i in f.members}
        tainted = set(region_out)
        for i, n in enumerate(nodes):
            if i in f.members:
                continue
            if any(v in tainted for v in n.inputs):
                tainted.update(n.outputs)
        if any(v in tainted for v in f.ext):
            return None
        f.members.sort()
        f.first, f.last = f.members[0], f.members[-1]
        # every external operand must exist when the region runs
        for v in f.ext:
            p = producer.get(v)
            if p is not None and p > f.last:
                return None
        return f

    # ------------------------------------------------------------------ execution
    def _module(self, J, classes, by_ptr):
        key = (J, classes, by_ptr)
        m = self._modules.get(key)
        if m is None:
            prog = self.program
            rows = [k for k, c in enumerate(classes) if c == "row"]
            scals = [k for k, c in enumerate(classes) if c == "scal"]
            pos = {k: ("row", rows.index(k)) if c == "row" else ("scal", scals.index(k))
                   for k, c in enumerate(classes)}

            def conv(ref):
                return pos[ref[1]] if ref[0] == "ext" else ref

            spec = {
                "J": J,
                "steps": [{"expr": prog.nodes[i].params["expr"], "args": [conv(r) for r in refs]}
                          for i, refs in self.steps],
                "row_dtypes": [prog.vars[self.ext[k]].dtype for k in rows],
                "scal_dtypes": [prog.vars[self.ext[k]].dtype for k in scals],
                "scal_by_ptr": [by_ptr[k] for k in scals],
                "sums": [conv(r) for _, r in self.sums],
                "gemv2": conv(self.g2[1]),
                "rows_per_iter": int(os.environ.get("AB_ROWFUSE_ROWS", "1")),
            }
            src = rowfuse_source(spec)
            m = self._modules[key] = {"src": src, "rows": rows, "scals": scals, "handle": None}
        return m

    def compile_all(self):
        """JIT the common variant (every external vector per-row except size-1 constants)."""
        prog = self.program
        classes = tuple("scal" if (prog.vars[v].static_shape or (None,))[-1:] == (1,) or prog.vars[v].ndim == 0
                        else "row" for v in self.ext)
        m = self._module(4, classes, tuple(False for _ in self.ext))
        _lib.compile_cubin(m["src"], "rowfuse")
        return 1

    def run(self, ex, env):
        """Execute the region; returns False (nothing done) if the operands do not fit."""
        prog = self.program
        X, w = env.get(self.X), env.get(self.w)
        if not isinstance(X, DeviceArray) or X.dtype != np.float32 or X.ndim != 2:
            return False
        N, D = X.shape
        if N == 0 or D == 0 or D % 4 or X.strides[1] != 1 or X.strides[0] % 4 or X.ptr % 16:
            return False
        J = (D + 127) // 128
        if J > MAX_J:
            return False
        if not isinstance(w, DeviceArray):
            w = ex.dev(w, dtype="float32")
        if w.dtype != np.float32 or w.shape != (D,):
            return False
        classes, by_ptr, vals = [], [], []
        for v in self.ext:
            a = env[v]
            dt = np.dtype(prog.vars[v].dtype)
            if isinstance(a, DeviceArray):
                if a.dtype != dt:
                    return False
                if a.size == 1:
                    classes.append("scal"); by_ptr.append(True); vals.append(a)
                elif a.ndim == 1 and a.shape[0] == N:
                    classes.append("row"); by_ptr.append(False); vals.append(a)
                else:
                    return False
            else:
                h = np.asarray(a)
                if h.size != 1:
                    if h.ndim == 1 and h.shape[0] == N:
                        classes.append("row"); by_ptr.append(False)
                        vals.append(ex.dev(h, dtype=dt.name))
                        continue
                    return False
                classes.append("scal"); by_ptr.append(False); vals.append(h.astype(dt).reshape(()))
        m = self._module(J, tuple(classes), tuple(by_ptr))
        if m["handle"] is None:
            m["handle"] = _lib.load_module(m["src"], "rowfuse")
        lib = _lib.load()
        dev = X.owner.device
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        rows_per_cta = WARPS * int(os.environ.get("AB_ROWFUSE_ROWS", "1"))
        grid = int(max(1, min((N + rows_per_cta - 1) // rows_per_cta, sms * 6)))
        Dp = J * 128
        gw_part = DeviceArray.empty((grid, Dp), "float32")
        sum_parts = [DeviceArray.empty((grid,), "float64") for _ in self.sums]

        keep = []

        def arg(ctype_value):
            keep.append(ctype_value)
            return C.cast(C.pointer(ctype_value), C.c_void_p)

        args = [arg(C.c_void_p(X.ptr)), arg(C.c_longlong(N)), arg(C.c_int(D)), arg(C.c_longlong(X.strides[0])),
                arg(C.c_void_p(w.ptr)), arg(C.c_longlong(w.strides[0])), arg(C.c_float(self.alpha1))]
        for k in m["rows"]:
            a = vals[k]
            args += [arg(C.c_void_p(a.ptr)), arg(C.c_longlong(a.strides[0]))]
        for k in m["scals"]:
            if by_ptr[k]:
                args.append(arg(C.c_void_p(vals[k].ptr)))
            else:
                args.append(arg(_CT[np.dtype(prog.vars[self.ext[k]].dtype).name](vals[k].item())))
        args += [arg(C.c_float(self.alpha2)), arg(C.c_void_p(gw_part.ptr))]
        for p in sum_parts:
            args.append(arg(C.c_void_p(p.ptr)))
        argv = (C.c_void_p * len(args))(*args)
        _lib.check(lib.ab_kernel_launch(m["handle"], b"ab_rowfused", grid, THREADS, 0, argv, stream_handle()))

        from . import kernels as K

        red32 = K.CAReduceKernel.get("add", "float32", "float64", "float32")
        gw = red32.launch(gw_part, [0])                     # [Dp] <- sum over CTAs
        env[prog.nodes[self.g2[0]].outputs[0]] = gw.index((slice(0, D),))
        for (i, _), part in zip(self.sums, sum_parts):
            node = prog.nodes[i]
            red = K.CAReduceKernel.get("add", "float64", "float64", node.params["out_dtype"])
            env[node.outputs[0]] = red.launch(part, [0])
        return True


_CT = {
    "float32": C.c_float, "float64": C.c_double, "int8": C.c_int8, "int16": C.c_int16,
    "int32": C.c_int32, "int64": C.c_int64, "uint8": C.c_uint8, "uint16": C.c_uint16,
    "uint32": C.c_uint32, "uint64": C.c_uint64, "bool": C.c_bool,
}
