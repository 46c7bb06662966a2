"""Recognise the Scan inner graphs of the family

    G      = Gemm(x_t, 1, s_hs, U, 1)                              # [B, gates*H]
    s'_k   = Elemwise nodes over G[:, g*H:(g+1)*H], s_0 .. s_{S-1}   # [B, H] each

and run the whole loop as one persistent kernel (``csrc/ab_scan_cell_kernel.cuh``).

The reference's Scan drives *any* inner function from a host loop
(``aesara/scan/op.py:1799-2103``); the general device loop is ``runtime/scan.py``.  This module
is the fast path for recurrences whose only contraction is one product with a loop-invariant
matrix: the LSTM cell BASELINE config 4 names (4 gates, states h and c — compiled ahead of time,
``ab_lstm_scan``), a tanh-RNN (1 gate, 1 state), minimal gated units (2-3 gates) and so on.
The inner program is traced with the actual operand shapes on address-only stand-ins — host
nodes and view nodes run for real, so the column ranges are whatever the graph computes, not
what their printed names suggest — and the Elemwise nodes become the kernel's cell
(``codegen/scan_cell.py``).  Anything else keeps the general loop.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from ..codegen.scan_cell import merge_cell, scan_cell_source
from . import lib as _lib
from .device import DeviceArray, c_strides, stream_handle

_FAKE_BASE = 1 << 44  # address-only stand-ins never dereferenced
MAX_GATES, MAX_STATES = 4, 3


def _canon(expr, roles):
    """Canonical string of a single-output scalar expression with named inputs;
    operands of commutative ops are sorted."""
    memo = {}

    def ref(r):
        if isinstance(r, dict):
            return f"const({r['const']!r})"
        if r[0] == "i":
            return roles[int(r[1:])]
        k = int(r[1:])
        if k not in memo:
            st = expr["stmts"][k]
            args = [ref(a) for a in st["args"]]
            if st["op"] in ("add", "mul"):
                args.sort()
            memo[k] = f"{st['op']}({','.join(args)})"
        return memo[k]

    return ref(expr["outputs"][0])


_C_NEW = "add(mul(c,sigmoid(gf)),mul(sigmoid(gi),tanh(gg)))"
_H_NEW = "mul(sigmoid(go),tanh(cn))"


class CellSpec:
    def __init__(self, gates, states, hs, steps, outputs, is_lstm):
        self.gates, self.states, self.hs = gates, states, hs
        self.steps, self.outputs, self.is_lstm = steps, outputs, is_lstm
        self._src = None
        self._handle = None

    def source(self):
        if self._src is None:
            name = "+".join(st[0].get("name", "?") for st in self.steps)
            merged = merge_cell(self.steps, self.gates, self.states, self.outputs, name=name)
            self._src = scan_cell_source(merged, self.gates, self.states)
        return self._src

    def compile(self):
        return _lib.compile_cubin(self.source(), "scan_cell")

    def handle(self):
        if self._handle is None:
            self._handle = _lib.load_module(self.source(), "scan_cell")
        return self._handle


class CellMatch:
    def __init__(self, inner_executor, n_states):
        self.ex = inner_executor
        self.n_states = n_states
        self._cache = {}

    def match(self, B, H, gates=4):
        """-> CellSpec or None for operands x_t [B, gates*H], states [B, H], U [H, gates*H]."""
        key = (B, H, gates)
        if key not in self._cache:
            try:
                self._cache[key] = self._trace(B, H, gates)
            except Exception:
                self._cache[key] = None
        return self._cache[key]

    def _trace(self, B, H, G):
        from .vm import _EXEC

        ex = self.ex
        prog = ex.program
        S = self.n_states
        if not (1 <= G <= MAX_GATES and 1 <= S <= MAX_STATES):
            return None
        if len(prog.inputs) != S + 2 or len(prog.outputs) != S:
            return None
        for vid in prog.inputs + prog.outputs:
            v = prog.vars[vid]
            if v.kind != "tensor" or v.dtype != "float32" or v.ndim != 2:
                return None
        shapes = [(B, G * H)] + [(B, H)] * S + [(H, G * H)]
        env = dict(ex._const_host)
        state_base = {}
        for k, (vid, shp) in enumerate(zip(prog.inputs, shapes)):
            base = _FAKE_BASE * (k + 1)
            if 1 <= k <= S:
                state_base[base] = k - 1
            env[vid] = DeviceArray(None, base, "float32", shp, c_strides(shp))
        next_base = [_FAKE_BASE * 16]

        def fake(shape):
            b = next_base[0]
            next_base[0] += _FAKE_BASE
            return DeviceArray(None, b, "float32", shape, c_strides(shape))

        gemm_out, hs, g_base = None, None, None
        elemwise = []          # (node, args, out)
        val_of = {}            # fake ptr -> step index
        for i, node in enumerate(prog.nodes):
            args = [env[v] for v in node.inputs]
            if node.op == "Gemm":
                if gemm_out is not None:
                    return None
                z, a, x, y, b = args
                if node.inputs[0] != prog.inputs[0] or node.inputs[3] != prog.inputs[S + 1]:
                    return None
                if node.inputs[2] not in prog.inputs[1 : S + 1]:
                    return None
                hs = prog.inputs.index(node.inputs[2]) - 1
                if not (np.ndim(a) == 0 and np.ndim(b) == 0 and float(np.asarray(a)) == 1.0
                        and float(np.asarray(b)) == 1.0):
                    return None
                gemm_out = env[node.outputs[0]] = fake((B, G * H))
                g_base = gemm_out.ptr
                continue
            dev_in = [a for a in args if isinstance(a, DeviceArray)]
            if node.op == "Elemwise" and dev_in:
                if len(node.outputs) != 1:
                    return None
                out = fake((B, H))
                val_of[out.ptr] = len(elemwise)
                elemwise.append((node, args, out))
                env[node.outputs[0]] = out
                continue
            if node.op in ("Shape_i", "Shape", "ScalarFromTensor", "TensorFromScalar", "ScalarOp",
                           "MakeVector", "Subtensor", "DimShuffle", "View", "Elemwise", "Assert"):
                outs = _EXEC[node.op](ex, i, node, args)
                if len(node.outputs) == 1:
                    env[node.outputs[0]] = outs
                else:
                    for vid, o in zip(node.outputs, outs):
                        env[vid] = o
                continue
            return None
        if gemm_out is None or not elemwise:
            return None

        def role(a):
            if not isinstance(a, DeviceArray):
                h = np.asarray(a)
                if h.size == 1 and h.dtype.kind == "f":
                    return ("const", float(h.reshape(-1)[0]))
                return None
            if a.shape != (B, H):
                return None
            if g_base <= a.ptr < g_base + G * H * 4 and a.strides == (G * H, 1):
                off = (a.ptr - g_base) // 4
                if off % H:
                    return None
                return ("gate", off // H)
            if a.ptr in state_base and a.strides == (H, 1):
                return ("state", state_base[a.ptr])
            if a.ptr in val_of and a.strides == (H, 1):
                return ("val", val_of[a.ptr])
            return None

        steps = []
        for node, args, _out in elemwise:
            refs = [role(a) for a in args]
            if None in refs:
                return None
            steps.append((node.params["expr"], refs))
        outputs = []
        for vid in prog.outputs:
            v = env.get(vid)
            if not isinstance(v, DeviceArray) or v.ptr not in val_of or v.strides != (H, 1):
                return None
            outputs.append(("val", val_of[v.ptr]))
        # every state must be used through the Gemm or the cell, and nothing may read x_t but the Gemm
        is_lstm = False
        if G == 4 and S == 2 and hs == 0 and len(steps) == 2 and outputs == [("val", 1), ("val", 0)]:
            names = {("gate", 0): "gi", ("gate", 1): "gf", ("gate", 2): "go", ("gate", 3): "gg",
                     ("state", 1): "c", ("val", 0): "cn"}
            try:
                r1 = [names[r] for r in steps[0][1]]
                r2 = [names[r] for r in steps[1][1]]
                is_lstm = (_canon(steps[0][0], r1) == _C_NEW and _canon(steps[1][0], r2) == _H_NEW)
            except KeyError:
                is_lstm = False
        return CellSpec(G, S, hs, steps, outputs, is_lstm)


def run_lstm(T, x, U, hbuf, cbuf, pos_h, pos_c):
    """The ahead-of-time LSTM member.  x: [>=T, B, 4H]; U: [H, 4H]; hbuf/cbuf: rings."""
    lib = _lib.load()
    B, H = hbuf.shape[1], hbuf.shape[2]
    nbytes = C.c_size_t()
    _lib.check(lib.ab_lstm_scan_workspace_bytes(B, H, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=hbuf.owner.device)
    _lib.check(lib.ab_lstm_scan(T, B, H, x.ptr, x.strides[0], x.strides[1], U.ptr, U.strides[0],
                                U.strides[1], hbuf.ptr, hbuf.shape[0], pos_h, cbuf.ptr,
                                cbuf.shape[0], pos_c, ws.data_ptr(), nbytes.value, stream_handle()))


def run_cell(spec, T, x, U, bufs, pos):
    """A generated cell: ``bufs`` are the state rings in Scan order, ``pos`` their positions."""
    lib = _lib.load()
    B, H = bufs[0].shape[1], bufs[0].shape[2]
    S = spec.states
    nbytes = C.c_size_t()
    _lib.check(lib.ab_cell_scan_workspace_bytes(spec.gates, B, H, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=bufs[0].owner.device)
    ptrs = (C.c_void_p * S)(*[b.ptr for b in bufs])
    lens = (C.c_int64 * S)(*[b.shape[0] for b in bufs])
    poss = (C.c_int64 * S)(*[int(p) for p in pos])
    _lib.check(lib.ab_cell_scan(spec.handle(), spec.gates, S, spec.hs, T, B, H, x.ptr, x.strides[0],
                                x.strides[1], U.ptr, U.strides[0], U.strides[1], ptrs, lens, poss,
                                ws.data_ptr(), nbytes.value, stream_handle()))


def eligible(x, U, bufs, T, gates):
    lib = _lib.load()
    if any(a.dtype != np.float32 for a in [x, U] + list(bufs)):
        return False
    if x.ndim != 3 or U.ndim != 2 or any(b.ndim != 3 for b in bufs):
        return False
    B, H = bufs[0].shape[1], bufs[0].shape[2]
    if any(b.shape[1:] != (B, H) for b in bufs) or x.shape[1:] != (B, gates * H) or U.shape != (H, gates * H):
        return False
    if not all(b.is_c_contiguous() for b in bufs):
        return False
    if x.strides[2] != 1 or x.strides[1] % 4 or x.strides[0] % 4 or x.ptr % 16:
        return False
    if any(b.ptr % 16 for b in bufs):
        return False
    # tensor-core tiles only pay off above a minimum size; tiny problems keep the general loop
    if B * H < 128 * 64 or T < 2:
        return False
    return bool(lib.ab_cell_scan_supported(gates, len(bufs), T, B, H))
