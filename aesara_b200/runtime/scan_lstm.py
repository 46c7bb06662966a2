"""Recognise an LSTM-cell inner graph of a ``Scan`` and run it as one persistent
kernel (``csrc/ab_scan_lstm.cu``).

The reference's Scan drives *any* inner function from a host loop
(``aesara/scan/op.py:1799-2103``).  Here the general loop is ``runtime/scan.py``;
this module is the fast path for the recurrence BASELINE config 4 names.  The
inner program is accepted only if, traced with the actual
operand shapes, it is *exactly*

    G     = Gemm(x_t, 1, h_prev, U, 1)                      # [B, 4H]
    c_new = sigmoid(G[:, H:2H]) * c_prev + sigmoid(G[:, :H]) * tanh(G[:, 3H:])
    h_new = sigmoid(G[:, 2H:3H]) * tanh(c_new)

The trace executes the inner program's host nodes and view nodes for real
(shape arithmetic, ``Subtensor`` offsets) on address-only stand-ins, so the
column ranges are whatever the graph computes, not what their printed names
suggest.  Anything else falls back to the general device loop.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as _lib
from .device import DeviceArray, c_strides, stream_handle

_FAKE_BASE = 1 << 44  # address-only stand-ins never dereferenced


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


class LstmMatch:
    def __init__(self, inner_executor):
        self.ex = inner_executor
        self._cache = {}

    def match(self, B, H):
        key = (B, H)
        if key not in self._cache:
            try:
                self._cache[key] = self._trace(B, H)
            except Exception:
                self._cache[key] = False
        return self._cache[key]

    def _trace(self, B, H):
        from .vm import _EXEC

        ex = self.ex
        prog = ex.program
        if len(prog.inputs) != 4 or len(prog.outputs) != 2:
            return False
        for vid in prog.inputs + prog.outputs:
            v = prog.vars[vid]
            if v.kind != "tensor" or v.dtype != "float32" or v.ndim != 2:
                return False
        shapes = [(B, 4 * H), (B, H), (B, H), (H, 4 * H)]
        bases = {}
        env = dict(ex._const_host)
        for k, (vid, shp) in enumerate(zip(prog.inputs, shapes)):
            base = _FAKE_BASE * (k + 1)
            bases[base] = ("in", k)
            env[vid] = DeviceArray(None, base, "float32", shp, c_strides(shp))
        next_base = [_FAKE_BASE * 16]

        def fake(shape, tag):
            b = next_base[0]
            next_base[0] += _FAKE_BASE
            bases[b] = tag
            return DeviceArray(None, b, "float32", shape, c_strides(shape))

        gemm_out = None
        elemwise = []
        for i, node in enumerate(prog.nodes):
            args = [env[v] for v in node.inputs]
            if node.op == "Gemm":
                if gemm_out is not None:
                    return False
                z, a, x, y, b = args
                if not (node.inputs[0] == prog.inputs[0] and node.inputs[2] == prog.inputs[1]
                        and node.inputs[3] == prog.inputs[3]):
                    return False
                if not (np.ndim(a) == 0 and np.ndim(b) == 0 and float(np.asarray(a)) == 1.0
                        and float(np.asarray(b)) == 1.0):
                    return False
                gemm_out = env[node.outputs[0]] = fake((B, 4 * H), ("G",))
                g_base = gemm_out.ptr
                continue
            dev_in = [a for a in args if isinstance(a, DeviceArray)]
            if node.op == "Elemwise" and dev_in:
                if len(node.outputs) != 1:
                    return False
                out = fake((B, H), ("E", len(elemwise)))
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
            return False
        if gemm_out is None or len(elemwise) != 2:
            return False

        def role(a, cn_ptr):
            if not isinstance(a, DeviceArray):
                return None
            if a.shape != (B, H):
                return None
            if g_base <= a.ptr < g_base + 4 * H * 4 and a.strides == (4 * H, 1):
                off = (a.ptr - g_base) // 4
                if off % H:
                    return None
                return ("gi", "gf", "go", "gg")[off // H]
            if a.ptr == _FAKE_BASE * 3 and a.strides == (H, 1):
                return "c"
            if cn_ptr is not None and a.ptr == cn_ptr and a.strides == (H, 1):
                return "cn"
            return None

        (n1, a1, o1), (n2, a2, o2) = elemwise
        r1 = [role(a, None) for a in a1]
        r2 = [role(a, o1.ptr) for a in a2]
        if None in r1 or None in r2:
            return False
        if _canon(n1.params["expr"], r1) != _C_NEW or _canon(n2.params["expr"], r2) != _H_NEW:
            return False
        if env[prog.outputs[0]] is not o2 or env[prog.outputs[1]] is not o1:
            return False
        return True


def run_lstm(T, x, U, hbuf, cbuf, pos_h, pos_c):
    """Launch the persistent kernel.  x: [>=T, B, 4H]; U: [H, 4H]; hbuf/cbuf: rings."""
    lib = _lib.load()
    B, H = hbuf.shape[1], hbuf.shape[2]
    nbytes = C.c_size_t()
    _lib.check(lib.ab_lstm_scan_workspace_bytes(B, H, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=hbuf.owner.device)
    _lib.check(lib.ab_lstm_scan(T, B, H, x.ptr, x.strides[0], x.strides[1], U.ptr, U.strides[0],
                                U.strides[1], hbuf.ptr, hbuf.shape[0], pos_h, cbuf.ptr,
                                cbuf.shape[0], pos_c, ws.data_ptr(), nbytes.value, stream_handle()))


def eligible(x, U, hbuf, cbuf, T):
    lib = _lib.load()
    if any(a.dtype != np.float32 for a in (x, U, hbuf, cbuf)):
        return False
    if x.ndim != 3 or U.ndim != 2 or hbuf.ndim != 3 or cbuf.ndim != 3:
        return False
    B, H = hbuf.shape[1], hbuf.shape[2]
    if cbuf.shape[1:] != (B, H) or x.shape[1:] != (B, 4 * H) or U.shape != (H, 4 * H):
        return False
    if not (hbuf.is_c_contiguous() and cbuf.is_c_contiguous()):
        return False
    if x.strides[2] != 1 or x.strides[1] % 4 or x.strides[0] % 4 or x.ptr % 16:
        return False
    if hbuf.ptr % 16 or cbuf.ptr % 16:
        return False
    # tensor-core tiles only pay off above a minimum size; tiny problems keep the general loop
    if B * H < 128 * 64 or T < 2:
        return False
    return bool(lib.ab_lstm_scan_supported(T, B, H))
