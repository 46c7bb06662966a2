"""Host-side evaluation of *metadata* scalar expressions.

About half of a dynamically-shaped optimised graph is int64 shape arithmetic
(``Shape_i`` → ``ScalarFromTensor`` → scalar ``Composite`` → ``Assert`` →
``AllocEmpty``).  Those values decide allocations and launch
geometry, so they must be known on the host before any kernel is launched;
they are a handful of integers, not tensor work.  This evaluator covers the
operator subset such graphs use; anything else on ≤1-element operands is sent
to the device kernel instead (never the other way round for real tensors).
"""

from __future__ import annotations

import numpy as np

MAX_HOST_ELEMS = 64


def _np(x):
    return np.asarray(x)


def _div_int(x, y):
    y_safe = np.where(y == 0, 1, y)
    return np.where(y == 0, 0, np.floor_divide(x, y_safe))


def _mod_int(x, y):
    y_safe = np.where(y == 0, 1, y)
    return np.where(y == 0, 0, np.mod(x, y_safe))


def _true_div(x, y):
    if x.dtype.kind in "biu" and y.dtype.kind in "biu":
        return x.astype(np.float64) / y
    return x / y


_OPS = {
    "add": lambda a, out: _fold(a, np.logical_or if out == np.bool_ else np.add),
    "mul": lambda a, out: _fold(a, np.logical_and if out == np.bool_ else np.multiply),
    "sub": lambda a, out: a[0] - a[1],
    "neg": lambda a, out: -a[0],
    "abs": lambda a, out: np.abs(a[0]),
    "sqr": lambda a, out: a[0] * a[0],
    "identity": lambda a, out: a[0],
    "second": lambda a, out: np.broadcast_to(a[1], np.broadcast_shapes(a[0].shape, a[1].shape)),
    "true_divide": lambda a, out: _true_div(a[0], a[1]),
    "int_div": lambda a, out: _div_int(a[0], a[1]) if out.kind in "biu" else np.floor(a[0] / a[1]),
    "mod": lambda a, out: _mod_int(a[0], a[1]) if out.kind in "biu" else np.mod(a[0], a[1]),
    "lt": lambda a, out: a[0] < a[1],
    "gt": lambda a, out: a[0] > a[1],
    "le": lambda a, out: a[0] <= a[1],
    "ge": lambda a, out: a[0] >= a[1],
    "eq": lambda a, out: a[0] == a[1],
    "neq": lambda a, out: a[0] != a[1],
    "switch": lambda a, out: np.where(a[0] != 0, a[1], a[2]),
    "maximum": lambda a, out: np.maximum(a[0], a[1]),
    "minimum": lambda a, out: np.minimum(a[0], a[1]),
    "and": lambda a, out: np.bitwise_and(a[0], a[1]),
    "or": lambda a, out: np.bitwise_or(a[0], a[1]),
    "xor": lambda a, out: np.bitwise_xor(a[0], a[1]),
    "invert": lambda a, out: np.logical_not(a[0]) if out == np.bool_ else np.invert(a[0]),
    "clip": lambda a, out: np.where(a[0] < a[1], a[1], np.where(a[0] > a[2], a[2], a[0])),
    "sgn": lambda a, out: np.sign(a[0]),
    "cast": lambda a, out: (a[0] != 0) if out == np.bool_ else a[0].astype(out),
    "ceil": lambda a, out: np.ceil(a[0].astype(out)),
    "floor": lambda a, out: np.floor(a[0].astype(out)),
    "reciprocal": lambda a, out: 1.0 / a[0].astype(out),
    "sqrt": lambda a, out: np.sqrt(a[0].astype(out)),
}


def _fold(args, fn):
    r = args[0]
    for x in args[1:]:
        r = fn(r, x)
    return r


def supports(expr):
    for st in expr["stmts"]:
        if st["op"] not in _OPS:
            return False
        # float IntDiv / Mod follow branchy C formulas in the reference (basic.py:2083-2121,
        # :2207-2236) that only the device bodies reproduce: never evaluate them here
        if st["op"] in ("int_div", "mod") and np.dtype(st["dtype"]).kind not in "biu":
            return False
    return True


def eval_expr(expr, inputs):
    """Evaluate an IR scalar expression on small host values."""
    temps = []
    inputs = [_np(x) for x in inputs]

    def ref(r):
        if isinstance(r, dict):
            return np.asarray(r["const"], dtype=r["dtype"])
        if r[0] == "i":
            return inputs[int(r[1:])]
        return temps[int(r[1:])]

    with np.errstate(all="ignore"):
        for st in expr["stmts"]:
            out = np.dtype(st["dtype"])
            val = _OPS[st["op"]]([ref(r) for r in st["args"]], out)
            temps.append(np.asarray(val).astype(out, copy=False))
    return [ref(r) for r in expr["outputs"]]
