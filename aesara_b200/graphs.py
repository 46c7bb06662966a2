"""Symbolic graphs of the BASELINE.json configs, built with the reference front-end.

Used by the golden-fixture generator (``tests/golden/make_golden.py``), by the
linker-level tests (when the front-end is importable) and by nothing on the
device path.  Definitions follow SURVEY.md §8(d).
"""

from __future__ import annotations

import numpy as np


def _at():
    from .compat.bootstrap import load_aesara

    aesara = load_aesara()
    import aesara.tensor as at

    return aesara, at


def cfg1_readme():
    """README example ``d = a/a + (M+a).dot(v)`` (f64)."""
    aesara, at = _at()
    a, v, M = at.dscalar("a"), at.dvector("v"), at.dmatrix("M")
    d = a / a + (M + a).dot(v)
    return [a, v, M], [d]


def cfg1_inputs(n=1000, seed=0):
    rng = np.random.default_rng(seed)
    return [np.float64(1.5), rng.standard_normal(n), rng.standard_normal((n, n))]


def cfg2_fused_elemwise():
    """``softplus(tanh(x) + y) * z`` on three f32 vectors."""
    aesara, at = _at()
    x, y, z = at.fvector("x"), at.fvector("y"), at.fvector("z")
    return [x, y, z], [at.softplus(at.tanh(x) + y) * z]


def cfg2_inputs(n, seed=0, yscale=1.0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n).astype("float32")
    y = (rng.standard_normal(n) * yscale).astype("float32")
    z = rng.standard_normal(n).astype("float32")
    return [x, y, z]


def cfg3_mlp(shared_weights=False, H=None):
    """2-layer tanh MLP, MSE loss, forward + ``aesara.grad`` wrt W1,b1,W2,b2.

    With ``shared_weights`` the parameters are ``aesara.shared`` (as in the
    survey); otherwise they are explicit inputs (device-resident benchmarking
    does not need the shared-variable machinery)."""
    aesara, at = _at()
    X, Y = at.fmatrix("X"), at.fmatrix("Y")
    if shared_weights:
        assert H is not None
        z = np.zeros
        W1 = aesara.shared(z((H, H), "float32"), name="W1")
        b1 = aesara.shared(z((H,), "float32"), name="b1")
        W2 = aesara.shared(z((H, H), "float32"), name="W2")
        b2 = aesara.shared(z((H,), "float32"), name="b2")
        ins = [X, Y]
    else:
        W1, b1 = at.fmatrix("W1"), at.fvector("b1")
        W2, b2 = at.fmatrix("W2"), at.fvector("b2")
        ins = [X, Y, W1, b1, W2, b2]
    h = at.tanh(X @ W1 + b1)
    out = h @ W2 + b2
    loss = at.mean((out - Y) ** 2)
    g = aesara.grad(loss, [W1, b1, W2, b2])
    return ins, [loss] + list(g)


def cfg3_inputs(B, H, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((B, H)).astype("float32")
    Y = rng.standard_normal((B, H)).astype("float32")
    W1 = (rng.standard_normal((H, H)) / np.sqrt(H)).astype("float32")
    W2 = (rng.standard_normal((H, H)) / np.sqrt(H)).astype("float32")
    b1 = np.zeros(H, "float32")
    b2 = np.zeros(H, "float32")
    return [X, Y, W1, b1, W2, b2]


def cfg4_lstm_scan():
    """LSTM cell over T steps with pre-projected inputs; returns h_T, c_T."""
    aesara, at = _at()
    x = at.ftensor3("x")  # [T, B, 4H]
    h0, c0 = at.fmatrix("h0"), at.fmatrix("c0")
    U = at.fmatrix("U")  # [H, 4H]

    def step(x_t, h_tm1, c_tm1, U_):
        H = h_tm1.shape[1]
        pre = x_t + at.dot(h_tm1, U_)
        i = at.sigmoid(pre[:, :H])
        f = at.sigmoid(pre[:, H : 2 * H])
        o = at.sigmoid(pre[:, 2 * H : 3 * H])
        g = at.tanh(pre[:, 3 * H :])
        c = f * c_tm1 + i * g
        h = o * at.tanh(c)
        return h, c

    (hs, cs), _ = aesara.scan(step, sequences=[x], outputs_info=[h0, c0], non_sequences=[U])
    return [x, h0, c0, U], [hs[-1], cs[-1]]


def cfg4_inputs(T, B, H, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((T, B, 4 * H)).astype("float32")
    U = (rng.standard_normal((H, 4 * H)) / np.sqrt(H)).astype("float32")
    h0 = np.zeros((B, H), "float32")
    c0 = np.zeros((B, H), "float32")
    return [x, h0, c0, U]


def cfg5_logreg():
    """Logistic-regression cost and gradient wrt (w, b)."""
    aesara, at = _at()
    X, y, w, b = at.fmatrix("X"), at.fvector("y"), at.fvector("w"), at.fscalar("b")
    p = at.sigmoid(X @ w + b)
    cost = at.mean(-y * at.log(p) - (1 - y) * at.log(1 - p))
    gw, gb = aesara.grad(cost, [w, b])
    return [X, y, w, b], [cost, gw, gb]


def cfg5_inputs(N, D, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D)).astype("float32")
    y = (rng.random(N) < 0.5).astype("float32")
    w = (rng.standard_normal(D) * 0.01).astype("float32")
    return [X, y, w, np.float32(0.0)]


def optimized_program(inputs, outputs, name=None, optimizer="fast_run", linker="cvm"):
    """Run the reference rewriter and lower the resulting fgraph.  Returns
    ``(program, reference_function)``; the reference function is compiled with
    the reference C-linker (``Mode("cvm")``) and is the parity oracle."""
    aesara, _ = _at()
    from aesara.compile.mode import Mode

    from .lower import lower_fgraph

    f = aesara.function(inputs, outputs, mode=Mode(linker, optimizer), on_unused_input="ignore")
    prog = lower_fgraph(f.maker.fgraph, name=name)
    return prog, f
