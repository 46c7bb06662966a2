"""ORACLE (test infrastructure only — never imported by the product path).

A NumPy interpreter for lowered programs (``aesara_b200.ir.Program``): the CPU
restatement of what the reference's per-Op ``perform``/C implementations
compute for every node kind on the hot path.  Each handler
cites the reference code it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
leg may import this module.
"""

import numpy as np

from .scalar_np import eval_expr

_H = {}


def _h(name):
    def deco(fn):
        _H[name] = fn
        return fn

    return deco


class OracleAssertionError(AssertionError):
    pass


def run_program(prog, inputs, return_env=False, trace=None):
    """Evaluate ``prog`` on NumPy ``inputs`` (list in ``prog.inputs`` order)."""
    env = {}
    for vid, v in enumerate(prog.vars):
        if v.const is not None:
            env[vid] = v.const
        elif v.const_other is not None:
            env[vid] = _other_const(v.const_other)
    if len(inputs) != len(prog.inputs):
        raise TypeError(f"expected {len(prog.inputs)} inputs, got {len(inputs)}")
    for vid, val in zip(prog.inputs, inputs):
        var = prog.vars[vid]
        if var.kind == "tensor":
            val = np.asarray(val, dtype=var.dtype)
        elif var.kind == "scalar":
            val = np.dtype(var.dtype).type(val)
        env[vid] = val
    for pos, node in enumerate(prog.nodes):
        args = [env[i] for i in node.inputs]
        outs = _H[node.op](node, args, prog)
        if not isinstance(outs, (list, tuple)):
            outs = [outs]
        for vid, o in zip(node.outputs, outs):
            var = prog.vars[vid]
            if var.kind == "tensor" and o is not None:
                o = np.asarray(o)
                if o.dtype != np.dtype(var.dtype):
                    o = o.astype(var.dtype)
            env[vid] = o
        if trace is not None:
            trace[pos] = [None if env[v] is None or isinstance(env[v], slice) else np.array(env[v], copy=True)
                          for v in node.outputs]
    outs = [env[v] for v in prog.outputs]
    if return_env:
        return outs, env
    return outs


def _other_const(c):
    if "none" in c:
        return None
    if "slice" in c:
        return slice(*c["slice"])
    raise NotImplementedError(c)


# -- elementwise -------------------------------------------------------------------
@_h("Elemwise")
def _elemwise(node, args, prog):
    # aesara/tensor/elemwise.py:725-820 (perform) with the broadcasting rule of
    # elemwise_cgen.py:72-125 (size-1 dims broadcast at run time, else must match)
    args = [np.asarray(a) for a in args]
    nd = args[0].ndim if args else 0
    for d in range(nd):
        sizes = {a.shape[d] for a in args if a.shape[d] != 1}
        if len(sizes) > 1:
            raise ValueError(
                f"Input dimension mismatch. (dim {d}: sizes {sorted(sizes)})"
            )
    outs = eval_expr(node.params["expr"], args)
    shape = np.broadcast_shapes(*[a.shape for a in args]) if args else ()
    res = []
    for k, o in enumerate(outs):
        o = np.broadcast_to(o, shape)
        ip = node.params.get("inplace", {}).get(str(k))
        if ip is not None and args[ip].shape == tuple(shape) and args[ip].flags.writeable:
            args[ip][...] = o
            res.append(args[ip])
        else:
            res.append(np.array(o))
    return res


@_h("ScalarOp")
def _scalarop(node, args, prog):
    outs = eval_expr(node.params["expr"], [np.asarray(a) for a in args])
    return [o.dtype.type(o.item()) if hasattr(o, "item") else o for o in outs]


@_h("DimShuffle")
def _dimshuffle(node, args, prog):
    # aesara/tensor/elemwise.py:222-239: transpose(shuffle + drop).reshape(with 1s)
    (x,) = args
    x = np.asarray(x)
    order = node.params["new_order"]
    kept = [o for o in order if o != "x"]
    drop = [i for i in range(x.ndim) if i not in kept]
    for d in drop:
        if x.shape[d] != 1:
            raise ValueError("DimShuffle: cannot drop a non-broadcastable dimension")
    y = x.transpose(kept + drop)
    shape = [1 if o == "x" else x.shape[o] for o in order]
    return y.reshape(shape)


_UFUNC = {
    "add": np.add, "mul": np.multiply, "maximum": np.maximum, "minimum": np.minimum,
    "and": np.bitwise_and, "or": np.bitwise_or, "xor": np.bitwise_xor,
}


@_h("CAReduce")
def _careduce(node, args, prog):
    # aesara/tensor/elemwise.py:1495-1513: cast to acc dtype, ufunc.reduce, cast out
    (x,) = args
    p = node.params
    x = np.asarray(x)
    acc = np.dtype(p["acc_dtype"])
    axis = tuple(p["axis"])
    xa = x.astype(acc)
    if p["scalar_op"] == "mul_without_zeros":  # aesara/tensor/math.py:2713-2754
        nz = xa != 0
        out = np.where(nz.any(axis=axis), np.prod(np.where(nz, xa, 1), axis=axis, dtype=acc), 0)
        return np.asarray(out).astype(p["out_dtype"])
    if p["scalar_op"] in ("maximum", "minimum") and any(x.shape[a] == 0 for a in axis):
        raise ValueError("zero-size array to reduction operation with no identity")
    uf = _UFUNC[p["scalar_op"]]
    if axis == ():
        out = xa
    else:
        out = uf.reduce(xa, axis=axis, dtype=acc)
    return np.asarray(out).astype(p["out_dtype"])


# -- row ops ---------------------------------------------------------------------------
@_h("Softmax")
def _softmax(node, args, prog):
    p = node.params
    axis = p["axis"]
    if p["mode"] == 2:  # aesara/tensor/special.py:38-43 SoftmaxGrad.perform
        dy, sm = args
        t = dy * sm
        return t - np.sum(t, axis=axis, keepdims=True) * sm
    (x,) = args
    x = np.asarray(x)
    xdev = x - np.max(x, axis=axis, keepdims=True)
    if p["mode"] == 0:  # special.py:440-478
        e = np.exp(xdev)
        return (e / np.sum(e, axis=axis, keepdims=True)).astype(x.dtype)
    # special.py:715-740
    return (xdev - np.log(np.sum(np.exp(xdev), axis=axis, keepdims=True))).astype(x.dtype)


@_h("MaxAndArgmax")
def _maxandargmax(node, args, prog):  # aesara/tensor/math.py:164-186
    (x,) = args
    x = np.asarray(x)
    axes = tuple(node.params["axes"])
    keep = [d for d in range(x.ndim) if d not in axes]
    xt = np.transpose(x, keep + list(axes))
    kept_shape = xt.shape[: len(keep)]
    flat = xt.reshape(kept_shape + (-1,))
    idx = np.argmax(flat, axis=-1).astype("int64")
    if node.params.get("argmax_only"):
        return idx
    return [np.max(x, axis=axes).astype(x.dtype), idx]


# -- BLAS family -------------------------------------------------------------------
@_h("Dot22")
def _dot22(node, args, prog):  # aesara/tensor/blas.py:1685-1694
    x, y = args
    return np.dot(x, y)


@_h("Dot22Scalar")
def _dot22scalar(node, args, prog):  # blas.py:1995-2005
    x, y, a = args
    return (np.asarray(a, dtype=x.dtype) * np.dot(x, y)).astype(x.dtype)


@_h("Dot")
def _dot(node, args, prog):  # aesara/tensor/math.py:1879
    return np.dot(args[0], args[1])


@_h("Gemm")
def _gemm(node, args, prog):  # aesara/tensor/blas.py:984-1017
    z, a, x, y, b = args
    a = np.asarray(a, dtype=z.dtype)
    b = np.asarray(b, dtype=z.dtype)
    if x.shape[1] != y.shape[0]:
        raise ValueError("Shape mismatch: x has %d cols but y has %d rows" % (x.shape[1], y.shape[0]))
    if not node.params["inplace"]:
        z = z.copy()
    if (x.shape[0] > z.shape[0]) or (y.shape[1] > z.shape[1]):
        z = np.broadcast_to(z, (max(x.shape[0], z.shape[0]), max(y.shape[1], z.shape[1]))).copy()
    if z.shape != (x.shape[0], y.shape[1]):
        raise ValueError("Shape mismatch: z vs x.y")
    if b == 0.0:
        z[:] = a * np.dot(x, y)
    else:
        z *= b
        z += a * np.dot(x, y)
    return z


@_h("Gemv")
def _gemv(node, args, prog):  # aesara/tensor/blas.py:279-318
    y, alpha, A, x, beta = args
    if A.shape[0] != y.shape[0] or A.shape[1] != x.shape[0]:
        raise ValueError(
            "Incompatible shapes for gemv "
            f"(beta * y + alpha * dot(A, x)). y: {y.shape}, A: {A.shape}, x: {x.shape}"
        )
    out = np.dot(A, x).astype(y.dtype)
    alpha = np.asarray(alpha, dtype=y.dtype)
    beta = np.asarray(beta, dtype=y.dtype)
    if alpha != 1:
        out *= alpha
    if beta != 0:  # when beta == 0 y may be uninitialised and must not be read
        out += beta * y
    if node.params["inplace"]:
        y[...] = out
        return y
    return out


@_h("Ger")
def _ger(node, args, prog):  # aesara/tensor/blas.py:381-392
    A, alpha, x, y = args
    if not node.params["inplace"]:
        A = A.copy()
    A += np.asarray(alpha, dtype=A.dtype) * np.outer(x, y)
    return A


# -- allocation / copies ---------------------------------------------------------------
@_h("AllocEmpty")
def _allocempty(node, args, prog):  # aesara/tensor/basic.py:3833
    # contents are unspecified in the reference; the oracle fills with NaN/-1 so any
    # consumer that reads them shows up in a parity test
    shape = tuple(int(s) for s in args)
    dt = np.dtype(node.params["dtype"])
    out = np.empty(shape, dtype=dt)
    out.fill(np.nan if dt.kind == "f" else 0)
    return out


@_h("Alloc")
def _alloc(node, args, prog):  # aesara/tensor/basic.py:1389 (perform :1468)
    v, *shape = args
    shape = tuple(int(s) for s in shape)
    return np.array(np.broadcast_to(v, shape))


@_h("AdvancedSubtensor")
def _advsub(node, args, prog):  # aesara/tensor/subtensor.py:2577 (perform :2641)
    x, *idx = args
    return np.asarray(x)[tuple(np.asarray(i) for i in idx)]


@_h("AdvancedIncSubtensor")
def _advincsub(node, args, prog):  # aesara/tensor/subtensor.py:2727 (perform :2768)
    x, y, *idx = args
    out = np.array(x, copy=True)
    idx = tuple(np.asarray(i) for i in idx)
    if node.params["set"]:
        out[idx] = y
    elif node.params.get("ignore_duplicates"):
        out[idx] += y
    else:
        np.add.at(out, idx, y)
    return out


@_h("BatchedDot")
def _batched_dot(node, args, prog):  # aesara/tensor/blas.py:2232 (perform :2277)
    x, y = (np.asarray(a) for a in args)
    if x.shape[0] != y.shape[0]:
        raise TypeError("BatchedDot: inputs must have the same size in axis 0")
    return np.stack([np.dot(x[b], y[b]) for b in range(x.shape[0])]) if x.shape[0] else \
        np.zeros((0,) + np.dot(x[:0].sum(0), y[:0].sum(0)).shape, dtype=np.result_type(x, y))


@_h("IfElse")
def _ifelse(node, args, prog):  # aesara/ifelse.py:44 (thunk :232-300)
    n = node.params["n_outs"]
    cond = bool(np.asarray(args[0]).item())
    vals = args[1 : 1 + n] if cond else args[1 + n : 1 + 2 * n]
    return [np.array(v, copy=True) for v in vals]


@_h("CumOp")
def _cumop(node, args, prog):  # aesara/tensor/extra_ops.py:253 (C code :325-375: accumulates in x's dtype)
    x = np.asarray(args[0])
    f = np.cumsum if node.params["mode"] == "add" else np.cumprod
    return f(x, axis=node.params["axis"], dtype=x.dtype)


@_h("ExtractDiag")
def _extract_diag(node, args, prog):  # aesara/tensor/basic.py:3480 (perform :3557)
    p = node.params
    return np.array(np.asarray(args[0]).diagonal(p["offset"], p["axis1"], p["axis2"]), copy=True)


@_h("AllocDiag")
def _alloc_diag(node, args, prog):  # aesara/tensor/basic.py:3600 (perform :3640), vector input
    return np.diag(np.asarray(args[0]), node.params["offset"])


@_h("Tri")
def _tri(node, args, prog):  # aesara/tensor/basic.py:1178 (perform :1196)
    n, m, k = (int(np.asarray(a).item()) for a in args)
    return np.tri(n, m, k, dtype=node.params["dtype"])


@_h("Eye")
def _eye(node, args, prog):  # aesara/tensor/basic.py:1318 (perform :1339)
    n, m, k = (int(np.asarray(a).item()) for a in args)
    return np.eye(n, m, k, dtype=node.params["dtype"])


@_h("ARange")
def _arange(node, args, prog):  # aesara/tensor/basic.py:2867 (perform :2937)
    start, stop, step = (np.asarray(a).item() for a in args)
    return np.arange(start, stop, step, dtype=node.params["dtype"])


@_h("BroadcastTo")
def _broadcast_to(node, args, prog):  # aesara/tensor/extra_ops.py:1613 (perform :1652)
    v, *shape = args
    return np.broadcast_to(v, tuple(int(s) for s in shape))


@_h("DeepCopy")
def _deepcopy(node, args, prog):  # aesara/compile/ops.py:149
    return np.array(args[0], copy=True)


@_h("View")
def _view(node, args, prog):  # aesara/compile/ops.py:37, tensor/shape.py:939
    return args[0]


@_h("Reshape")
def _reshape(node, args, prog):  # aesara/tensor/shape.py:589
    x, shp = args
    return np.reshape(x, tuple(int(s) for s in np.asarray(shp)))


# -- host metadata ----------------------------------------------------------------------
@_h("Shape_i")
def _shape_i(node, args, prog):  # aesara/tensor/shape.py:189
    return np.asarray(np.shape(args[0])[node.params["i"]], dtype="int64")


@_h("Shape")
def _shape(node, args, prog):  # aesara/tensor/shape.py:47
    return np.asarray(np.shape(args[0]), dtype="int64")


@_h("ScalarFromTensor")
def _sft(node, args, prog):  # aesara/tensor/basic.py:594
    a = np.asarray(args[0])
    return a.dtype.type(a.item())


@_h("TensorFromScalar")
def _tfs(node, args, prog):  # aesara/tensor/basic.py:539
    return np.asarray(args[0])


@_h("MakeVector")
def _makevector(node, args, prog):  # aesara/tensor/basic.py:1629
    return np.asarray(args, dtype=node.params["dtype"]).reshape(len(args))


@_h("Assert")
def _assert(node, args, prog):  # aesara/raise_op.py:28-120
    val, *conds = args
    if not all(bool(np.all(c)) for c in conds):
        raise OracleAssertionError(node.params["msg"])
    return val


# -- indexing -----------------------------------------------------------------------------
def build_index(idx_list, runtime):
    """Rebuild the Python index tuple from the template (``get_idx_list``,
    aesara/tensor/subtensor.py:185) consuming ``runtime`` scalars in order."""
    it = iter(runtime)

    def elem(e):
        if e is None:
            return None
        if e == "in":
            return int(next(it))
        return int(e)

    out = []
    for entry in idx_list:
        if "slice" in entry:
            out.append(slice(*[elem(e) for e in entry["slice"]]))
        else:
            out.append(elem(entry["index"]))
    return tuple(out)


@_h("Subtensor")
def _subtensor(node, args, prog):  # aesara/tensor/subtensor.py:682 (perform :760)
    x, *rt = args
    return np.asarray(x)[build_index(node.params["idx_list"], rt)]


@_h("IncSubtensor")
def _incsubtensor(node, args, prog):  # aesara/tensor/subtensor.py:1454 (perform :1560)
    x, y, *rt = args
    p = node.params
    if not p["inplace"]:
        x = x.copy()
    idx = build_index(p["idx_list"], rt)
    if p["set"]:
        x[idx] = y
    else:
        x[idx] += y
    return x


@_h("AdvancedSubtensor1")
def _advsub1(node, args, prog):  # aesara/tensor/subtensor.py:1953-1990
    x, idx = args
    return np.take(np.asarray(x), np.asarray(idx), axis=0)


@_h("AdvancedIncSubtensor1")
def _advincsub1(node, args, prog):  # aesara/tensor/subtensor.py:2128 (perform)
    x, y, idx = args
    p = node.params
    if not p["inplace"]:
        x = x.copy()
    if p["set"]:
        x[idx] = y
    else:
        np.add.at(x, idx, y)
    return x


@_h("Join")
def _join(node, args, prog):  # aesara/tensor/basic.py:2142 (perform: np.concatenate)
    axis, *tensors = args
    return np.concatenate(tensors, axis=int(axis))


@_h("Split")
def _split(node, args, prog):  # aesara/tensor/basic.py:1882
    x, axis, splits = args
    axis = int(axis)
    splits = [int(s) for s in np.asarray(splits)]
    if sum(splits) != np.shape(x)[axis]:
        raise ValueError("Split: the split sizes do not sum to the input length along the axis")
    outs, start = [], 0
    for s in splits:
        sl = [slice(None)] * np.ndim(x)
        sl[axis] = slice(start, start + s)
        outs.append(np.array(np.asarray(x)[tuple(sl)]))
        start += s
    return outs


# -- Scan ---------------------------------------------------------------------------------
@_h("Scan")
def _scan(node, args, prog):
    """aesara/scan/op.py:1673-2160 (``Scan.perform``), restated without the
    storage-reuse bookkeeping: circular output buffers of length ``store_steps``,
    taps gathered at ``(pos + tap) % store_steps``, final rotation (:2105-2134)
    and zero-fill of never-written rows (:2139-2159)."""
    info = node.params["info"]
    inner = node.params["inner"]
    destroy = {int(k) for k in node.params.get("destroy_map", {})}
    n_seqs = info["n_seqs"]
    mm_in, mm_out = info["mit_mot_in_slices"], info["mit_mot_out_slices"]
    ms_in, ss_in = info["mit_sot_in_slices"], info["sit_sot_in_slices"]
    n_mit_mot, n_mit_sot, n_sit_sot = len(mm_in), len(ms_in), len(ss_in)
    n_nit_sot, n_shared = info["n_nit_sot"], info["n_shared_outs"]
    n_outs = n_mit_mot + n_mit_sot + n_sit_sot
    tap_array = mm_in + ms_in + ss_in
    mintaps = [min(t) for t in tap_array] + [0] * n_nit_sot

    n_steps = int(args[0])
    if n_steps < 0:
        raise IndexError(f"Scan was asked to run for negative number of step {n_steps}")
    seqs = args[1 : 1 + n_seqs]
    for idx, s in enumerate(seqs):
        if s.shape[0] < n_steps:
            raise ValueError(
                f"Sequence {idx} has shape {s.shape} but the Scan's required number of steps is {n_steps}"
            )
    o0 = 1 + n_seqs
    states = args[o0 : o0 + n_outs]
    shared0 = args[o0 + n_outs : o0 + n_outs + n_shared]
    nit_len = [int(a) for a in args[o0 + n_outs + n_shared : o0 + n_outs + n_shared + n_nit_sot]]
    non_seqs = args[o0 + n_outs + n_shared + n_nit_sot :]

    store_steps = [s.shape[0] for s in states] + nit_len
    bufs = []
    for idx, s in enumerate(states):
        bufs.append(s if idx in destroy else s.copy())
    nit_bufs = [None] * n_nit_sot
    shared_vals = list(shared0)
    if n_steps == 0:
        outs = bufs + [
            np.empty((0,) * prog.vars[node.outputs[n_outs + j]].ndim,
                     dtype=prog.vars[node.outputs[n_outs + j]].dtype)
            for j in range(n_nit_sot)
        ] + shared_vals
        return outs

    pos = [(-mintaps[idx]) % store_steps[idx] for idx in range(n_outs + n_nit_sot)]
    i, cond = 0, True
    while i < n_steps and cond:
        inner_in = [s[i] for s in seqs]
        for idx, taps in enumerate(tap_array):
            for t in taps:
                inner_in.append(bufs[idx][(pos[idx] + t) % store_steps[idx]])
        inner_in += shared_vals
        inner_in += list(non_seqs)
        inner_out = run_program(inner, [np.array(a) if isinstance(a, np.ndarray) else a for a in inner_in])
        k = 0
        for g, taps in enumerate(mm_in):
            for out_slice in mm_out[g]:
                bufs[g][out_slice + pos[g]] = inner_out[k]
                k += 1
        for j in range(n_mit_mot, n_outs):
            bufs[j][pos[j]] = inner_out[k]
            k += 1
        for j in range(n_nit_sot):
            val = np.asarray(inner_out[k])
            if i == 0:
                nit_bufs[j] = np.empty((store_steps[n_outs + j],) + val.shape,
                                       dtype=prog.vars[node.outputs[n_outs + j]].dtype)
            nit_bufs[j][pos[n_outs + j]] = val
            k += 1
        for j in range(n_shared):
            shared_vals[j] = inner_out[k]
            k += 1
        if info["as_while"]:
            cond = not bool(np.asarray(inner_out[k]).item())
        pos = [(p + 1) % s for p, s in zip(pos, store_steps)]
        i += 1

    allb = bufs + nit_bufs
    for idx in range(n_mit_mot, n_outs + n_nit_sot):
        st = store_steps[idx]
        if st < i - mintaps[idx] and pos[idx] < st:
            allb[idx][...] = np.roll(allb[idx], -pos[idx], axis=0)
        elif st > i - mintaps[idx]:
            allb[idx][i - mintaps[idx] :] = 0
            if i < n_steps:
                allb[idx] = allb[idx][: -(n_steps - i)]
    return allb + shared_vals
