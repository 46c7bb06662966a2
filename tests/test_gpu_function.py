"""GPU: the drop-in boundary on hardware.

``aesara.function(..., mode="B200")`` -> ``FunctionMaker`` -> ``B200Linker.accept /
make_all`` -> ``B200VM`` -> ``ProgramExecutor`` -> C ABI -> CUDA, compared IN THE SAME
PROCESS with the reference's own C-linker (``mode="FAST_RUN"`` = ``Mode("cvm","fast_run")``,
``aesara/compile/mode.py:446-452``; contract ``compile/function/types.py:791-1048``,
``link/vm.py:1212-1324``).  The reference front-end is the travelling copy under
``oracle/_ref`` (``oracle/ref.py``): a checker, never on the product path.
"""
import numpy as np
import pytest

from aesara_b200.compat import bootstrap

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not bootstrap.available(), reason="reference front-end (oracle/_ref) not present")]


@pytest.fixture(scope="module")
def env():
    import torch

    from aesara_b200.runtime import lib

    lib.check(lib.load().ab_init(0))
    torch.cuda.set_device(0)
    aesara = bootstrap.load_aesara()
    import aesara_b200.linker as L

    return aesara, L, lib.load()


def _close(got, want, blas, what, rtol=1e-5):
    from tests._cases import assert_matches

    assert_matches(np.asarray(got), np.asarray(want), blas=blas, rtol=rtol, what=what)


def _cfg(name):
    from aesara_b200 import graphs as G

    return {
        "cfg1": (G.cfg1_readme, lambda: G.cfg1_inputs(300)),
        "cfg2": (G.cfg2_fused_elemwise, lambda: G.cfg2_inputs(100003, yscale=12.0)),
        "cfg3": (G.cfg3_mlp, lambda: G.cfg3_inputs(384, 256)),
        "cfg4": (G.cfg4_lstm_scan, lambda: G.cfg4_inputs(6, 256, 64)),
        "cfg5": (G.cfg5_logreg, lambda: G.cfg5_inputs(5000, 128)),
    }[name]


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_function_call_matches_the_c_linker(env, name):
    """The five BASELINE graphs through ``Function.__call__`` on the device vs the
    reference C-linker on the host, same seeded inputs, same process."""
    aesara, L, lib = env
    build, inputs = _cfg(name)
    i, o = build()
    f = aesara.function(i, o, mode=L.mode(), on_unused_input="ignore")
    i2, o2 = build()
    g = aesara.function(i2, o2, mode="FAST_RUN", on_unused_input="ignore")
    assert isinstance(f.vm, L.B200VM) and not isinstance(g.vm, L.B200VM)
    vals = inputs()
    before = lib.ab_launch_count()
    got = f(*vals)
    launched = lib.ab_launch_count() - before
    assert launched > 0, "no kernel of libaesara_b200.so ran"
    want = g(*vals)
    blas = name != "cfg2"
    for k, (a, b) in enumerate(zip(got, want)):
        assert isinstance(a, np.ndarray)
        _close(a, b, blas, f"{name} output {k}: B200 Function vs C-linker Function")
    if name == "cfg4":
        scan = [st["runner"] for st in f.vm.executor._state if "runner" in st][0]
        assert scan.used_fast_path, "mode='B200' did not reach the persistent Scan kernel"
    # second call: storage cells are reused, results identical
    again = f(*vals)
    for a, b in zip(again, got):
        np.testing.assert_array_equal(a, b)


def test_mode_name_string(env):
    aesara, L, lib = env
    import aesara.tensor as at

    x = at.fvector("x")
    f = aesara.function([x], at.exp(x) + 1, mode="B200")
    assert isinstance(f.vm, L.B200VM)
    xv = np.linspace(-3, 3, 1000).astype("float32")
    _close(f(xv), np.exp(xv) + 1, False, "mode='B200'")


def test_updates_on_device_shared_variables_training_loop(env):
    """SGD on the cfg3 graph with ``updates=``: parameters in ``aesara_b200.shared`` stay on
    the device between calls (no download of the update, no upload at the next call) and
    follow the C-linker's trajectory with ordinary ``aesara.shared``."""
    aesara, L, lib = env
    import aesara.tensor as at
    from aesara_b200.runtime.device import DeviceArray
    from aesara_b200.sharedvar import shared

    rng = np.random.default_rng(0)
    B, H = 256, 128
    W0 = (rng.standard_normal((H, H)) / np.sqrt(H)).astype("float32")
    b0 = (rng.standard_normal(H) * 0.1).astype("float32")

    def build(mk):
        X, Y = at.fmatrix("X"), at.fmatrix("Y")
        W, b = mk(W0.copy(), "W"), mk(b0.copy(), "b")
        loss = at.mean((at.tanh(X @ W + b) - Y) ** 2)
        gW, gb = aesara.grad(loss, [W, b])
        return [X, Y], loss, [(W, W - 0.5 * gW), (b, b - 0.5 * gb)], (W, b)

    ins, loss, upd, (Wd, bd) = build(lambda v, n: shared(v, name=n))
    f = aesara.function(ins, loss, updates=upd, mode=L.mode())
    ins2, loss2, upd2, (Wh, bh) = build(lambda v, n: aesara.shared(v, name=n))
    g = aesara.function(ins2, loss2, updates=upd2, mode="FAST_RUN")
    Xv = rng.standard_normal((B, H)).astype("float32")
    Yv = np.tanh(rng.standard_normal((B, H))).astype("float32")
    for step in range(4):
        ld, lh = f(Xv, Yv), g(Xv, Yv)
        _close(ld, lh, True, f"loss at step {step}", rtol=2e-5)
        assert Wd.is_on_device() and bd.is_on_device()
        assert isinstance(Wd.get_value(borrow=True, return_internal_type=True), DeviceArray)
    _close(Wd.get_value(), Wh.get_value(), True, "W after 4 updates", rtol=5e-5)
    _close(bd.get_value(), bh.get_value(), True, "b after 4 updates", rtol=5e-5)
    # set_value with a device array, then one more step
    Wd.set_value(DeviceArray.from_numpy(W0))
    Wh.set_value(W0)
    bd.set_value(b0)
    bh.set_value(b0)
    _close(f(Xv, Yv), g(Xv, Yv), True, "loss after set_value", rtol=2e-5)


def test_ordinary_shared_variable_updates(env):
    aesara, L, lib = env
    import aesara.tensor as at

    x = at.fvector("x")
    acc = aesara.shared(np.zeros(2000, "float32"), name="acc")
    f = aesara.function([x], (acc * 2).sum(), updates=[(acc, acc + x)], mode=L.mode())
    xv = np.arange(2000, dtype="float32") / 100
    assert float(f(xv)) == 0.0
    r = f(xv)
    np.testing.assert_allclose(r, 2 * xv.sum(), rtol=1e-6)
    assert isinstance(acc.container.storage[0], np.ndarray)
    np.testing.assert_allclose(acc.get_value(), 2 * xv, rtol=1e-6)


def test_trust_input_device_arrays_and_device_outputs(env):
    """SURVEY 8d timing protocol: ``f.trust_input = True``, device-resident inputs, outputs
    left on the device."""
    aesara, L, lib = env
    from aesara_b200.runtime.device import DeviceArray

    build, inputs = _cfg("cfg3")
    i, o = build()
    f = aesara.function(i, o, mode=L.mode(device_outputs=True))
    f.trust_input = True
    vals = inputs()
    dev = [DeviceArray.from_numpy(v) for v in vals]
    got = f(*dev)
    assert all(isinstance(a, DeviceArray) for a in got)
    i2, o2 = build()
    want = aesara.function(i2, o2, mode="FAST_RUN")(*vals)
    for k, (a, b) in enumerate(zip(got, want)):
        _close(a.to_numpy(), b, True, f"device in/out output {k}")


def test_cuda_graph_linker_option(env):
    """``mode(cuda_graph=True)``: the evaluation is captured once and replayed
    (runtime/graph.py) when the arguments are device-resident at repeating addresses."""
    aesara, L, lib = env
    from aesara_b200.runtime.device import DeviceArray

    build, inputs = _cfg("cfg1")
    i, o = build()
    f = aesara.function(i, o, mode=L.mode(device_outputs=True, cuda_graph=True))
    f.trust_input = True
    a, v, M = inputs()
    dv, dM = DeviceArray.from_numpy(v), DeviceArray.from_numpy(M)
    want = 1.0 + (M + a).dot(v)
    for _ in range(4):
        (out,) = f(np.float64(a), dv, dM)
    assert f.vm._replay is not None and f.vm._replay.replays >= 2
    np.testing.assert_allclose(out.to_numpy(), want, rtol=1e-12)


def test_profile_gets_cuda_event_times(env):
    aesara, L, lib = env
    build, inputs = _cfg("cfg3")
    i, o = build()
    f = aesara.function(i, o, mode=L.mode(), profile=True)
    vals = inputs()
    for _ in range(3):
        f(*vals)
    prof = f.profile
    assert prof.fct_callcount == 3
    times = {type(n.op).__name__: t for (fg, n), t in prof.apply_time.items() if t > 0}
    assert any(k in times for k in ("Dot22", "Gemm")), times
    assert sum(prof.apply_time.values()) > 0
    assert max(prof.apply_callcount.values()) == 3


def test_output_subset_on_device(env):
    aesara, L, lib = env
    import aesara.tensor as at

    x = at.fmatrix("x")
    cnt = aesara.shared(np.zeros((), "float32"), name="cnt")
    outs = [at.tanh(x), at.exp(x).sum(axis=0), x @ x.T]
    f = aesara.function([x], outs, updates=[(cnt, cnt + x.sum())], mode=L.mode())
    xv = np.random.default_rng(1).standard_normal((300, 200)).astype("float32")
    r = f(xv, output_subset=[1])
    assert len(r) == 1
    _close(r[0], np.exp(xv.astype(np.float64)).sum(0).astype("float32"), False, "subset output")
    np.testing.assert_allclose(cnt.get_value(), xv.sum(dtype=np.float64), rtol=1e-5)
    full = f(xv)
    assert len(full) == 3
    _close(full[2], xv @ xv.T, True, "full call after a subset call")


def test_errors_are_reraised_with_the_apply_node(env):
    """Shape errors come back as the reference's exception types through
    ``raise_with_op`` (link/utils.py:270) with the failing Apply node named."""
    aesara, L, lib = env
    import aesara.tensor as at

    x, y = at.fmatrix("x"), at.fmatrix("y")
    f = aesara.function([x, y], at.dot(x, y) + 1, mode=L.mode())
    with pytest.raises(ValueError) as ei:
        f(np.zeros((200, 30), "float32"), np.zeros((31, 200), "float32"))
    msg = str(ei.value)
    assert "Apply node that caused the error" in msg and "Inputs shapes" in msg
    assert f.vm.position_of_error >= 0
    # the function is still usable
    r = f(np.ones((200, 30), "float32"), np.ones((30, 200), "float32"))
    np.testing.assert_allclose(r, 31.0)


def test_dual_run_checker_against_c_thunks_on_device(env):
    """``debug.check_function``: every node of the device run vs the reference's C thunk of
    the same Apply (the DualLinker pattern, link/c/basic.py:1934)."""
    aesara, L, lib = env
    from aesara_b200.debug import check_function

    for name in ("cfg3", "cfg5"):
        build, inputs = _cfg(name)
        i, o = build()
        report = check_function(i, o, inputs(), reference_linker="c", rtol=2e-5)
        assert len(report) >= 5
        assert all(isinstance(r[2], float) for r in report)


def test_shared_constructor_takes_device_arrays(env):
    aesara, L, lib = env
    import aesara.tensor as at
    from aesara_b200.runtime.device import DeviceArray
    from aesara_b200.sharedvar import B200SharedVariable

    w = np.random.default_rng(2).standard_normal((500, 40)).astype("float32")
    W = aesara.shared(DeviceArray.from_numpy(w), name="W")
    assert isinstance(W, B200SharedVariable) and W.is_on_device()
    x = at.fmatrix("x")
    f = aesara.function([x], (x @ W).sum(axis=0), mode=L.mode())
    xv = np.random.default_rng(3).standard_normal((700, 500)).astype("float32")
    _close(f(xv), (xv.astype(np.float64) @ w).sum(0).astype("float32"), True, "x @ W(dev)", rtol=2e-5)


def test_scan_variants_through_the_linker(env):
    """Scan inner graphs are rewritten at lowering (lower.optimized_inner_fgraph), not by a
    side effect of another linker: a tanh-RNN with a gradient and a while-loop, compiled for
    the device FIRST, then by the C-linker."""
    aesara, L, lib = env
    import aesara.tensor as at

    x = at.ftensor3("x")
    h0 = at.fmatrix("h0")
    W = at.fmatrix("W")

    def step(x_t, h, W_):
        return at.tanh(x_t + at.dot(h, W_))

    hs, _ = aesara.scan(step, sequences=[x], outputs_info=[h0], non_sequences=[W])
    loss = (hs[-1] ** 2).sum()
    gW = aesara.grad(loss, W)
    f = aesara.function([x, h0, W], [loss, gW], mode=L.mode())
    g = aesara.function([x, h0, W], [loss, gW], mode="FAST_RUN")
    rng = np.random.default_rng(4)
    T, B, H = 7, 96, 48
    vals = [rng.standard_normal((T, B, H)).astype("float32") * 0.5,
            rng.standard_normal((B, H)).astype("float32") * 0.1,
            (rng.standard_normal((H, H)) / np.sqrt(H)).astype("float32")]
    for k, (a, b) in enumerate(zip(f(*vals), g(*vals))):
        _close(a, b, True, f"rnn grad output {k}", rtol=3e-5)


def test_host_chunks_pipeline_equals_whole_batch(env):
    """``mode(host_chunks=K)``: NumPy arguments of a batch-map graph are uploaded and evaluated in
    row blocks (shard.ChunkedHostExecutor) and the outputs combined as shardplan derives from the
    graph (mean for the MLP loss and gradients, concat for a per-row output); a graph that is not
    a batch map is evaluated whole."""
    aesara, L, lib = env
    import aesara.tensor as at
    import torch

    from aesara_b200 import graphs as G

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
        t.numpy()[...] = a
        return t

    i, o = G.cfg3_mlp()
    f = aesara.function(i, o, mode=L.mode(host_chunks=4))
    i2, o2 = G.cfg3_mlp()
    g = aesara.function(i2, o2, mode=L.mode())
    vals = G.cfg3_inputs(16384 + 37, 256)          # ragged: the row blocks differ by one row
    keep = [pinned(v) for v in vals]
    got = f(*[k.numpy() for k in keep])
    assert f.vm.executor.chunks_run == 4
    assert [m[0] for m in f.maker.linker.shard_plan.outputs] == ["mean"] * 5
    want = g(*vals)
    for k, (a, b) in enumerate(zip(got, want)):
        _close(a, b, True, f"chunked MLP output {k}", rtol=3e-5)
    # per-row output: concat
    x, y, z = at.fvectors("x", "y", "z")
    h = aesara.function([x, y, z], [at.softplus(at.tanh(x) + y) * z, (x * y).sum()], mode=L.mode(host_chunks=3))
    n = 50000
    xs = G.cfg2_inputs(n)
    out, tot = h(*xs)
    assert h.vm.executor.chunks_run == 3 and out.shape == (n,)
    ref = aesara.function([x, y, z], [at.softplus(at.tanh(x) + y) * z, (x * y).sum()], mode="FAST_RUN")(*xs)
    _close(out, ref[0], False, "chunked elementwise output")
    _close(tot, ref[1], True, "chunked sum output", rtol=2e-5)
    # device-resident arguments are evaluated in one piece
    from aesara_b200.runtime.device import DeviceArray

    f.trust_input = True
    f(*[DeviceArray.from_numpy(v) for v in vals])
    assert f.vm.executor.chunks_run == 1
    # not a batch map (a maximum over the batch axis): no chunking, same answer as the C-linker
    m = aesara.function([x], x.max() + x.sum(), mode=L.mode(host_chunks=4))
    assert not hasattr(m.vm.executor, "chunks_run")
    _close(m(xs[0]), np.float32(xs[0].max() + xs[0].sum(dtype=np.float64)), True, "replicas-only graph")
