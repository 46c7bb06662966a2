"""CPU (needs the reference front-end): the drop-in boundary.

`aesara.function(..., mode=B200)` must go through B200Linker.accept/make_all and
give Function-level semantics identical to the reference VM (storage cells,
shared-variable updates, output order).  No GPU here, so the *device executor*
is replaced by the oracle interpreter for these tests only — what is under test
is the linker/VM glue and the lowering, not the kernels (tests/test_gpu_parity.py
covers those on the B200 box)."""
import numpy as np
import pytest

from aesara_b200.compat import bootstrap

pytestmark = pytest.mark.skipif(not bootstrap.available(), reason="reference front-end not available")


@pytest.fixture()
def aes():
    aesara = bootstrap.load_aesara()
    import aesara_b200.linker as L

    return aesara, L


class _OracleExecutor:
    """Stand-in with the ProgramExecutor call signature (tests only)."""

    def __init__(self, program, **kw):
        self.program = program

    def __call__(self, *inputs):
        from oracle.program_np import run_program

        return run_program(self.program, [np.array(i) if isinstance(i, np.ndarray) else i for i in inputs])


def test_mode_and_linker_are_registered(aes):
    aesara, L = aes
    from aesara.compile.mode import get_mode, predefined_linkers

    assert "b200" in predefined_linkers
    m = get_mode("B200")
    assert isinstance(m.linker, L.B200Linker)
    from aesara.compile.mode import get_target_language

    assert get_target_language(m) == ("c",)


@pytest.mark.parametrize("cfg", ["cfg2_fused", "cfg3_mlp", "cfg5_logreg", "cfg4_lstm"])
def test_lowering_reproduces_committed_fixture(aes, cfg):
    """The program the linker lowers today equals the committed fixture (so the
    fixtures the GPU box runs are what the linker would execute)."""
    from aesara_b200 import graphs as G
    from tests._cases import load_case

    build = {"cfg2_fused": G.cfg2_fused_elemwise, "cfg3_mlp": G.cfg3_mlp,
             "cfg5_logreg": G.cfg5_logreg, "cfg4_lstm": G.cfg4_lstm_scan}[cfg]
    i, o = build()
    prog, _ = G.optimized_program(i, o, name=cfg)
    want, _, _ = load_case(cfg)
    assert [n.op for n in prog.nodes] == [n.op for n in want.nodes]
    assert [n.params.get("expr", {}).get("name") for n in prog.nodes] == [
        n.params.get("expr", {}).get("name") for n in want.nodes
    ]


def test_function_semantics_through_the_linker(aes, monkeypatch):
    aesara, L = aes
    import aesara.tensor as at
    import aesara_b200.runtime.vm as vm

    monkeypatch.setattr(vm, "ProgramExecutor", _OracleExecutor)
    x = at.fvector("x")
    acc = aesara.shared(np.zeros(5, "float32"), name="acc")
    out = at.tanh(x) * 2 + acc
    f = aesara.function([x], [out, out.sum()], updates=[(acc, acc + x)], mode=L.mode())
    assert isinstance(f.vm, L.B200VM)
    g = aesara.function([x], [out, out.sum()], updates=[], mode="FAST_RUN")
    xv = np.arange(5, dtype="float32")
    r1 = f(xv)
    np.testing.assert_allclose(r1[0], np.tanh(xv) * 2, rtol=1e-6)
    np.testing.assert_allclose(acc.get_value(), xv)          # update applied by the VM
    r2 = f(xv)
    np.testing.assert_allclose(r2[0], np.tanh(xv) * 2 + xv, rtol=1e-6)
    np.testing.assert_allclose(acc.get_value(), 2 * xv)
    assert len(r2) == 2 and r2[1].shape == ()
    acc.set_value(np.zeros(5, "float32"))
    np.testing.assert_allclose(g(xv)[0], r1[0], rtol=1e-6)


def test_unsupported_op_is_a_hard_error(aes):
    """No CPU fallback for tensor work: an Op without a device implementation
    fails at link time, naming the Op."""
    aesara, L = aes
    import aesara.tensor as at

    x = at.fmatrix("x")
    with pytest.raises(NotImplementedError, match="no device implementation"):
        aesara.function([x], at.linalg.det(x), mode=L.mode())


def test_linker_copies_bind_to_one_graph(aes):
    aesara, L = aes
    import aesara.tensor as at
    from aesara.graph.fg import FunctionGraph

    x = at.fvector("x")
    fg1 = FunctionGraph([x], [x * 2], clone=True)
    fg2 = FunctionGraph([x], [x + 1], clone=True)
    lk = L.B200Linker()
    a = lk.accept(fg1)
    b = a.accept(fg2)
    assert a is lk and b is not lk and b.fgraph is fg2
    assert lk.clone(allow_gc=False).allow_gc is False
