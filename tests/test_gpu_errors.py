"""GPU: error behaviour at the boundary.  The reference raises from its
thunks with these exception types / messages (tensor/blas.py:645-665, blas_c.py:381-392,
elemwise_cgen.py:116, subtensor.py, scan/op.py:1740-1760, raise_op.py); the executor raises
the same types on the host before (or instead of) launching, wrapped in NodeError that records
the failing position — what B200VM turns into `position_of_error` for `raise_with_op`."""
import numpy as np
import pytest

from tests._cases import load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt():
    import torch

    from aesara_b200.runtime import lib
    from aesara_b200.runtime.vm import ProgramExecutor

    lib.check(lib.load().ab_init(0))
    torch.cuda.set_device(0)
    return ProgramExecutor


def _raises(ex, ins, exc, match):
    from aesara_b200.runtime.vm import NodeError

    with pytest.raises(NodeError) as ei:
        ex(*ins)
    err = ei.value
    assert isinstance(err.original, exc), f"expected {exc.__name__}, got {type(err.original).__name__}: {err.original}"
    assert match.lower() in str(err.original).lower(), str(err.original)
    assert 0 <= err.position < len(ex.program.nodes) and ex.position_of_error == err.position
    return err


def test_gemv_shape_error(rt):
    prog, ins, _ = load_case("cfg1_readme")
    a, v, M = [np.array(x) for x in ins]
    _raises(rt(prog), [a, v[:-3], M], ValueError, "Incompatible shapes for gemv")


def test_dot22_and_gemm_shape_errors(rt):
    prog, ins, _ = load_case("blas_dot22_layouts")
    a, b, c = [np.array(x) for x in ins]
    _raises(rt(prog), [a, b[:-1], c], ValueError, "Shape mismatch: x has")
    prog, ins, _ = load_case("blas_gemm_alpha_beta")
    z, x, y, al, be = [np.array(v) for v in ins]
    _raises(rt(prog), [z[:-2], x, y, al, be], ValueError, "mismatch")


def test_elemwise_broadcast_error(rt):
    prog, ins, _ = load_case("ew_fusion_multi")
    x, y, z, iv = [np.array(v) for v in ins]
    _raises(rt(prog), [x, y[:, :-1], z, iv], ValueError, "Input dimension mismatch")


def test_index_errors(rt):
    prog, ins, _ = load_case("indexing_embedding")
    vals = [np.array(v) for v in ins]
    bad = list(vals)
    bad[1] = vals[1].copy()
    bad[1][0] = 1000  # row index far out of range
    _raises(rt(prog), bad, IndexError, "out of")
    prog, ins, _ = load_case("adv_index_pairs")
    vals = [np.array(v) for v in ins]
    bad = list(vals)
    bad[3] = vals[3].copy()
    bad[3][2] = 99  # column index of x[i, j]
    _raises(rt(prog), bad, IndexError, "out of")
    mism = list(vals)
    mism[3] = vals[3][:-2]  # i and j of different lengths
    _raises(rt(prog), mism, IndexError, "broadcast")


def test_scan_too_many_steps(rt):
    prog, ins, _ = load_case("scan_seq_taps_shared_nsteps")
    x, k = np.array(ins[0]), ins[1]
    _raises(rt(prog), [x, np.int64(x.shape[0] + 5)], ValueError, "required number of steps")


def test_reshape_size_error(rt):
    prog, ins, _ = load_case("reshape_flatten_noncontig")
    x, t, n = np.array(ins[0]), np.array(ins[1]), ins[2]
    _raises(rt(prog), [x, t, np.int64(7)], ValueError, "reshape")  # 60 elements into (7, -1)


def test_reduction_over_empty_axis_without_identity(rt):
    prog, ins, _ = load_case("careduce_nan")
    _raises(rt(prog), [np.zeros((0, 9), "float32")], ValueError, "zero-size array to reduction")


def test_join_dimension_error(rt):
    prog, ins, _ = load_case("join_split_reshape")
    a, b, c, v = [np.array(x) for x in ins]
    _raises(rt(prog), [a, b[:, :-1], c, v], ValueError, "must match")


def test_wrong_rank_and_arity_are_type_errors(rt):
    prog, ins, _ = load_case("cfg2_fused")
    ex = rt(prog)
    with pytest.raises(TypeError, match="dimensions"):
        ex(np.zeros((2, 2), "float32"), np.zeros(4, "float32"), np.zeros(4, "float32"))
    with pytest.raises(TypeError, match="expected 3 inputs"):
        ex(np.zeros(4, "float32"))
