"""Shared helpers: load golden fixtures, compare with the tolerances stated here.

Tolerance policy (BASELINE north star; the reference's own test tolerances are
aesara/tensor/math.py:85-96: f32 rtol=atol=1e-5, f64 rtol=1e-5/atol=1e-8):
  * integer / bool outputs: bit-exact;
  * float outputs of element-wise / reduction graphs: rtol 1e-5 (+ atol 1e-6 f32, 1e-10 f64);
  * float outputs downstream of a dot product: rtol 1e-5 relative to the
    largest magnitude of the output (norm-wise), because the summation order
    differs between BLAS, NumPy and the device kernels.
"""
import glob
import os

import numpy as np

from aesara_b200.ir import Program

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BLAS_OPS = {"Dot22", "Dot22Scalar", "Gemm", "Gemv", "Ger", "Dot", "Scan"}


def case_names():
    return sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN, "*.json")))


def load_case(name):
    prog = Program.load(os.path.join(GOLDEN, name + ".json"))
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    ins = [z[f"in_{k}"] for k in range(len(prog.inputs))]
    outs = [z[f"out_{k}"] for k in range(len(prog.outputs))]
    return prog, ins, outs


def uses_blas(prog):
    return any(n.op in BLAS_OPS for n in prog.nodes)


def assert_matches(got, want, blas=False, rtol=1e-5, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype, f"{what}: dtype {got.dtype} != {want.dtype}"
    if want.dtype.kind in "biu":
        np.testing.assert_array_equal(got, want, err_msg=what)
        return
    if blas:
        scale = float(np.max(np.abs(want))) if want.size else 1.0
        atol = rtol * max(scale, 1e-30)
        np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, err_msg=what)
    else:
        atol = 1e-6 if want.dtype == np.float32 else 1e-10
        np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, equal_nan=True, err_msg=what)
