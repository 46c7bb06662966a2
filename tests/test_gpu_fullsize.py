"""GPU: independent truth at BASELINE sizes (VERDICT r1 "compared only with themselves").

The driver-benchmarked computations — cfg3 at B=65536, H=4096 under BOTH GEMM policies, cfg4 at
T=128, B=8192, H=1024, cfg5 at N=2^24, D=512 — are compared with something that is not this
backend: a float64 evaluation (tests/_truth.py) or, for the Scan, the NumPy oracle on sampled
batch rows (rows are independent through all T steps).  Measured errors are printed (-s) and
written to gpurun_out/fullsize_errors.json.
"""
import json
import os

import numpy as np
import pytest

from tests._cases import load_case
from tests._truth import logreg_truth, mlp_truth, nerr

pytestmark = pytest.mark.gpu
_ERRORS = {}


@pytest.fixture(scope="module")
def rt():
    import torch

    from aesara_b200.runtime import lib
    from aesara_b200.runtime.vm import ProgramExecutor

    lib.check(lib.load().ab_init(0))
    torch.cuda.set_device(0)
    return ProgramExecutor


def _record(name, errs):
    _ERRORS[name] = errs
    print(f"\n{name}: " + ", ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "fullsize_errors.json"), "w") as f:
        json.dump(_ERRORS, f, indent=1, sort_keys=True)


@pytest.mark.parametrize("precision,tol", [(0, 1e-5), (2, 2e-2)])
def test_cfg3_full_size_vs_float64(rt, precision, tol):
    """loss, db1, db2 in full; dW1 and dW2 on five 128x128 blocks (corners, centre, a random
    one).  fp32-faithful (3xTF32, segmented accumulation) <= 1e-5 norm-wise even for the
    K = 65536 weight-gradient products; the bf16 policy inside its stated 2e-2."""
    import torch

    from aesara_b200.runtime.device import DeviceArray

    prog, _, _ = load_case("cfg3_mlp")
    B, H = 65536, 4096
    g = torch.Generator(device="cuda").manual_seed(17)
    X = torch.randn(B, H, device="cuda", generator=g)
    Y = torch.randn(B, H, device="cuda", generator=g)
    W1 = torch.randn(H, H, device="cuda", generator=g) / H ** 0.5
    W2 = torch.randn(H, H, device="cuda", generator=g) / H ** 0.5
    b1 = torch.randn(H, device="cuda", generator=g) * 0.1
    b2 = torch.randn(H, device="cuda", generator=g) * 0.1
    ex = rt(prog, precision=precision, host_outputs=False)
    loss, dW1, db1, dW2, db2 = ex(*[DeviceArray.from_torch(t) for t in (X, Y, W1, b1, W2, b2)])
    # bf16 policy: the three GEMM-epilogue regions; fp32-faithful: node by node but for Sqr -> Sum
    assert ex.fused_regions_run == (3 if precision == 2 else 1)
    blocks = [(0, 0), (H - 128, H - 128), (0, H - 128), (H // 2, H // 2 - 128), (1152, 2944)]
    truth = mlp_truth(X, Y, W1, b1, W2, b2, blocks)
    dW1t = dW1.owner.view(torch.float32)[: H * H].view(H, H)
    dW2t = dW2.owner.view(torch.float32)[: H * H].view(H, H)
    errs = {"loss": abs(float(np.asarray(loss)) - float(truth["loss"])) / float(truth["loss"]),
            "db1": nerr(np.asarray(db1), truth["db1"]), "db2": nerr(np.asarray(db2), truth["db2"])}
    # blocks: error relative to the largest entry of the whole gradient (norm-wise, as elsewhere)
    for nm, dev, tr in (("dW1", dW1t, truth["dW1"]), ("dW2", dW2t, truth["dW2"])):
        scale = float(dev.abs().max())
        worst = 0.0
        for (r0, c0), t in zip(blocks, tr):
            worst = max(worst, float((dev[r0:r0 + 128, c0:c0 + 128].double() - t).abs().max()) / scale)
        errs[nm] = worst
    _record(f"cfg3_B65536_H4096_precision{precision}", errs)
    for k, v in errs.items():
        assert v <= tol, f"{k}: {v:.3e} > {tol}"


def test_cfg4_full_size_sampled_rows_vs_oracle(rt):
    """h_T, c_T of 64 sampled batch rows after 128 steps vs the NumPy oracle on those rows."""
    import torch

    from aesara_b200.runtime.device import DeviceArray
    from oracle.program_np import run_program

    prog, _, _ = load_case("cfg4_lstm")
    T, B, H = 128, 8192, 1024
    g = torch.Generator(device="cuda").manual_seed(23)
    x = torch.randn(T, B, 4 * H, device="cuda", generator=g)
    U = torch.randn(H, 4 * H, device="cuda", generator=g) / H ** 0.5
    h0 = torch.zeros(B, H, device="cuda")
    c0 = torch.zeros(B, H, device="cuda")
    ex = rt(prog, host_outputs=False)
    hT, cT = ex(*[DeviceArray.from_torch(t) for t in (x, h0, c0, U)])
    scan = [st["runner"] for st in ex._state if "runner" in st][0]
    assert scan.used_fast_path
    rows = np.sort(np.random.default_rng(5).choice(B, size=64, replace=False))
    ridx = torch.as_tensor(rows, device="cuda")
    xs = x[:, ridx, :].cpu().numpy()
    want_h, want_c = run_program(prog, [xs, np.zeros((64, H), "float32"), np.zeros((64, H), "float32"),
                                        U.cpu().numpy()])
    got_h = hT.to_numpy()[rows]      # hs[-1] / cs[-1] are views into the Scan's output rings
    got_c = cT.to_numpy()[rows]
    errs = {"h_T": nerr(got_h, want_h), "c_T": nerr(got_c, want_c)}
    _record("cfg4_T128_B8192_H1024", errs)
    assert errs["h_T"] <= 1e-5 and errs["c_T"] <= 1e-5, errs


def test_cfg5_full_size_vs_float64(rt):
    import torch

    from aesara_b200.runtime.device import DeviceArray

    prog, _, _ = load_case("cfg5_logreg")
    N, D = 1 << 24, 512
    g = torch.Generator(device="cuda").manual_seed(29)
    X = torch.randn(N, D, device="cuda", generator=g)
    y = (torch.rand(N, device="cuda", generator=g) < 0.5).float()
    w = torch.randn(D, device="cuda", generator=g) * 0.01
    ex = rt(prog, host_outputs=False)
    cost, gw, gb = ex(DeviceArray.from_torch(X), DeviceArray.from_torch(y), DeviceArray.from_torch(w), np.float32(0.05))
    assert ex.fused_regions_run == 1
    truth = logreg_truth(X, y, w, 0.05)
    errs = {"cost": abs(float(np.asarray(cost)) - float(truth["cost"])) / float(truth["cost"]),
            "grad_w": nerr(np.asarray(gw), truth["grad_w"]),
            "grad_b": abs(float(np.asarray(gb)) - float(truth["grad_b"])) / max(abs(float(truth["grad_b"])),
                                                                                 float(truth["grad_w"].abs().max()))}
    _record("cfg5_N16777216_D512", errs)
    assert errs["cost"] <= 1e-5 and errs["grad_w"] <= 1e-5 and errs["grad_b"] <= 1e-5, errs
