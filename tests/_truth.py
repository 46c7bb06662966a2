"""float64 ground truth for the BASELINE graphs at BASELINE sizes (checker code only).

The NumPy oracle and the reference C-linker need tens of seconds to minutes at these sizes
(cfg3: ~30 s per evaluation on 128 cores), and both are float32: what the 1e-5 bar is argued
against at full size is a float64 evaluation of the same mathematics.  It runs on the
GPU through torch.float64 — torch is used here as a *checker*, in row chunks so that the
float64 intermediates fit; nothing on the product path or inside a timed region calls it.
"""
import numpy as np
import torch


def nerr(got, want):
    """Norm-wise error max|got - want| / max|want| (float64)."""
    g = torch.as_tensor(np.asarray(got, np.float64)) if not torch.is_tensor(got) else got.double().cpu()
    w = want.double().cpu() if torch.is_tensor(want) else torch.as_tensor(np.asarray(want, np.float64))
    return float((g - w).abs().max() / w.abs().max().clamp_min(1e-300))


def mlp_truth(X, Y, W1, b1, W2, b2, blocks, chunk=8192):
    """cfg3: loss, db1, db2 in full and dW1 / dW2 on the given [r0, c0] 128x128
    blocks, all float64, accumulated over row chunks.  Inputs are float32 CUDA tensors."""
    B, H = X.shape
    W1d, W2d, b1d, b2d = W1.double(), W2.double(), b1.double(), b2.double()
    n = float(B * W2.shape[1])
    loss = torch.zeros((), dtype=torch.float64, device=X.device)
    db1 = torch.zeros(W1.shape[1], dtype=torch.float64, device=X.device)
    db2 = torch.zeros(W2.shape[1], dtype=torch.float64, device=X.device)
    dW1 = [torch.zeros(128, 128, dtype=torch.float64, device=X.device) for _ in blocks]
    dW2 = [torch.zeros(128, 128, dtype=torch.float64, device=X.device) for _ in blocks]
    for s in range(0, B, chunk):
        x = X[s:s + chunk].double()
        h = torch.tanh(x @ W1d + b1d)
        diff = h @ W2d + b2d - Y[s:s + chunk].double()
        loss += (diff * diff).sum()
        dout = diff * (2.0 / n)
        db2 += dout.sum(0)
        dpre = (dout @ W2d.T) * (1.0 - h * h)
        db1 += dpre.sum(0)
        for k, (r0, c0) in enumerate(blocks):
            dW2[k] += h[:, r0:r0 + 128].T @ dout[:, c0:c0 + 128]
            dW1[k] += x[:, r0:r0 + 128].T @ dpre[:, c0:c0 + 128]
        del x, h, diff, dout, dpre
    return {"loss": loss / n, "db1": db1, "db2": db2, "dW1": dW1, "dW2": dW2}


def logreg_truth(X, y, w, b, chunk=1 << 20):
    """cfg5: cost, grad_w, grad_b in float64 over row chunks."""
    N = X.shape[0]
    wd = w.double()
    cost = torch.zeros((), dtype=torch.float64, device=X.device)
    gw = torch.zeros(X.shape[1], dtype=torch.float64, device=X.device)
    gb = torch.zeros((), dtype=torch.float64, device=X.device)
    for s in range(0, N, chunk):
        x = X[s:s + chunk].double()
        yy = y[s:s + chunk].double()
        z = x @ wd + float(b)
        # -y log p - (1-y) log(1-p) = softplus(z) - y z
        cost += (torch.nn.functional.softplus(z) - yy * z).sum()
        r = (torch.sigmoid(z) - yy) / N
        gw += x.T @ r
        gb += r.sum()
        del x, yy, z, r
    return {"cost": cost / N, "grad_w": gw, "grad_b": gb}
