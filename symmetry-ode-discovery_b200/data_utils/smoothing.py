"""GP smoother of the data generators on the GPU, call-compatible with the reference's
`data_utils/smoothing.py:155-196` (`num_diff_gp`, used by `gen_data` when `--smoothing gp`, `data_utils/ode.py:43-45`).

The reference builds a GPPCA model per state dimension with r = n_traj components (`smoothing.py:181-182`) — all of
them, so the factor loading A is a full orthogonal matrix, A·Aᵀ = I, and the predictive mean
`K_new (K + σ²I)⁻¹ Y A Aᵀ` (`:137-143`) is plain GP regression with the RBF kernel. Its 10⁴×10⁴ float64 `inv`/`eigh`
calls (≈ 5 min per file on the CPU) become one Cholesky factorisation and two kernel-matrix
products per dimension in float64 on the device (cuSOLVER / cuBLAS through torch: library calls, no custom kernel).
No CPU fallback: raises without a CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["num_diff_gp"]

_EPS = 0.001   # forward-difference step of the posterior mean (`smoothing.py:186`)


def _rbf(ta: torch.Tensor, tb: torch.Tensor, sigma_out: float, sigma_in: float) -> torch.Tensor:
    return sigma_out ** 2 * torch.exp(-(ta[:, None] - tb[None, :]) ** 2 / (2.0 * sigma_in ** 2))


def num_diff_gp(x, dt, noise_level, std_base, sigma_in=None, device=None):
    """x: (seq_len, n_trajs, input_dim) noisy states (numpy or tensor). Returns (dX, X_smooth) as float64 numpy arrays
    of the same shape, like the reference (note the order)."""
    if not torch.cuda.is_available():
        raise RuntimeError("data_utils.smoothing.num_diff_gp runs on the GPU (cuSOLVER Cholesky); no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    xt = torch.as_tensor(np.asarray(x), dtype=torch.float64).to(dev)
    T, n_traj, dim = xt.shape
    t = torch.arange(T, dtype=torch.float64, device=dev) * float(dt)
    s_in = float(dt) if sigma_in is None else float(sigma_in)
    dX, Xs = torch.empty_like(xt), torch.empty_like(xt)
    for d in range(dim):
        s_out = float(std_base[d])
        s = float(noise_level) * s_out
        K = _rbf(t, t, s_out, s_in)
        A = K.clone()
        A.diagonal().add_(s * s)
        L = torch.linalg.cholesky(A)
        del A
        alpha = torch.cholesky_solve(xt[:, :, d].contiguous(), L)
        del L
        X0 = K @ alpha
        del K
        X1 = _rbf(t + _EPS, t, s_out, s_in) @ alpha
        Xs[:, :, d] = X0
        dX[:, :, d] = (X1 - X0) / _EPS
    return dX.cpu().numpy(), Xs.cpu().numpy()
