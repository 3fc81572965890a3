"""Drop-in replacement for the reference's `data_utils/ode.py` (solve_ode_batch, gen_data).

The reference integrates an arbitrary Python right-hand side with NumPy float64 RK4, one Python iteration and
four vectorised RHS calls per step (`data_utils/ode.py:7-28`). Every ODE the reference ships is a member of the
SINDy library — f(x) = Θ(x)·Ξᵀ with the truth Ξ of `evaluation/eval_eq.py:88-105` — so here the right-hand side
is a `LibraryODE` (library + coefficient matrix) and the whole rollout is ONE CUDA kernel in float64
(`sb_rollout`, record_dx mode).

The reference's generators (`data_utils/damped_oscillator.py:20-24`, `growth.py:18-22`, `lotka.py:33-41`,
`selkov.py:18-22`) hand `solve_ode_batch` a Python callable. A Python callable cannot run on the device, so it is
IDENTIFIED first (`identify_library_ode`): evaluated once on a few hundred probe points, fitted in float64 onto the
candidate libraries, verified on held-out points to 1e-11 — if it is a member of a SINDy library (all four shipped
systems are) the rollout runs on the GPU with the recovered coefficients; if not, `TypeError` (there is no CPU
fallback for the integration itself).
"""
from __future__ import annotations

import numpy as np
import torch

from sindy_b200 import native
from sindy_b200.native import Library

__all__ = ["LibraryODE", "solve_ode_batch", "gen_data", "identify_library_ode"]


class LibraryODE:
    """dx/dt = Θ(x)·Ξᵀ. Callable on NumPy arrays (host evaluation for samplers/tests of small inputs is NOT
    provided: calling it runs the CUDA forward kernel)."""

    def __init__(self, library: Library, Xi, name: str = "ode"):
        self.library = library
        self.Xi = np.asarray(Xi, dtype=np.float64)
        if self.Xi.shape != (library.dim, library.K):
            raise ValueError(f"Xi has shape {self.Xi.shape}, expected {(library.dim, library.K)}")
        self.name = name

    def with_coefficients(self, Xi):
        return LibraryODE(self.library, Xi, self.name)

    def __call__(self, x, **kwargs):
        dev = torch.device("cuda", torch.cuda.current_device())
        xt = torch.as_tensor(np.asarray(x), dtype=torch.float64, device=dev)
        w = torch.as_tensor(self.Xi, dtype=torch.float64, device=dev)
        shape = xt.shape
        _, dx, _ = native.rollout(xt.reshape(-1, shape[-1]), w, self.library, 0.0, 1, 1, "rk4", record_dx=True,
                                  want_traj=True, want_last=False)
        return dx[0].reshape(shape).cpu().numpy()


# candidate libraries, specialised kernels first: (poly_order, include_sine, include_exp)
_CANDIDATES = [(2, False, False), (3, False, False), (5, False, False), (1, False, False), (4, False, False),
               (2, False, True), (3, False, True), (2, True, False), (3, True, False), (2, True, True),
               (3, True, True)]


def _theta_host(x, lib):
    """Θ(x) in float64 on the host for the PROBE points of identify_library_ode only (a few hundred rows): the
    reference's column order (`sindy.py:7-30`), products formed from the exponent table."""
    e = lib.exponents().numpy()                                   # (K, d); transcendental rows are all-zero
    n_poly = lib.K - lib.dim * (int(lib.include_sine) + int(lib.include_exp))
    cols = [np.prod(x[:, None, :] ** e[None, :n_poly, :], axis=2)]
    if lib.include_sine:
        cols.append(np.sin(x))
    if lib.include_exp:
        cols.append(np.exp(x))
    return np.concatenate(cols, axis=1)


def identify_library_ode(ode, x0, name=None, **kwargs):
    """LibraryODE equal to the Python right-hand side `ode(x, **kwargs)` (reference signature: x (..., d) -> dx).

    `ode` is called twice on the host, on 4·K+64 fit points and 256 verification points drawn from the bounding box
    of the initial conditions; the coefficients come from a float64 least-squares fit, re-fitted on their support.
    Raises TypeError if no candidate library reproduces the callable to 1e-11 (relative to its largest output)."""
    x0 = np.asarray(x0, dtype=np.float64).reshape(-1, np.asarray(x0).shape[-1])
    d = x0.shape[1]
    lo, hi = x0.min(0), x0.max(0)
    pad = np.maximum(0.5 - (hi - lo) / 2, 0.05)                   # box at least 1 wide per coordinate
    lo, hi = lo - pad, hi + pad
    rng = np.random.default_rng(20240516)
    fn = lambda q: np.asarray(ode(q.copy(), **kwargs), dtype=np.float64)  # noqa: E731
    x_ver = rng.uniform(lo, hi, (256, d))
    y_ver = fn(x_ver)
    scale = max(float(np.abs(y_ver).max()), 1e-300)
    for p, sine, exp_ in _CANDIDATES:
        lib = Library(d, p, sine, exp_)
        K = lib.K
        x_fit = rng.uniform(lo, hi, (4 * K + 64, d))
        y_fit = fn(x_fit)
        th = _theta_host(x_fit, lib)
        coef = np.linalg.lstsq(th, y_fit, rcond=None)[0]          # (K, d)
        Xi = np.zeros((d, K))
        for i in range(d):                                        # re-fit on the support: well conditioned, ~1e-16
            sup = np.abs(coef[:, i]) > 1e-9 * max(np.abs(coef[:, i]).max(), 1e-300)
            if sup.any():
                Xi[i, sup] = np.linalg.lstsq(th[:, sup], y_fit[:, i], rcond=None)[0]
        if np.abs(_theta_host(x_ver, lib) @ Xi.T - y_ver).max() <= 1e-11 * scale:
            return LibraryODE(lib, Xi, name or getattr(ode, "__name__", "ode"))
    raise TypeError("solve_ode_batch: the right-hand side is not a member of a supported SINDy library (polynomial "
                    "degree <= 5, optionally sin/exp columns); arbitrary Python callables cannot run on the GPU and "
                    "there is no CPU fallback")


def solve_ode_batch(ode, x0, dt=0.002, num_steps=2000, solver='rk4', device=None, return_tensors=False, **kwargs):
    """x, dx of shape (num_steps, *x0.shape), float64: row i is the state after i RK4 steps and its derivative
    (row 0 = x0), exactly the layout of the reference (`data_utils/ode.py:7-28`). `ode`: a LibraryODE, or a Python
    right-hand side `ode(x, **kwargs)` that is a member of the SINDy library (identified, see above)."""
    if solver != 'rk4':
        raise NotImplementedError
    if not isinstance(ode, LibraryODE):
        if not callable(ode):
            raise TypeError("solve_ode_batch: `ode` must be a LibraryODE or a callable right-hand side")
        ode = identify_library_ode(ode, x0.detach().cpu().numpy() if torch.is_tensor(x0) else x0, **kwargs)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    x0t = torch.as_tensor(np.asarray(x0) if not torch.is_tensor(x0) else x0).to(device=dev, dtype=torch.float64)
    w = torch.as_tensor(ode.Xi, dtype=torch.float64, device=dev)
    x, dx, _ = native.rollout(x0t, w, ode.library, dt, num_steps, 1, "rk4", record_dx=True, want_traj=True,
                              want_last=False)
    x = x.view(num_steps, *x0t.shape)
    dx = dx.view(num_steps, *x0t.shape)
    if return_tensors:
        return x, dx
    return x.cpu().numpy(), dx.cpu().numpy()


def gen_data(ode, init_fn, n_ics=1000, dt=0.002, num_steps=2000, subsample_rate=1, noise=0.0,
             multiplicative_noise=False, smoothing=None, **kwargs):
    """Trajectories (n_ics, num_steps/subsample_rate, dim) and derivatives, as the reference's gen_data
    (`data_utils/ode.py:30-49`): noise and finite differences on the host; `smoothing='gp'` runs the GP smoother of
    data_utils/smoothing.py on the GPU (needs kwargs['gp_sigma_in'] like the reference)."""
    x0 = init_fn(n_ics)
    x, dx = solve_ode_batch(ode, x0, dt=dt, num_steps=num_steps, **kwargs)
    if noise > 0:
        x_std = np.std(x, axis=(0, 1))
        if multiplicative_noise:
            x *= (1 + np.random.randn(*x.shape) * noise)
        else:
            x += np.random.randn(*x.shape) * noise * x_std
        if smoothing is None:
            dx[:-1, :] = np.diff(x, axis=0) / dt
        elif smoothing == 'gp':
            from .smoothing import num_diff_gp
            print('Smoothing with Gaussian process...')
            dx, x = num_diff_gp(x, dt, noise_level=noise, std_base=x_std, sigma_in=kwargs['gp_sigma_in'])
    x = np.transpose(x[::subsample_rate], (1, 0, 2))
    dx = np.transpose(dx[::subsample_rate], (1, 0, 2))
    return x, dx
