"""Drop-in replacement for the reference's `data_utils/ode.py` (solve_ode_batch, gen_data).

The reference integrates an arbitrary Python right-hand side with NumPy float64 RK4, one Python iteration and
four vectorised RHS calls per step (`data_utils/ode.py:7-28`). Every ODE the reference ships is a member of the
SINDy library — f(x) = Θ(x)·Ξᵀ with the truth Ξ of `evaluation/eval_eq.py:88-105` — so here the right-hand side
is a `LibraryODE` (library + coefficient matrix) and the whole rollout is ONE CUDA kernel in float64
(`sb_rollout`, record_dx mode). A plain Python callable cannot run on the device: it is rejected, there is no
CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from sindy_b200 import native
from sindy_b200.native import Library

__all__ = ["LibraryODE", "solve_ode_batch", "gen_data"]


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


def solve_ode_batch(ode, x0, dt=0.002, num_steps=2000, solver='rk4', device=None, return_tensors=False, **kwargs):
    """x, dx of shape (num_steps, *x0.shape), float64: row i is the state after i RK4 steps and its derivative
    (row 0 = x0), exactly the layout of the reference. `ode` must be a LibraryODE."""
    if solver != 'rk4':
        raise NotImplementedError
    if not isinstance(ode, LibraryODE):
        raise TypeError("solve_ode_batch needs a LibraryODE (library + coefficients); arbitrary Python "
                        "right-hand sides cannot run on the GPU and there is no CPU fallback")
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
    x, dx = solve_ode_batch(ode, x0, dt=dt, num_steps=num_steps)
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
