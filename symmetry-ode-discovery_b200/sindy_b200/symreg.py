"""Linear Lie-derivative symmetry regulariser (reference `train.py:503-507`, intended formula `jvp(...)[1]`):

    L_sym(Ξ) = Σ_v Σ_n ‖ J_h(z_n)·(v z_n) − v·h(z_n) ‖²        (sum over the batch, one term per generator v)

Two evaluations, both differentiable w.r.t. Ξ:

* `lie_loss_per_sample` — literal: J_h(z)(vz) through the CUDA JVP operator (any library, incl. sin/exp columns);
* `lie_loss_from_gram`  — for polynomial libraries J_Θ(z)·v·z = M_v·Θ(z) with a constant K×K matrix M_v (the map
  the reference builds symbolically in `sindy.py:123-144`), hence the defect is (W M_v − v W)·Θ(z) =: A_v·Θ(z) and
  L_sym = Σ_v tr(A_v G A_vᵀ) with the Gram matrix G = ΘᵀΘ: ONE data pass (the moment kernel) serves every
  generator, and loss and gradient are K×K algebra in fp64.
"""
from __future__ import annotations

import itertools
from typing import Sequence

import numpy as np
import torch

from . import native, ops
from .native import Library


def lie_matrix(lib: Library, v) -> torch.Tensor:
    """M_v (K×K, fp64) with J_Θ(z)·v·z = M_v·Θ(z) for a polynomial library."""
    if lib.include_sine or lib.include_exp:
        raise ValueError("the Gram form of the Lie-derivative regulariser needs a polynomial library")
    d, p = lib.dim, lib.poly_order
    cols = [()]
    for n in range(1, p + 1):
        cols.extend(itertools.combinations_with_replacement(range(d), n))
    where = {c: k for k, c in enumerate(cols)}
    vm = np.asarray(torch.as_tensor(v).detach().cpu(), dtype=np.float64)
    M = np.zeros((len(cols), len(cols)))
    for k, c in enumerate(cols):
        for pos, j in enumerate(c):
            rest = c[:pos] + c[pos + 1:]
            for l in range(d):
                if vm[j, l] != 0.0:
                    M[k, where[tuple(sorted(rest + (l,)))]] += vm[j, l]
    return torch.from_numpy(M)


def gram(x: torch.Tensor, lib: Library) -> torch.Tensor:
    """G = Θ(x)ᵀΘ(x) (K×K fp64) in one pass (moment kernel for the specialised polynomial libraries)."""
    flags = native.SB_STEP_GRAM
    out = native.train_step(x, None, None, lib, flags)
    return native.unpack_step(out, lib, flags)["gram"]


def lie_loss_from_gram(G: torch.Tensor, W: torch.Tensor, gens: Sequence[torch.Tensor],
                       Ms: Sequence[torch.Tensor]) -> torch.Tensor:
    """Σ_v tr(A_v G A_vᵀ), A_v = W M_v − v W; differentiable w.r.t. W (fp64 result)."""
    Wd = W.double()
    loss = Wd.new_zeros(())
    for v, M in zip(gens, Ms):
        A = Wd @ M.to(Wd.device) - v.to(Wd.device, torch.float64) @ Wd
        loss = loss + ((A @ G) * A).sum()
    return loss


def quadratic_form(lib: Library, gens: Sequence[torch.Tensor], G: torch.Tensor) -> torch.Tensor:
    """H ((d·K)×(d·K), fp64, symmetric) with Σ_v tr(A_v G A_vᵀ) = wᵀ H w for w = vec(W) row-major:
    vec(W M_v − v W) = (I_d ⊗ M_vᵀ − v ⊗ I_K) w =: L_v w, so H = Σ_v L_vᵀ (I_d ⊗ G) L_v. G = ΘᵀΘ of the (whole,
    all-reduced) data set makes H a constant of the fit: sb_fit_step evaluates the regulariser and its gradient 2Hw
    in the fused kernel's epilogue with no further data pass."""
    d, K = lib.dim, lib.K
    dev = G.device
    Gd = G.to(torch.float64)
    eye_d = torch.eye(d, dtype=torch.float64, device=dev)
    eye_K = torch.eye(K, dtype=torch.float64, device=dev)
    IG = torch.kron(eye_d, Gd)
    H = torch.zeros(d * K, d * K, dtype=torch.float64, device=dev)
    for v in gens:
        v = torch.as_tensor(v).to(dev, torch.float64)
        M = lie_matrix(lib, v).to(dev)
        L = torch.kron(eye_d, M.T.contiguous()) - torch.kron(v, eye_K)
        H = H + L.T @ IG @ L
    return 0.5 * (H + H.T)


def lie_loss_per_sample(z: torch.Tensor, W: torch.Tensor, gens: Sequence[torch.Tensor], lib: Library) -> torch.Tensor:
    """Σ_v ‖J_h(z)(v z) − v h(z)‖² with the CUDA forward / JVP operators (any library)."""
    h = ops.sindy_forward(z, W, lib)
    loss = h.new_zeros(())
    for v in gens:
        v = v.to(z.device, z.dtype)
        vz = torch.einsum('ij,...j->...i', v, z)
        jv = ops.sindy_jvp(z, vz, W, lib)
        loss = loss + ((jv - torch.einsum('ij,...j->...i', v, h)) ** 2).sum()
    return loss
