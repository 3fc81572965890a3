"""CPU oracle for the SINDy hot path — TEST INFRASTRUCTURE ONLY.

A plain NumPy / CPU-PyTorch restatement of the reference's algorithms for this path. Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import it; the product
(`symmetry-ode-discovery_b200/`) never does.

Pinning: the reference has NO tests, so its own golden vectors do not exist. The oracle is
pinned instead against outputs of the unmodified reference imported from /root/reference in the build
container (`oracle/gen_golden.py` -> `tests/golden/*.npz`, checked by `tests/test_oracle_golden.py`) and
against the truth tables / known answers of SURVEY.md §8c. Degrees 4-5 of the polynomial library are an
extension the reference does not have: for those the oracle is the definition ("parity unpinned" for
poly_order > 3).

Every function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

import itertools
import math

import numpy as np
import torch

# ---------------------------------------------------------------------------------------------------------
# library  (sindy.py:7-30, 68-77, 179-189, 201-203)
# ---------------------------------------------------------------------------------------------------------


def poly_tuples(d: int, p: int):
    """Index tuples of the polynomial columns in the reference's order (nested i<=j<=k loops)."""
    cols = [()]
    for n in range(1, p + 1):
        cols.extend(itertools.combinations_with_replacement(range(d), n))
    return cols


def term_count(d: int, p: int, sine: bool = False, exp: bool = False) -> int:
    """sindy.py:179-189 get_term_num, continued past degree 3."""
    return sum(math.comb(d + n - 1, n) for n in range(p + 1)) + d * (int(bool(sine)) + int(bool(exp)))


def exponents(d: int, p: int) -> np.ndarray:
    E = np.zeros((len(poly_tuples(d, p)), d), dtype=np.int64)
    for k, c in enumerate(poly_tuples(d, p)):
        for j in c:
            E[k, j] += 1
    return E


def theta(x: np.ndarray, p: int, sine: bool = False, exp: bool = False) -> np.ndarray:
    """Θ(x) in the dtype of x; products formed left to right like `x[..., i] * x[..., j] * x[..., k]`."""
    x = np.asarray(x)
    d = x.shape[-1]
    cols = []
    for c in poly_tuples(d, p):
        if len(c) == 0:
            cols.append(np.ones(x.shape[:-1], dtype=x.dtype))
            continue
        v = x[..., c[0]]
        for j in c[1:]:
            v = v * x[..., j]
        cols.append(v)
    if sine:
        cols += [np.sin(x[..., j]) for j in range(d)]
    if exp:
        cols += [np.exp(x[..., j]) for j in range(d)]
    return np.stack(cols, axis=-1)


def dtheta(x: np.ndarray, p: int, sine: bool = False, exp: bool = False) -> np.ndarray:
    """Jacobian of Θ: (..., K, d), float64, from the exponent table (d x^a / dx_j = a_j x^(a-e_j))."""
    x = np.asarray(x, dtype=np.float64)
    d = x.shape[-1]
    E = exponents(d, p)
    J = np.zeros(x.shape[:-1] + (term_count(d, p, sine, exp), d))
    for k in range(E.shape[0]):
        for j in range(d):
            if E[k, j] == 0:
                continue
            e = E[k].copy()
            e[j] -= 1
            J[..., k, j] = E[k, j] * np.prod(x ** e, axis=-1)
    k = E.shape[0]
    if sine:
        for j in range(d):
            J[..., k, j] = np.cos(x[..., j]); k += 1
    if exp:
        for j in range(d):
            J[..., k, j] = np.exp(x[..., j]); k += 1
    return J


def d2theta_uv(x, u, p, sine=False, exp=False):
    """sum_{j,l} d²Θ_k/dx_j dx_l · u_j (returned as (..., K, d) over l), float64."""
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    d = x.shape[-1]
    E = exponents(d, p)
    H = np.zeros(x.shape[:-1] + (term_count(d, p, sine, exp), d))
    for k in range(E.shape[0]):
        for j in range(d):
            for l in range(d):
                e = E[k].copy()
                c = e[j]
                if c == 0:
                    continue
                e[j] -= 1
                c2 = e[l]
                if c2 == 0:
                    continue
                e[l] -= 1
                H[..., k, l] += c * c2 * np.prod(x ** e, axis=-1) * u[..., j]
    k = E.shape[0]
    if sine:
        for j in range(d):
            H[..., k, j] = -np.sin(x[..., j]) * u[..., j]; k += 1
    if exp:
        for j in range(d):
            H[..., k, j] = np.exp(x[..., j]) * u[..., j]; k += 1
    return H


# ---------------------------------------------------------------------------------------------------------
# model and losses  (sindy.py:79-82; train.py:645-690)
# ---------------------------------------------------------------------------------------------------------


def forward(x, W, p, sine=False, exp=False):
    """h(x) = Θ(x)·Wᵀ with Θ in fp32 (as the reference) and the contraction in fp64."""
    th = theta(np.asarray(x, dtype=np.float32), p, sine, exp).astype(np.float64)
    return th @ np.asarray(W, dtype=np.float64).T


def jvp(x, u, W, p, sine=False, exp=False):
    """J_h(x)·u — what `jvp(regressor, x, u)[1]` returns (model_utils.py:56, train.py:505)."""
    J = dtheta(x, p, sine, exp)                                  # (..., K, d)
    t = np.einsum('...kj,...j->...k', J, np.asarray(u, dtype=np.float64))
    return t @ np.asarray(W, dtype=np.float64).T


def backward(x, gy, W, p, sine=False, exp=False):
    """(gW, gx) for cotangent gy of forward: gW = gyᵀΘ, gx = J_hᵀ gy."""
    th = theta(np.asarray(x, dtype=np.float32), p, sine, exp).astype(np.float64)
    gy = np.asarray(gy, dtype=np.float64)
    gW = gy.reshape(-1, gy.shape[-1]).T @ th.reshape(-1, th.shape[-1])
    c = gy @ np.asarray(W, dtype=np.float64)                     # (..., K)
    gx = np.einsum('...k,...kj->...j', c, dtheta(x, p, sine, exp))
    return gW, gx


def jvp_backward(x, u, g, W, p, sine=False, exp=False):
    """Cotangents (gW, gx, gu) of jvp() for cotangent g."""
    J = dtheta(x, p, sine, exp)
    u = np.asarray(u, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    t = np.einsum('...kj,...j->...k', J, u)
    gW = g.reshape(-1, g.shape[-1]).T @ t.reshape(-1, t.shape[-1])
    c = g @ np.asarray(W, dtype=np.float64)
    gu = np.einsum('...k,...kj->...j', c, J)
    gx = np.einsum('...k,...kl->...l', c, d2theta_uv(x, u, p, sine, exp))
    return gW, gx, gu


def train_step_sums(x, dx, W, p, sine=False, exp=False):
    """The packed sums of sb_train_step: Σr², Σ r_i Θ_k, ΘᵀΘ, ΘᵀẊ (fp64 from fp32 Θ)."""
    th = theta(np.asarray(x, dtype=np.float32), p, sine, exp).astype(np.float64)
    th = th.reshape(-1, th.shape[-1])
    dx = np.asarray(dx, dtype=np.float64).reshape(-1, np.shape(dx)[-1])
    r = th @ np.asarray(W, dtype=np.float64).T - dx
    return {"sum_sq": float((r ** 2).sum()), "n": th.shape[0], "grad_raw": r.T @ th, "gram": th.T @ th,
            "b": th.T @ dx}


def mse_loss_and_grad(x, dx, W, p, sine=False, exp=False):
    """MSELoss(regressor(x), dx) and d/dW (train.py:663-664, 689): mean over N·d elements."""
    s = train_step_sums(x, dx, W, p, sine, exp)
    nd = s["n"] * np.shape(dx)[-1]
    return s["sum_sq"] / nd, 2.0 * s["grad_raw"] / nd


def torch_theta(x: torch.Tensor, p: int, sine=False, exp=False) -> torch.Tensor:
    """The reference's cat-of-products idiom (sindy.py:7-30, 81) for any degree; autograd-capable."""
    d = x.shape[-1]
    cols = []
    for c in poly_tuples(d, p):
        if len(c) == 0:
            cols.append(torch.ones(*x.shape[:-1], 1, device=x.device))
            continue
        v = x[..., c[0]]
        for j in c[1:]:
            v = v * x[..., j]
        cols.append(v.view(*x.shape[:-1], 1))
    if sine:
        cols.append(torch.sin(x))
    if exp:
        cols.append(torch.exp(x))
    return torch.cat(cols, dim=-1)


class TorchPolyN(torch.nn.Module):
    """Degree-n block of the library as an nn.Module in the idiom of the reference's SINDyPoly2 / SINDyPoly3
    (sindy.py:13-24: nested non-decreasing index loops, products formed left to right, one column each, torch.cat).
    The reference stops at n = 3; appending TorchPolyN(4) and TorchPolyN(5) to `SINDyRegression.terms` extends ITS
    forward pass to the degree-5 library of BASELINE config 5 without touching its code (bench.py --impl reference)."""

    def __init__(self, n: int):
        super().__init__()
        self.n = n

    def forward(self, x):
        d = x.shape[-1]
        cols = []
        for c in itertools.combinations_with_replacement(range(d), self.n):
            v = x[..., c[0]]
            for j in c[1:]:
                v = v * x[..., j]
            cols.append(v.unsqueeze(-1))
        return torch.cat(cols, dim=-1)


class TorchRegressor(torch.nn.Module):
    """Port of SINDyRegression.forward (sindy.py:79-82) for any degree: used by the benchmark's CPU arm only when no
    reference checkout (baseline/_ref) is available."""

    def __init__(self, d, p):
        super().__init__()
        self.p = p
        self.Xi = torch.nn.Parameter(torch.randn(d, term_count(d, p)))
        self.mask = torch.ones_like(self.Xi)

    def forward(self, x):
        return torch_theta(x, self.p) @ (self.Xi * self.mask).T


def torch_closure(x, dx, Xi, mask, p, sine=False, exp=False, w_sindy_x=1.0, w_sindy_reg=0.0):
    """One evaluation of the LBFGS closure of train.py:645-690 (no sym-reg): returns (loss, Xi.grad).
    This is the CPU baseline the benchmark times ("port" of the reference closure for any degree)."""
    Xi = Xi.detach().clone().requires_grad_(True)
    pred = torch_theta(x, p, sine, exp) @ (Xi * mask).T
    loss = w_sindy_x * torch.nn.functional.mse_loss(pred, dx)
    loss = loss + w_sindy_reg * torch.norm(Xi, 1)
    loss.backward()
    return loss.detach(), Xi.grad


def torch_adam_loop(x, dx, Xi, mask, p, n_steps, lr, w_sindy_x=1.0, w_sindy_reg=0.0, optimizer="adam", sine=False,
                    exp=False):
    """`n_steps` iterations of the Adam loop of train.py:512-530 without sym-reg on one fixed batch
    (forward, MSELoss, L1 of the unmasked parameters, backward, optimizer.step()) with torch's own optimiser on CPU.
    Returns (final Xi, [loss at every step before its update], last gradient)."""
    Xi = torch.nn.Parameter(Xi.detach().clone())
    opt = torch.optim.Adam([Xi], lr=lr) if optimizer == "adam" else torch.optim.SGD([Xi], lr=lr)
    losses = []
    for _ in range(n_steps):
        pred = torch_theta(x, p, sine, exp) @ (Xi * mask).T
        loss = w_sindy_x * torch.nn.functional.mse_loss(pred, dx) + w_sindy_reg * torch.norm(Xi, 1)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return Xi.detach().clone(), losses, Xi.grad.detach().clone()


# ---------------------------------------------------------------------------------------------------------
# linear Lie-derivative regulariser  (train.py:503-507, intended formula `jvp(...)[1]`)
# ---------------------------------------------------------------------------------------------------------


def lie_reg_linear(z, W, gens, p, sine=False, exp=False):
    """sum_v sum_n ||J_h(z_n)(v z_n) − v h(z_n)||² (sum over the batch, not mean)."""
    z = np.asarray(z, dtype=np.float64)
    h = forward(z, W, p, sine, exp)
    total = 0.0
    for v in gens:
        v = np.asarray(v, dtype=np.float64)
        vz = z @ v.T
        total += float(((jvp(z, vz, W, p, sine, exp) - h @ v.T) ** 2).sum())
    return total


def lie_matrix(d, p, L):
    """M with J_Θ(z)·L·z = M·Θ(z) (sindy.py:123-144 computes it symbolically)."""
    cols = poly_tuples(d, p)
    where = {c: k for k, c in enumerate(cols)}
    L = np.asarray(L, dtype=np.float64)
    M = np.zeros((len(cols), len(cols)))
    for k, c in enumerate(cols):
        for pos, j in enumerate(c):
            rest = c[:pos] + c[pos + 1:]
            for l in range(d):
                M[k, where[tuple(sorted(rest + (l,)))]] += L[j, l]
    return M


# ---------------------------------------------------------------------------------------------------------
# reversed regulariser with precomputed g(x), J_g(x)  (model_utils.py:126-170, 172-211)
# ---------------------------------------------------------------------------------------------------------


def symmreg_r_precomputed(x, gx_list, Jgx_list, W, p, sine=False, exp=False):
    loss = 0.0
    for gx, Jg in zip(gx_list, Jgx_list):
        d = np.shape(x)[-1]
        Jg = np.asarray(Jg, dtype=np.float64).reshape(-1, d, d)  # the reference returns (B, 1, d, d)
        pushed = np.einsum('bij,bj->bi', Jg, forward(x, W, p, sine, exp))
        loss += float(np.mean((pushed - forward(gx, W, p, sine, exp)) ** 2))
    return loss


# ---------------------------------------------------------------------------------------------------------
# STLSQ  (sindy.py:192-195, 250-324) — restated with torch.linalg.lstsq on the CPU like the reference
# ---------------------------------------------------------------------------------------------------------


def set_threshold(Xi, mask, thr):
    """sindy.py:192-195: strict `>` and AND with the previous mask."""
    return np.logical_and(np.abs(Xi) > thr, mask > 0).astype(np.float32)


def stlsq_one_step(x, y, mask, w_reg, thr, p, sine=False, exp=False):
    """Unconstrained solve_SINDy_one_step (sindy.py:250-315): returns (Xi, new_mask, converged)."""
    th = torch.from_numpy(theta(np.asarray(x, dtype=np.float32), p, sine, exp))
    yt = torch.from_numpy(np.asarray(y, dtype=np.float32))
    K, d = th.shape[1], yt.shape[1]
    A = torch.cat([th, w_reg * torch.eye(K)], dim=0)
    B = torch.cat([yt, torch.zeros(K, d)], dim=0)
    m = torch.from_numpy(np.asarray(mask)) > 0
    if not bool(torch.all(m)):
        A = torch.block_diag(*([A] * d))[:, m.flatten()]
        B = B.T.reshape(-1)
        sol = torch.linalg.lstsq(A, B).solution
        Xi = torch.zeros(d, K)
        Xi[m] = sol
    else:
        Xi = torch.linalg.lstsq(A, B).solution.T
    Xi = Xi.numpy()
    new_mask = set_threshold(Xi, np.asarray(mask), thr)
    return Xi, new_mask, bool(np.allclose(new_mask, mask))


def stlsq(x, y, w_reg, thr, p, sine=False, exp=False, max_iter=5):
    """solve_SINDy (sindy.py:318-324): mask reset, at most max_iter steps."""
    d = np.shape(y)[-1]
    mask = np.ones((d, term_count(np.shape(x)[-1], p, sine, exp)), dtype=np.float32)
    Xi = None
    for _ in range(max_iter):
        Xi, mask, conv = stlsq_one_step(x, y, mask, w_reg, thr, p, sine, exp)
        if conv:
            break
    return Xi, mask


# ---------------------------------------------------------------------------------------------------------
# WSINDy  (sindy.py:332-395)
# ---------------------------------------------------------------------------------------------------------


def wsindy_test_functions(n_steps, dt, t_max, n_test=50):
    """V, V' (n_test × T) in fp32 with the reference's operation order (sindy.py:337-347)."""
    t = torch.arange(n_steps) * dt
    dt_t = t[1] - t[0]
    k = torch.arange(1, n_test + 1, dtype=torch.float32).view(-1, 1)
    g = math.sqrt(2 / t_max) * torch.sin(k * torch.pi * t / t_max)
    gd = math.sqrt(2 / t_max) * k * np.pi / t_max * torch.cos(k * np.pi * t / t_max)
    return (dt_t * g).numpy(), (dt_t * gd).numpy()


def wsindy_integrals(x, dt, t_max, p, sine=False, exp=False, n_test=50):
    """G = VΘ(x), b = −V'x (sindy.py:361-362), fp64 contraction of the fp32 factors."""
    x = np.asarray(x, dtype=np.float32)
    V, Vd = wsindy_test_functions(x.shape[0], dt, t_max, n_test)
    th = theta(x, p, sine, exp).astype(np.float64)
    return V.astype(np.float64) @ th, -(Vd.astype(np.float64) @ x.astype(np.float64))


def wsindy_one_step(x, mask, w_reg, thr, dt, t_max, p, sine=False, exp=False, n_test=50, solve_dtype=torch.float32):
    """WSINDyWrapper.solve (sindy.py:352-395) with torch.linalg.lstsq in fp32 like the reference.
    solve_dtype=torch.float64: the same algorithm with the fp32-built test functions and data promoted to float64
    before the products and the least-squares solve — the answer the reference's fp32 LAPACK call is a noisy estimate
    of. Needed where the fp32 solve is not reproducible (Sel'kov, w = 0: sigma_min/sigma_max = 1.6e-4; the reference's
    own result for BASELINE config 4 changes completely with the MKL thread count)."""
    x32 = np.asarray(x, dtype=np.float32)
    V, Vd = wsindy_test_functions(x32.shape[0], dt, t_max, n_test)
    V, Vd = torch.from_numpy(V).to(solve_dtype), torch.from_numpy(Vd).to(solve_dtype)
    xt = torch.from_numpy(x32).to(solve_dtype)
    G = V @ torch.from_numpy(theta(x32, p, sine, exp)).to(solve_dtype)
    b = -Vd @ xt
    K, d = G.shape[1], xt.shape[1]
    G_aug = torch.cat([V.T @ G, math.sqrt(w_reg) * torch.eye(K, dtype=solve_dtype)], dim=0)
    b_aug = torch.cat([V.T @ b, torch.zeros(K, d, dtype=solve_dtype)], dim=0)
    m = torch.from_numpy(np.asarray(mask)) > 0
    if not bool(torch.all(m)):
        G_aug = torch.block_diag(*([G_aug] * d))[:, m.flatten()]
        b_aug = b_aug.T.reshape(-1)
        sol = torch.linalg.lstsq(G_aug, b_aug).solution
        Xi = torch.zeros(d, K, dtype=solve_dtype)
        Xi[m] = sol
    else:
        Xi = torch.linalg.lstsq(G_aug, b_aug).solution.T
    Xi = Xi.to(torch.float32).numpy()
    new_mask = set_threshold(Xi, np.asarray(mask), thr)
    return Xi, new_mask, bool(np.allclose(new_mask, mask))


# ---------------------------------------------------------------------------------------------------------
# integrators  (data_utils/ode.py:7-28; model_utils.py:223-255)
# ---------------------------------------------------------------------------------------------------------


def library_rhs(Xi, p, sine=False, exp=False):
    Xi = np.asarray(Xi)

    def f(x):
        return theta(x, p, sine, exp) @ Xi.T.astype(x.dtype)

    return f


def solve_ode_batch(f, x0, dt=0.002, num_steps=2000):
    """NumPy float64 RK4 with dx recorded at every row (data_utils/ode.py:7-28)."""
    x0 = np.asarray(x0, dtype=np.float64)
    x = np.zeros((num_steps,) + x0.shape)
    dx = np.zeros_like(x)
    x[0] = x0
    for i in range(num_steps):
        d1 = f(x[i])
        dx[i] = d1
        if i == num_steps - 1:
            break
        k1 = dt * d1
        k2 = dt * f(x[i] + 0.5 * k1)
        k3 = dt * f(x[i] + 0.5 * k2)
        k4 = dt * f(x[i] + k3)
        x[i + 1] = x[i] + (k1 + 2 * k2 + 2 * k3 + k4) / 6
    return x, dx


def odeint(f, x0, t, dt, method='euler', full_traj=False):
    """model_utils.py:223-255 on NumPy arrays in the dtype of x0 (fp32 for parity with torch)."""
    n_steps = int(t / dt)
    x = np.asarray(x0)
    one = x.dtype.type
    traj = []
    for _ in range(n_steps):
        if method == 'euler':
            x = x + one(dt) * f(x)
        elif method == 'rk4':
            k1 = f(x)
            k2 = f(x + one(dt / 2) * k1)
            k3 = f(x + one(dt / 2) * k2)
            k4 = f(x + one(dt) * k3)
            x = x + one(dt / 6) * (k1 + one(2) * k2 + one(2) * k3 + k4)
        else:
            raise ValueError('Unrecognized ODEInt method.')
        if full_traj:
            traj.append(x)
    return np.stack(traj, axis=0) if full_traj else x


# ---------------------------------------------------------------------------------------------------------
# GP smoother of the data generators  (data_utils/smoothing.py:155-196 num_diff_gp, GPPCA0 :17-152)
# ---------------------------------------------------------------------------------------------------------


def num_diff_gp(x, dt, noise_level, std_base, sigma_in=None, eps=0.001):
    """x: (T, n_traj, dim) noisy states. Returns (dX, X) like the reference.

    The reference keeps r = n_traj principal components (`smoothing.py:181-182`), i.e. ALL of them, so the factor
    loading A is a full orthogonal matrix and X_hat = K_new (K + s^2 I)^-1 Y A A^T (`:137-143`) is plain GP regression
    with the RBF kernel K = s_out^2 exp(-(t_i - t_j)^2 / (2 s_in^2)), s_out = std_base[d], s = noise_level * std_base[d],
    s_in = sigma_in or dt (`:29-32`); the derivative is the forward difference of the posterior mean over eps = 0.001
    (`:186-195`)."""
    x = np.asarray(x, dtype=np.float64)
    T = x.shape[0]
    t = np.arange(T) * dt
    dX, Xs = np.empty_like(x), np.empty_like(x)
    for d in range(x.shape[2]):
        s_out, s = std_base[d], noise_level * std_base[d]
        s_in = (t[1] - t[0]) if sigma_in is None else sigma_in
        K = s_out ** 2 * np.exp(-(t[:, None] - t[None, :]) ** 2 / (2 * s_in ** 2))
        K2 = s_out ** 2 * np.exp(-((t + eps)[:, None] - t[None, :]) ** 2 / (2 * s_in ** 2))
        alpha = np.linalg.solve(K + s ** 2 * np.eye(T), x[:, :, d])
        Xs[:, :, d] = K @ alpha
        dX[:, :, d] = (K2 @ alpha - Xs[:, :, d]) / eps
    return dX, Xs
