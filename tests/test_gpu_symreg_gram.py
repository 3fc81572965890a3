"""GPU parity for the moment-based Gram and the linear Lie-derivative regulariser."""
import numpy as np
import pytest
import torch

from oracle import sindy_oracle as O

pytestmark = pytest.mark.gpu


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype).cuda()


def rel(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("d,p", [(2, 2), (2, 3), (3, 2), (3, 3), (3, 5)])
@pytest.mark.parametrize("n", [2, 7, 5001, 400_003])
def test_moment_gram_vs_oracle(d, p, n, monkeypatch):
    from sindy_b200 import native, symreg
    rng = np.random.default_rng(n + 10 * d + p)
    lib = native.Library(d, p)
    assert native.train_step_variant(lib, native.SB_STEP_GRAM) == "moments"
    monkeypatch.setenv("SB_MOMENTS_MIN_SAMPLES", "0")     # force the power-sum kernel at test sizes
    x = rng.uniform(-1.2, 1.2, (n, d)).astype(np.float32)
    th = O.theta(x, p).astype(np.float64)
    G = symreg.gram(dev(x), lib)
    assert rel(G, th.T @ th) < 2e-5
    assert torch.equal(G, G.T)                                   # both triangles read the same power sum
    assert torch.equal(G, symreg.gram(dev(x), lib))              # deterministic
    # all sections at once: header, residual sums, Gram, ΘᵀẊ from the specialised kernels
    dx = rng.standard_normal((n, d)).astype(np.float32)
    W = rng.standard_normal((d, lib.K)).astype(np.float32)
    parts = native.unpack_step(native.train_step(dev(x), dev(dx), dev(W), lib, 15), lib, 15)
    ref = O.train_step_sums(x, dx, W, p)
    assert float(parts["n"]) == n and abs(float(parts["sum_sq"]) - ref["sum_sq"]) < 2e-5 * ref["sum_sq"]
    for key in ("grad_raw", "gram", "b"):
        assert rel(parts[key], ref[key]) < 2e-5, key


def test_lie_regulariser_matches_reference_golden(golden):
    import sindy
    g = golden("jvp")
    for (d, p, s, e) in [(2, 2, 0, 1), (2, 3, 1, 0), (3, 3, 0, 0), (3, 2, 1, 1)]:
        t = f"d{d}p{p}s{s}e{e}"
        reg = sindy.SINDyRegression(d, p, bool(s), bool(e), threshold=0.05, device="cuda", constrain_constant=True)
        reg.Xi.data = dev(g[t + "_Xi"])
        x, gens = dev(g[t + "_x"]), list(dev(g[t + "_gens"]))
        methods = ["jvp"] + (["gram"] if not (s or e) else [])
        for m in methods:
            reg.zero_grad()
            loss = reg.lie_reg_loss(x, gens, method=m)
            loss.backward()
            assert abs(float(loss) - float(g[t + "_lie"])) < 2e-5 * float(g[t + "_lie"]), (t, m)
            assert rel(reg.Xi.grad, g[t + "_grad_lie"]) < 1e-4, (t, m)


def test_lie_matrix_is_the_symbolic_map(golden):
    from sindy_b200 import native, symreg
    g = golden("stlsq")
    so2 = np.array([[0.0, 1.0], [-1.0, 0.0]])
    np.testing.assert_allclose(symreg.lie_matrix(native.Library(2, 2), so2).numpy(), g["so2_M"], atol=1e-6)
    np.testing.assert_allclose(symreg.lie_matrix(native.Library(2, 2), np.diag([2.0, 1.0])).numpy(), g["scaling2_M"], atol=1e-6)


def test_sharded_step_with_symreg_matches_oracle(monkeypatch):
    from sindy_b200 import native
    monkeypatch.setenv("SB_MOMENTS_MIN_SAMPLES", "0")
    from sindy_b200.dist import ShardedTrainStep
    rng = np.random.default_rng(2)
    d, p = 3, 5
    lib = native.Library(d, p)
    n = 30_011
    x = rng.uniform(-1, 1, (n, d)).astype(np.float32)
    dx = rng.standard_normal((n, d)).astype(np.float32)
    Xi = (0.3 * rng.standard_normal((d, lib.K))).astype(np.float32)
    mask = (rng.random((d, lib.K)) > 0.2).astype(np.float32)
    so3 = np.zeros((3, 3, 3))
    k = 0
    for i in range(3):
        for j in range(i):
            so3[k, i, j], so3[k, j, i] = 1, -1
            k += 1
    w_sym = 0.1
    step = ShardedTrainStep(lib, dev(x), dev(dx), sym_gens=list(so3), w_sym=w_sym)
    loss, grad = step.step(dev(Xi), dev(mask), 0.0)
    W = Xi * mask
    l_mse, g_mse = O.mse_loss_and_grad(x, dx, W, p)
    l_sym = O.lie_reg_linear(x, W, so3, p)
    assert abs(float(loss) - (l_mse + w_sym * l_sym)) < 3e-5 * (l_mse + w_sym * l_sym)
    # gradient of the regulariser by central differences of the oracle on a few entries
    eps = 1e-4
    for (i, kk) in [(0, 1), (1, 7), (2, 30), (0, 55)]:
        Wp, Wm = W.astype(np.float64).copy(), W.astype(np.float64).copy()
        Wp[i, kk] += eps; Wm[i, kk] -= eps
        fd = (O.lie_reg_linear(x, Wp, so3, p) - O.lie_reg_linear(x, Wm, so3, p)) / (2 * eps)
        want = g_mse[i, kk] * mask[i, kk] + w_sym * fd * mask[i, kk]
        assert abs(float(grad[i, kk]) - want) < 2e-4 * max(1.0, abs(want)), (i, kk, float(grad[i, kk]), want)
    # graph replay gives the same numbers
    gstep = ShardedTrainStep(lib, dev(x), dev(dx), sym_gens=list(so3), w_sym=w_sym, use_graph=True)
    l2, g2 = gstep.step(dev(Xi), dev(mask), 0.0)
    assert abs(float(l2) - float(loss)) < 1e-6 * abs(float(loss)) and rel(g2, grad) < 1e-6
