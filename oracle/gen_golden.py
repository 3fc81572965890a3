"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the reference does not exist on the GPU box):

    python oracle/gen_golden.py

Every fixture stores seeded inputs and the reference's outputs for one slice of the hot path
(SURVEY.md §8a rows a1-a13). The fixtures pin both the CPU oracle (tests/test_oracle_golden.py) and the CUDA
path (tests/test_gpu_parity.py).
"""
import os
import sys

REF = os.environ.get("SINDY_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
os.environ.setdefault("WANDB_MODE", "disabled")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.autograd.functional import jvp  # noqa: E402

import sindy as ref_sindy  # noqa: E402  (the reference's)
import model_utils as ref_mu  # noqa: E402
from data_utils import ode as ref_ode  # noqa: E402
from data_utils import damped_oscillator, growth, lotka, selkov  # noqa: E402
from evaluation.eval_eq import sindy_truth  # noqa: E402

import standins  # noqa: E402  (tests/standins.py: seeded frozen autoencoder / generator stand-ins)

assert ref_sindy.__file__.startswith(REF), ref_sindy.__file__
torch.set_num_threads(4)


def make_reg(d, p, sine, exp, L_list=(), threshold=0.05, constrain_constant=True, seed=0):
    torch.manual_seed(seed)
    return ref_sindy.SINDyRegression(d, p, sine, exp, L_list=list(L_list), threshold=threshold, device="cpu",
                                     constrain_constant=constrain_constant)


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


LIBS = [(2, 2, 0, 0), (2, 2, 0, 1), (2, 3, 0, 0), (2, 3, 1, 0), (3, 3, 1, 1), (3, 2, 0, 0), (1, 3, 1, 1), (4, 3, 0, 0)]


def gen_model():
    """a1-a4: Θ, forward, closure loss (MSE + L1) and gradient; known answers."""
    out = {}
    for (d, p, s, e) in LIBS:
        tag = f"d{d}p{p}s{s}e{e}"
        reg = make_reg(d, p, bool(s), bool(e), seed=d * 10 + p)
        g = torch.Generator().manual_seed(100 + d + p)
        n = 257
        x = torch.rand(n, d, generator=g) * 2 - 1
        dx = torch.randn(n, d, generator=g)
        mask = (torch.rand(d, reg.Xi.shape[1], generator=g) > 0.3).float()
        reg.mask.data = mask
        th = reg.eval_Theta_at(x)
        y = reg(x)
        loss_x = torch.nn.MSELoss()(y, dx)
        l1 = sum(torch.norm(q, 1) for q in reg.parameters())
        (loss_x + 0.01 * l1).backward()
        out.update({f"{tag}_x": x, f"{tag}_dx": dx, f"{tag}_Xi": reg.Xi.detach(), f"{tag}_mask": mask,
                    f"{tag}_theta": th, f"{tag}_y": y.detach(), f"{tag}_loss_x": loss_x.detach(),
                    f"{tag}_l1": l1.detach(), f"{tag}_grad": reg.Xi.grad})
        # 3-D leading shape (B, n_comps, d) as the autoencoder path feeds it
        x3 = x[:64].reshape(32, 2, d)
        out[f"{tag}_y3"] = reg(x3).detach()
    # SURVEY §8c known answers
    reg = make_reg(3, 3, False, False)
    out["ka_theta_235"] = reg.eval_Theta_at(torch.tensor([[2.0, 3.0, 5.0]]))[0]
    reg = make_reg(2, 1, False, False)
    reg.Xi.data = torch.tensor([[0.05, 0.0500001, -0.05], [1.0, -1.0, 0.0]])
    reg.set_threshold(0.05)
    out["ka_threshold_mask"] = reg.mask
    out["ka_int_steps"] = np.array([int(0.1 / 0.01), int(0.3 / 0.1), int(3.0 / 0.1), int(2.0 / 0.02)])
    for name, tab in sindy_truth.items():
        out[f"truth_{name}"] = tab
    save("model", **out)


def gen_jvp():
    """a5/a6 building blocks: jvp(regressor, x, u)[1], the linear Lie-derivative loss (intended formula of
    train.py:503-507) and its gradient, and the gradient THROUGH a jvp (double-vjp with create_graph)."""
    out = {}
    for (d, p, s, e) in [(2, 2, 0, 1), (2, 3, 1, 0), (3, 3, 0, 0), (3, 2, 1, 1)]:
        tag = f"d{d}p{p}s{s}e{e}"
        reg = make_reg(d, p, bool(s), bool(e), seed=7 + d)
        g = torch.Generator().manual_seed(5 + d * p)
        n = 129
        x = torch.rand(n, d, generator=g) * 2 - 1
        u = torch.randn(n, d, generator=g)
        jv = jvp(reg, x, u)[1]
        # loss through the jvp, differentiated w.r.t. Xi, x-path (Hessian) exercised through an Euler step
        def two_steps(z):
            z1 = z + 0.1 * reg(z)
            return z1 + 0.1 * reg(z1)
        jv2 = jvp(two_steps, x, u, create_graph=True, strict=True)[1]
        tgt = torch.randn(n, d, generator=g)
        loss = torch.mean((jv2 - tgt) ** 2) / torch.mean(jv2 ** 2)
        reg.zero_grad()
        loss.backward()
        gens = [torch.randn(d, d, generator=g) for _ in range(2)]
        lie = 0.0
        for v in gens:
            lie = lie + torch.norm(jvp(reg, x, torch.einsum('ij,bj->bi', v, x), create_graph=True)[1]
                                   - torch.einsum('ij,bj->bi', v, reg(x))) ** 2
        grad_two = reg.Xi.grad.clone()
        reg.zero_grad()
        lie.backward()
        out.update({f"{tag}_x": x, f"{tag}_u": u, f"{tag}_Xi": reg.Xi.detach(), f"{tag}_jv": jv.detach(),
                    f"{tag}_tgt": tgt, f"{tag}_loss_two": loss.detach(), f"{tag}_grad_two": grad_two,
                    f"{tag}_gens": torch.stack(gens), f"{tag}_lie": lie.detach(), f"{tag}_grad_lie": reg.Xi.grad})
    save("jvp", **out)


def gen_stlsq():
    """a9/a13: solve_SINDy / solve_SINDy_one_step, unconstrained and constrained; Q and M known answers."""
    out = {}
    rng = np.random.default_rng(3)
    # selkov-like data, d=2 p=3
    n = 1500
    x = rng.uniform(0.2, 1.5, size=(n, 2)).astype(np.float32)
    truth = sindy_truth["selkov"]
    reg = make_reg(2, 3, False, False)
    th = reg.eval_Theta_at(torch.from_numpy(x)).numpy().astype(np.float64)
    y = (th @ truth.T + 0.01 * rng.standard_normal((n, 2))).astype(np.float32)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    for w in (0.0, 0.3):
        reg = make_reg(2, 3, False, False)
        reg.reset_mask()
        masks, xis = [], []
        for it in range(5):
            _, conv = ref_sindy.solve_SINDy_one_step(reg, xt, yt, w, 0.05)
            masks.append(reg.mask.clone()); xis.append(reg.Xi.detach().clone())
            if conv:
                break
        tag = f"selkov_w{int(w * 10)}"
        out.update({f"{tag}_masks": torch.stack(masks), f"{tag}_xis": torch.stack(xis)})
    out.update({"selkov_x": x, "selkov_y": y})
    # d=3 p=3 Lorenz-like (unit box), partial start mask
    n = 2000
    x3 = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    Xi3 = np.zeros((3, 20))
    Xi3[0, 1], Xi3[0, 2] = -10, 10
    Xi3[1, 1], Xi3[1, 2], Xi3[1, 6] = 2.8, -1, -1
    Xi3[2, 5], Xi3[2, 3] = 1, -8 / 3
    reg = make_reg(3, 3, False, False)
    th = reg.eval_Theta_at(torch.from_numpy(x3)).numpy().astype(np.float64)
    y3 = (th @ Xi3.T + 0.01 * rng.standard_normal((n, 3))).astype(np.float32)
    reg = make_reg(3, 3, False, False)
    ref_sindy.solve_SINDy(reg, torch.from_numpy(x3), torch.from_numpy(y3), 0.0, 0.1)
    out.update({"lorenz_x": x3, "lorenz_y": y3, "lorenz_Xi": reg.Xi.detach(), "lorenz_mask": reg.mask,
                "lorenz_truth": Xi3})
    # constraint bases: so2 and scaling2
    so2 = torch.tensor([[0.0, 1.0], [-1.0, 0.0]])
    sc2 = torch.diag(torch.tensor([2.0, 1.0]))
    for name, L in (("so2", so2), ("scaling2", sc2)):
        reg = make_reg(2, 2, False, False, L_list=[L])
        out.update({f"{name}_Q": reg.Q, f"{name}_M": reg.get_M_list()[0], f"{name}_kron": int(reg.use_kron_product),
                    f"{name}_beta0": reg.beta.detach(), f"{name}_const0": reg.const.detach(),
                    f"{name}_Xi0": reg.get_Xi().detach()})
    so3ish = torch.tensor([[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 0.0]])  # singular -> non-kron layout
    reg = make_reg(3, 2, False, False, L_list=[so3ish])
    out.update({"sing3_Q": reg.Q, "sing3_M": reg.get_M_list()[0], "sing3_kron": int(reg.use_kron_product)})
    # constrained STLSQ on dosc (so2) and growth (scaling2), with and without free constants
    xd = rng.uniform(-1.5, 1.5, size=(1000, 2)).astype(np.float32)
    yd = (make_reg(2, 2, False, False).eval_Theta_at(torch.from_numpy(xd)).numpy().astype(np.float64)
          @ sindy_truth["dosc"].T + 0.02 * rng.standard_normal((1000, 2))).astype(np.float32)
    xg = rng.uniform(0.2, 1.0, size=(1000, 2)).astype(np.float32)
    yg = (make_reg(2, 2, False, False).eval_Theta_at(torch.from_numpy(xg)).numpy().astype(np.float64)
          @ sindy_truth["growth"].T + 0.005 * rng.standard_normal((1000, 2))).astype(np.float32)
    out.update({"dosc_x": xd, "dosc_y": yd, "growth_x": xg, "growth_y": yg})
    for name, L, xx, yy in (("dosc_so2", so2, xd, yd), ("growth_sc2", sc2, xg, yg)):
        for cc in (True, False):
            reg = make_reg(2, 2, False, False, L_list=[L], constrain_constant=cc)
            res = []
            for it in range(3):
                _, conv = ref_sindy.solve_SINDy_one_step(reg, torch.from_numpy(xx), torch.from_numpy(yy), 0.0, 0.05)
                res.append((reg.get_Xi().detach().clone(), reg.mask.clone()))
                if conv:
                    break
            tag = f"{name}_cc{int(cc)}"
            out.update({f"{tag}_xis": torch.stack([r[0] for r in res]), f"{tag}_masks": torch.stack([r[1] for r in res]),
                        f"{tag}_beta": reg.beta.detach(), f"{tag}_const": reg.const.detach()})
    save("stlsq", **out)


def gen_wsindy():
    """a10: V, V', G, b and the solve sequence on a selkov trajectory (config 4 shape, shortened)."""
    out = {}
    np.random.seed(11)
    x0 = selkov.generate_random_ics(3)
    x, _ = ref_ode.solve_ode_batch(selkov.selkov, x0, dt=0.002, num_steps=4001)
    x = x + 0.01 * np.random.randn(*x.shape) * x.std(axis=(0, 1))
    traj = torch.from_numpy(x[:, 0]).float()[:4000]
    T, dt = traj.shape[0], 0.002
    t = torch.arange(T) * dt
    t_max = T * dt
    for w in (0.0, 0.05, 0.01):
        reg = make_reg(2, 3, False, False, threshold=0.075)
        wr = ref_sindy.WSINDyWrapper(reg, t, t_max, device="cpu")
        if w == 0.0:
            G = wr.V @ reg.eval_Theta_at(traj)
            b = -wr.V_drv @ traj
            out.update({"V": wr.V[:, ::40], "V_drv": wr.V_drv[:, ::40], "G": G, "b": b})
        masks, xis = [], []
        for it in range(6):
            _, conv = wr.solve(traj, w, 0.075)
            masks.append(reg.mask.clone()); xis.append(reg.Xi.detach().clone())
            if conv:
                break
        out.update({f"w{int(w * 100)}_masks": torch.stack(masks), f"w{int(w * 100)}_xis": torch.stack(xis)})
    # a batch of trajectories for the batched integrals
    out.update({"traj": traj, "dt": dt, "t_max": t_max, "trajs3": torch.from_numpy(x[:4000]).float().permute(1, 0, 2)})
    save("wsindy", **out)


def gen_rollout():
    """a11/a12: solve_ode_batch (float64) for the four shipped systems, and torch fp32 odeint."""
    out = {}
    np.random.seed(5)
    systems = {
        "dosc": (damped_oscillator.dosc, damped_oscillator.generate_random_ics, 0.002, 2, False),
        "growth": (growth.growth, growth.generate_random_ics, 0.002, 2, False),
        "lv": (lotka.lotka_volterra, lotka.generate_random_ics, 0.002, 2, True),
        "selkov": (selkov.selkov, selkov.generate_random_ics, 0.002, 3, False),
    }
    for name, (f, ics, dt, p, exp) in systems.items():
        x0 = ics(7)
        x, dx = ref_ode.solve_ode_batch(f, x0, dt=dt, num_steps=301)
        out.update({f"{name}_x0": x0, f"{name}_x": x[::10], f"{name}_dx": dx[::10]})
    # long rollout, longest config: 10^4 steps (dosc), stored every 500th
    x0 = np.array([[1.0, 0.0], [0.3, -1.2]])
    x, dx = ref_ode.solve_ode_batch(damped_oscillator.dosc, x0, dt=0.002, num_steps=10000)
    out.update({"dosc_long_x0": x0, "dosc_long_x": x[::500], "dosc_long_dx": dx[::500]})
    # odeint fp32 through the reference regressor (rk4 full_traj / euler final), lv library with exp
    for name, (d, p, s, e) in {"selkov": (2, 3, False, False), "lv": (2, 2, False, True)}.items():
        reg = make_reg(d, p, s, e)
        reg.Xi.data = torch.tensor(sindy_truth[name], dtype=torch.float32)
        x0 = torch.from_numpy(out[f"{name}_x0"]).float()
        with torch.no_grad():
            tr = ref_mu.odeint(reg, x0, 1.0, 0.002, method='rk4', full_traj=True)
            eu = ref_mu.odeint(reg, x0, 0.1, 0.01, method='euler')
        out.update({f"{name}_odeint_rk4": tr[::25], f"{name}_odeint_euler": eu})
    reg = make_reg(3, 3, False, False, seed=4)
    reg.Xi.data = 0.3 * reg.Xi.data
    x0 = torch.rand(5, 3, generator=torch.Generator().manual_seed(1)) - 0.5
    with torch.no_grad():
        tr = ref_mu.odeint(reg, x0, 0.5, 0.01, method='rk4', full_traj=True)
    out.update({"rand3_Xi": reg.Xi.detach(), "rand3_x0": x0, "rand3_odeint_rk4": tr})
    save("rollout", **out)


def gen_symmreg():
    """a6/a7/a8: symmreg_i / symmreg_r / symmreg_f with seeded frozen stand-ins for the LaLiGAN autoencoder and
    generator (the trained checkpoints are not shipped); loss and gradient w.r.t. Ξ."""
    out = {}
    d = 2
    ae, gen = standins.make_standins(seed=0, input_dim=d, n_comps=2, hidden=32)
    out.update({f"ae_{k}": v for k, v in ae.state_dict().items()})
    out["gen_basis"] = torch.stack(gen.get_full_basis_list())
    out["gen_elems"] = torch.stack(gen.get_deterministic_group_elems())
    g = torch.Generator().manual_seed(2)
    x = torch.rand(96, d, generator=g) * 1.5 + 0.2
    for (p, s, e, tag) in [(2, False, True, "lv"), (3, False, False, "cubic")]:
        reg = make_reg(d, p, s, e, seed=3)
        reg.Xi.data = 0.3 * reg.Xi.data

        def forward_step(z):
            return ref_mu.odeint(reg, z, 0.1, 0.01)

        x_fx = torch.stack([x, forward_step(x)], dim=1)
        li = ref_mu.symmreg_i(x_fx, ae, gen, f=forward_step, require_grad=True)
        reg.zero_grad(); li.backward()
        gi = reg.Xi.grad.clone()
        lr = ref_mu.symmreg_r(x, ae, gen, h=reg, require_grad=True)
        reg.zero_grad(); lr.backward()
        gr = reg.Xi.grad.clone()
        x_fx = torch.stack([x, forward_step(x)], dim=1)
        lf = ref_mu.symmreg_f(x_fx, ae, gen, f=forward_step, require_grad=True)
        reg.zero_grad(); lf.backward()
        gf = reg.Xi.grad.clone()
        # the reference's own precompute (vmap(jacfwd) stacks a single sample on the wrong axis, so its Jacobian is
        # NOT J_g; kept as `_Jgx_ref` to pin our drop-in's identical behaviour) ...
        gx_list, Jgx_ref = ref_mu.precompute_symmreg_r(x, ae, gen)
        # ... and the true per-sample Jacobian of the same group transform, which reproduces symmreg_r's loss
        Jg_true = []
        for gi_, gmat in enumerate(gen.get_deterministic_group_elems(scale=0.01)):
            def move(xx, gmat=gmat):
                zz = ae.encode(torch.stack([xx, xx], dim=1)) - ae.encoder[-2].bias
                gz = torch.einsum('jk,...k->...j', gmat, zz.reshape(zz.shape[0], -1)).reshape(zz.shape)
                return ae.decode(gz + ae.encoder[-2].bias)[:, 0]
            J = torch.autograd.functional.jacobian(lambda xx: move(xx).sum(0), x)  # (d, B, d)
            Jg_true.append(J.permute(1, 0, 2))
        out.update({f"{tag}_Xi": reg.Xi.detach(), f"{tag}_li": li.detach(), f"{tag}_gi": gi, f"{tag}_lr": lr.detach(),
                    f"{tag}_gr": gr, f"{tag}_lf": lf.detach(), f"{tag}_gf": gf,
                    f"{tag}_gx": torch.stack(gx_list), f"{tag}_Jgx_ref": torch.stack(Jgx_ref),
                    f"{tag}_Jgx": torch.stack(Jg_true)})
    out["x"] = x
    save("symmreg", **out)


def gen_lbfgs():
    """a3 end to end: the reference's train_SIGED_lbfgs on noisy dosc data (config-1 shape) from a fixed Ξ init;
    final coefficients and mask."""
    import train as ref_train
    import wandb
    wandb.init(mode="disabled")
    out = {}
    rng = np.random.default_rng(9)
    n = 2500
    x = rng.uniform(-1.5, 1.5, size=(n, 2)).astype(np.float32)
    th = make_reg(2, 2, False, False).eval_Theta_at(torch.from_numpy(x)).numpy().astype(np.float64)
    dx = (th @ sindy_truth["dosc"].T + 0.05 * rng.standard_normal((n, 2))).astype(np.float32)
    loader = [(torch.from_numpy(x), torch.from_numpy(dx))]
    for tag, L_list, lr in (("sindy", [], 0.1), ("esindy", [torch.tensor([[0.0, 1.0], [-1.0, 0.0]])], 0.1)):
        reg = make_reg(2, 2, False, False, L_list=L_list, seed=0)
        init = {k: v.detach().clone() for k, v in reg.state_dict().items()}
        cwd = os.getcwd()
        os.chdir("/tmp")
        try:
            ref_train.train_SIGED_lbfgs(
                train_loader=loader, test_loader=loader, num_epochs=200, device="cpu", log_interval=1000,
                save_interval=100000, save_dir="golden_tmp", autoencoder=torch.nn.Identity(), generator=torch.nn.Identity(),
                regressor=reg, regressor_dst=None, use_latent=False, distill_latent=False, lr_sindy=lr, w_sindy_z=0.0,
                w_sindy_x=1.0, sindy_reg_type='l1', w_sindy_reg=0.0, sym_reg_type='i', w_sym_reg=0.0, st_freq=50,
                threshold=0.05, int_t=0.1, int_dt=0.01, print_eq=False)
        finally:
            os.chdir(cwd)
        Xi = reg.get_Xi().detach() if reg.constraint else reg.Xi.detach()
        out.update({f"{tag}_Xi": Xi, f"{tag}_mask": reg.mask})
        out.update({f"{tag}_init_{k}": v for k, v in init.items()})
    out.update({"x": x, "dx": dx})
    save("lbfgs", **out)


def gen_adam():
    """a3/a6 through the Adam loop: the reference's train_SIGED (data-space branch: MSE + w_sym_reg*symmreg_i + L1,
    thresholding every st_freq epochs) with the frozen stand-in autoencoder/generator."""
    import train as ref_train
    import wandb
    wandb.init(mode="disabled")
    out = {}
    ae, gen = standins.make_standins(seed=1, input_dim=2, n_comps=2, hidden=16)
    out.update({f"ae_{k}": v for k, v in ae.state_dict().items()})
    out["gen_basis"] = torch.stack(gen.get_full_basis_list())
    rng = np.random.default_rng(21)
    n = 512
    x = rng.uniform(0.3, 1.2, size=(n, 2)).astype(np.float32)
    th = make_reg(2, 2, False, False).eval_Theta_at(torch.from_numpy(x)).numpy().astype(np.float64)
    dx = (th @ sindy_truth["growth"].T + 0.01 * rng.standard_normal((n, 2))).astype(np.float32)
    xt, dxt = torch.from_numpy(x), torch.from_numpy(dx)
    loader = [(xt[i:i + 128], dxt[i:i + 128]) for i in range(0, n, 128)]
    reg = make_reg(2, 2, False, False, seed=5)
    reg.Xi.data = 0.2 * reg.Xi.data
    out["init_Xi"] = reg.Xi.detach().clone()
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        ref_train.train_SIGED(
            train_loader=loader, test_loader=loader, num_epochs=6, device="cpu", log_interval=1000,
            save_interval=100000, save_dir="golden_tmp", autoencoder=ae, discriminator=torch.nn.Identity(), generator=gen,
            lr_ae=1e-3, lr_d=1e-3, lr_g=1e-3, w_recon=0.0, w_gan=0.0, w_reg_norm=0.0, w_reg_ortho=0.0,
            w_reg_closure=0.0, use_original_x=False, gan_st_freq=5, gan_st_thres=0.3, ae_arch='mlp', regressor=reg,
            use_latent=False, lr_sindy=2e-2, w_sindy_z=0.0, w_sindy_x=1.0, sindy_reg_type='l1', w_sindy_reg=1e-3,
            w_sym_reg=0.05, st_freq=3, threshold=0.02, int_t=0.1, int_dt=0.01, print_eq=False, print_li=False)
    finally:
        os.chdir(cwd)
    out.update({"x": x, "dx": dx, "final_Xi": reg.Xi.detach(), "final_mask": reg.mask})
    # the same loop with w_sym_reg = 0 (the regulariser is still evaluated and logged by the reference but does not
    # move the parameters): pins the one-launch fused iteration of train.train_SIGED
    reg0 = make_reg(2, 2, False, False, seed=5)
    reg0.Xi.data = out["init_Xi"].clone()
    os.chdir("/tmp")
    try:
        ref_train.train_SIGED(
            train_loader=loader, test_loader=loader, num_epochs=9, device="cpu", log_interval=1000,
            save_interval=100000, save_dir="golden_tmp", autoencoder=ae, discriminator=torch.nn.Identity(), generator=gen,
            lr_ae=1e-3, lr_d=1e-3, lr_g=1e-3, w_recon=0.0, w_gan=0.0, w_reg_norm=0.0, w_reg_ortho=0.0,
            w_reg_closure=0.0, use_original_x=False, gan_st_freq=5, gan_st_thres=0.3, ae_arch='mlp', regressor=reg0,
            use_latent=False, lr_sindy=2e-2, w_sindy_z=0.0, w_sindy_x=0.8, sindy_reg_type='l1', w_sindy_reg=1e-3,
            w_sym_reg=0.0, st_freq=3, threshold=0.02, int_t=0.1, int_dt=0.01, print_eq=False, print_li=False)
    finally:
        os.chdir(cwd)
    out.update({"nosym_final_Xi": reg0.Xi.detach(), "nosym_final_mask": reg0.mask})
    save("adam", **out)


def gen_smoothing():
    """GP smoother + derivative of the data generators (`data_utils/smoothing.py:155-196` num_diff_gp, called from
    `data_utils/ode.py:43-45`), on small noisy damped-oscillator and Lotka-Volterra batches."""
    from data_utils.smoothing import num_diff_gp
    out = {}
    rng = np.random.default_rng(5)
    for name, f, dt, T, n_traj, noise, sigma_in in (("dosc", damped_oscillator.dosc, 0.01, 300, 4, 0.2, 0.1),
                                                    ("lv", lotka.lotka_volterra, 0.02, 200, 3, 0.05, None)):
        x0 = rng.uniform(0.5, 1.5, (n_traj, 2))
        x, _ = ref_ode.solve_ode_batch(f, x0, dt=dt, num_steps=T)
        std = np.std(x, axis=(0, 1))
        xn = x + rng.standard_normal(x.shape) * noise * std
        dX, Xs = num_diff_gp(xn.copy(), dt, noise_level=noise, std_base=std, sigma_in=sigma_in)
        out.update({name + "_x": xn, name + "_std": std, name + "_dt": dt, name + "_noise": noise,
                    name + "_sigma_in": (np.nan if sigma_in is None else sigma_in), name + "_dX": dX, name + "_X": Xs})
    save("smoothing", **out)


def gen_api():
    """The reference's public surface on the hot path as data: parameter names and defaults of the
    functions / methods the drop-in modules must reproduce. Stored as JSON next to the numeric fixtures."""
    import inspect
    import json
    import train as ref_train

    def sig(fn):
        out = []
        for name, p in inspect.signature(fn).parameters.items():
            d = None if p.default is inspect._empty else repr(p.default)
            out.append([name, str(p.kind), d])
        return out

    api = {}
    for modname, mod in (("sindy", ref_sindy), ("model_utils", ref_mu), ("data_utils.ode", ref_ode), ("train", ref_train)):
        entry = {}
        for name, obj in vars(mod).items():
            if name.startswith("_") or getattr(obj, "__module__", None) != mod.__name__:
                continue
            if inspect.isfunction(obj):
                entry[name] = {"kind": "function", "params": sig(obj)}
            elif inspect.isclass(obj):
                methods = {m: sig(f) for m, f in vars(obj).items() if inspect.isfunction(f) and (not m.startswith("_") or m == "__init__")}
                entry[name] = {"kind": "class", "methods": methods}
        api[modname] = entry
    path = os.path.join(OUT, "api_surface.json")
    with open(path, "w") as f:
        json.dump(api, f, indent=1, sort_keys=True)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["model", "jvp", "stlsq", "wsindy", "rollout", "symmreg", "lbfgs", "adam", "smoothing", "api"]
    for w in which:
        globals()["gen_" + w]()
