"""GPU parity of the training loops: the Adam loop with the infinitesimal symmetry regulariser (reference
`train.py:382-614`, data-space branch) against a golden run of the unmodified reference, and the closure-free
LBFGS fit on cached sufficient statistics against the standard fused-closure fit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype).cuda()


def rel(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-30)


def test_golden_adam_loop_with_symmreg_i(golden):
    import sindy
    import train
    import standins
    g = golden("adam")
    sd = {k[3:]: torch.as_tensor(g[k]) for k in g.keys() if k.startswith("ae_")}
    ae, gen = standins.make_standins(seed=1, input_dim=2, n_comps=2, hidden=16, state_dict=sd, device="cuda")
    assert rel(torch.stack(gen.get_full_basis_list()), g["gen_basis"]) == 0
    x, dx = dev(g["x"]), dev(g["dx"])
    loader = [(x[i:i + 128], dx[i:i + 128]) for i in range(0, x.shape[0], 128)]
    reg = sindy.SINDyRegression(2, 2, False, False, threshold=0.02, device="cuda", constrain_constant=True)
    reg.Xi.data = dev(g["init_Xi"])
    train.train_SIGED(
        train_loader=loader, test_loader=loader, num_epochs=6, device="cuda", log_interval=1000, save_interval=100000,
        save_dir=None, autoencoder=ae, discriminator=torch.nn.Identity(), generator=gen, lr_ae=1e-3, lr_d=1e-3,
        lr_g=1e-3, w_recon=0.0, w_gan=0.0, w_reg_norm=0.0, w_reg_ortho=0.0, w_reg_closure=0.0, use_original_x=False,
        gan_st_freq=5, gan_st_thres=0.3, ae_arch='mlp', regressor=reg, use_latent=False, lr_sindy=2e-2, w_sindy_z=0.0,
        w_sindy_x=1.0, sindy_reg_type='l1', w_sindy_reg=1e-3, w_sym_reg=0.05, st_freq=3, threshold=0.02, int_t=0.1,
        int_dt=0.01, print_eq=False, print_li=False)
    assert np.array_equal(reg.mask.cpu().numpy(), g["final_mask"])
    # 24 Adam steps through symmreg_i: fp32 differences accumulate, the trajectories stay together to ~1e-3
    assert rel(reg.Xi, g["final_Xi"]) < 2e-3, rel(reg.Xi, g["final_Xi"])


def test_cached_gram_lbfgs_matches_fused_closure_fit(golden):
    import sindy
    import train
    g = golden("lbfgs")
    x, dx = dev(g["x"]), dev(g["dx"])
    loader = [(x, dx)]
    fits = []
    for cached in (False, True):
        reg = sindy.SINDyRegression(2, 2, False, False, threshold=0.05, device="cuda", constrain_constant=True)
        reg.Xi.data = dev(g["sindy_init_Xi"])
        train.train_SIGED_lbfgs(
            train_loader=loader, test_loader=loader, num_epochs=200, device="cuda", log_interval=1000,
            save_interval=100000, save_dir=None, autoencoder=torch.nn.Identity(), generator=torch.nn.Identity(),
            regressor=reg, regressor_dst=None, use_latent=False, distill_latent=False, lr_sindy=0.1, w_sindy_z=0.0,
            w_sindy_x=1.0, sindy_reg_type='l1', w_sindy_reg=0.0, sym_reg_type='i', w_sym_reg=0.0, st_freq=50,
            threshold=0.05, int_t=0.1, int_dt=0.01, print_eq=False, cached_gram=cached)
        fits.append((reg.Xi.detach().clone(), reg.mask.clone()))
    assert torch.equal(fits[0][1], fits[1][1])
    assert np.array_equal(fits[1][1].cpu().numpy(), g["sindy_mask"])
    assert rel(fits[1][0] * fits[1][1], g["sindy_Xi"] * g["sindy_mask"]) < 1e-3
    # value and gradient of the quadratic form agree with the data pass
    reg = sindy.SINDyRegression(2, 2, False, False, threshold=0.05, device="cuda", constrain_constant=True)
    stats = reg.sufficient_statistics(x, dx)
    la = reg.mse_loss(x, dx); la.backward(); ga = reg.Xi.grad.clone(); reg.zero_grad()
    lb = reg.mse_loss_from_statistics(stats); lb.backward()
    assert abs(float(la) - float(lb)) < 1e-5 * abs(float(la)) and rel(reg.Xi.grad, ga) < 1e-4
