"""Training loops over the B200 SINDy operators, call-compatible with the reference's `train.py`
(`train_SIGED_lbfgs` :617-852, `train_SIGED` :382-614 equation part, `train_WSINDy` :855-869, `train_SINDy` :872-887).

The reference's own `train.py` also runs unchanged on top of the drop-in `sindy.py` / `model_utils.py`; these
loops exist because they use the fused one-pass train step (`SINDyRegression.mse_loss`: loss and dL/dΞ from a
single sweep over the batch, no N×K or N×d intermediate, one host sync per closure instead of 3-4) and because
the GPU box has no copy of the reference. Optimiser, thresholding schedule, convergence tests, NaN guard,
printed messages and checkpoint names follow the reference. `train_lassi` (LaLiGAN symmetry discovery,
`train.py:16-269`) is outside the hot path: when a reference checkout is importable (it follows this directory on
`sys.path`, or $SINDY_B200_REFERENCE names it) its own `train_lassi` is re-exported, running on this repo's `sindy` /
`model_utils` operators, so that `main.py:90-91` keeps working; without a reference the name is simply absent.
"""
from __future__ import annotations

import os
from copy import deepcopy

import numpy as np
import torch

from model_utils import (EulerFlowMap, encode_constant_component, group_action_and_jacobian, make_fsymmreg_pttrain, make_rsymmreg_pttrain,
                         make_symmreg_pttrain, odeint, symmreg_r_precomputed)
from sindy import solve_SINDy_one_step

__all__ = ["train_SIGED_lbfgs", "train_SIGED", "train_WSINDy", "train_SINDy"]


def _reference_train_module():
    """The reference's train.py loaded under a private name (None if no reference checkout is reachable)."""
    import importlib.util
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for base in [os.environ.get("SINDY_B200_REFERENCE")] + list(sys.path):
        if not base or os.path.abspath(base) == here:
            continue
        path = os.path.join(os.path.abspath(base), "train.py")
        if os.path.isfile(path) and os.path.isfile(os.path.join(os.path.dirname(path), "gan.py")):
            try:
                spec = importlib.util.spec_from_file_location("_sindy_b200_reference_train", path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                return mod
            except Exception:  # noqa: BLE001  (a reference that does not import leaves train_lassi undefined)
                return None
    return None


_ref_train = _reference_train_module()
if _ref_train is not None and hasattr(_ref_train, "train_lassi"):
    train_lassi = _ref_train.train_lassi
    __all__.append("train_lassi")

_TOL = 1e-3  # LBFGS convergence tolerance on the parameter update (reference `train.py:643`)


def _log(metrics):
    """wandb.log when a run is active (the reference logs unconditionally, `train.py:761`)."""
    try:
        import wandb
        if wandb.run is not None:
            wandb.log(metrics)
    except ImportError:
        pass


def _save(regressor, save_dir, name):
    if save_dir is None:
        return
    os.makedirs(f'saved_models/{save_dir}', exist_ok=True)
    torch.save(regressor.state_dict(), f'saved_models/{save_dir}/{name}')


def _snapshot(module):
    return [p.detach().clone() for p in module.parameters()]


def _moved(module, old):
    with torch.no_grad():
        return sum(torch.norm(p - q) for p, q in zip(module.parameters(), old))


def _l1(module):
    return sum(torch.norm(p, 1) for p in module.parameters())


def _lbfgs_phase(regressor, data_loss, num_epochs, lr_sindy, sindy_reg_type, w_sindy_reg, st_freq, threshold,
                 log_interval, save_interval, save_dir, on_log=None):
    """LBFGS with sequential thresholding (reference `train.py:692-766` and `:805-852`).

    `data_loss(losses)` returns the differentiable data term and records scalars in `losses`.
    Whenever the parameters stop moving (< 1e-3 per epoch): threshold + fresh optimiser; if they also have not
    moved since the previous thresholding: converged. Every `st_freq` epochs without convergence: threshold anyway.
    """
    if sindy_reg_type not in ('l1', 'none'):
        raise ValueError(f'Unknown regularization type: {sindy_reg_type}')
    losses = {}
    state = {'opt': torch.optim.LBFGS(regressor.parameters(), lr=lr_sindy)}

    def closure():
        state['opt'].zero_grad()
        loss = data_loss(losses)
        if sindy_reg_type == 'l1':
            reg = _l1(regressor)
            losses['loss_sindy_reg'] = reg.detach()
            loss = loss + w_sindy_reg * reg
        loss.backward()
        return loss

    prev, since_threshold = _snapshot(regressor), _snapshot(regressor)
    n_iters = 0
    for epoch in range(num_epochs):
        n_iters += 1
        state['opt'].step(closure)
        if any(torch.isnan(p).any() for p in regressor.parameters()):
            print(f'NaN encountered at iteration {epoch}; exit training.')
            break
        metrics = deepcopy({k: float(v) for k, v in losses.items()})
        if _moved(regressor, prev) < _TOL:
            if _moved(regressor, since_threshold) < _TOL:
                print(f'Final convergence reached at iteration {epoch}; exit training.')
                _save(regressor, save_dir, f'regressor_{epoch}.pt')
                break
            n_iters = 0
            regressor.set_threshold(threshold)
            state['opt'] = torch.optim.LBFGS(regressor.parameters(), lr=lr_sindy)
            since_threshold = _snapshot(regressor)
            print(f'Convergence reached at iteration {epoch}; apply parameter thresholding and reset optimizer.')
        elif st_freq > 0 and n_iters % st_freq == 0:
            n_iters = 0
            regressor.set_threshold(threshold)
            state['opt'] = torch.optim.LBFGS(regressor.parameters(), lr=lr_sindy)
            print('Max number of LBFGS iterations reached; apply parameter thresholding and reset optimizer.')
        prev = _snapshot(regressor)

        if (epoch + 1) % log_interval == 0:
            print(', '.join([f'Epoch {epoch}'] + [f'{k}: {metrics[k]:.4f}' for k in metrics]))
            if on_log is not None:
                on_log(epoch, metrics)
        _log(metrics)
        if (epoch + 1) % save_interval == 0:
            _save(regressor, save_dir, f'regressor_{epoch}.pt')


def train_SIGED_lbfgs(
    train_loader, test_loader, num_epochs, device, log_interval, save_interval, save_dir,  # global
    autoencoder, generator,  # symmetry discovery model
    regressor, regressor_dst, use_latent, distill_latent, lr_sindy, w_sindy_z, w_sindy_x,  # SINDy
    sindy_reg_type, w_sindy_reg, sym_reg_type, w_sym_reg, st_freq, threshold, int_t, int_dt,  # SINDy
    **kwargs
):
    if distill_latent and not use_latent:
        raise ValueError('Cannot distill without first learning latent space equation. Set use_latent=True.')
    train_data = next(iter(train_loader))
    x, dx = (t.to(device) for t in train_data)
    symm_loss = {'i': make_symmreg_pttrain, 'f': make_fsymmreg_pttrain,
                 'r': make_rsymmreg_pttrain}[sym_reg_type](autoencoder, generator) if w_sym_reg > 0.0 else None
    autoencoder.eval()
    generator.eval()
    mse = torch.nn.MSELoss()
    # cached_gram=True (extension): the batch is fixed for the whole fit (`train.py:626-629`), so the
    # MSE is an exact quadratic in Ξ given G = ΘᵀΘ, b = ΘᵀẊ, Σẋ²: ONE data pass for the fit, every closure is K×K algebra
    stats = None
    if kwargs.get('cached_gram') and not use_latent:
        stats = regressor.sufficient_statistics(x, dx)
    group_cache = None
    if w_sym_reg > 0.0 and sym_reg_type == 'r' and not use_latent and kwargs.get('precompute_group', True):
        group_cache = group_action_and_jacobian(x, autoencoder, generator)
    # symmreg_i encodes [x, f(x)]: the x half never changes during the fit and needs no gradient
    z_x = encode_constant_component(autoencoder, x) if (w_sym_reg > 0.0 and sym_reg_type == 'i' and not use_latent) else None

    def data_loss(losses):
        if use_latent:
            z, _ = autoencoder(x)
            dz = autoencoder.compute_dz(x, dx)
            dz_pred = regressor(z)
            dx_pred = autoencoder.compute_dx(z, dz_pred)
            loss_z, loss_x = mse(dz_pred, dz), mse(dx_pred, dx)
            losses['loss_sindy_z'], losses['loss_sindy_x'] = loss_z.detach(), loss_x.detach()
            return w_sindy_z * loss_z + w_sindy_x * loss_x
        if stats is not None:
            loss_x = regressor.mse_loss_from_statistics(stats)   # closure-free: K×K algebra, no data pass
        else:
            loss_x = regressor.mse_loss(x, dx)      # fused: value and dL/dΞ from one pass
        losses['loss_sindy_x'] = loss_x.detach()
        loss = w_sindy_x * loss_x
        if w_sym_reg > 0.0:
            if sym_reg_type in ('i', 'f'):
                # the flow map of `train.py:669-673` as an object that knows its own JVP: symmreg_i then takes
                # (f(x), J_f(x)·v) from one fused launch instead of a double vjp through the Euler steps
                forward_step = EulerFlowMap(regressor, int_t, int_dt)
                x_fx = torch.stack([x, forward_step(x)], dim=1)
                loss_sym = symm_loss(x_fx, f=forward_step, **({'z_x': z_x} if z_x is not None else {}))
            elif group_cache:
                # g(x), J_g(x) do not depend on Ξ: formed once for the (fixed) batch, then one streaming launch per
                # group element and closure (`model_utils.py:126-170`)
                loss_sym = symmreg_r_precomputed(x, group_cache[0], group_cache[1], regressor)
            else:
                loss_sym = symm_loss(x, h=regressor)
            losses['loss_sym_reg'] = loss_sym.detach()
            loss = loss + w_sym_reg * loss_sym
        return loss

    def on_log(epoch, metrics):
        # the reference's "test" loop evaluates the training batch once per validation batch (`train.py:739-751`)
        with torch.no_grad():
            if not use_latent:
                val = float(regressor.mse_loss(x, dx))
                print(f'Epoch {epoch}, test_loss_sindy_z: 0.0000, test_loss_sindy_x: {val:.4f}')
                metrics.update({'test_loss_sindy_z': 0.0, 'test_loss_sindy_x': val})
        if kwargs.get('print_eq'):
            regressor.print()

    _lbfgs_phase(regressor, data_loss, num_epochs, lr_sindy, sindy_reg_type, w_sindy_reg, st_freq, threshold,
                 log_interval, save_interval, save_dir, on_log)

    if not distill_latent:
        return
    print('\n=== Phase 2: distill equation from latent to data space ===\n')
    with torch.no_grad():
        z, _ = autoencoder(x)
        dx_latent = autoencoder.compute_dx(z, regressor(z))

    def distill_loss(losses):
        loss_x = regressor_dst.mse_loss(x, dx_latent)
        losses['loss_sindy_x'] = loss_x.detach()
        return w_sindy_x * loss_x

    def on_log_dst(epoch, metrics):
        if kwargs.get('print_eq'):
            regressor_dst.print()

    _lbfgs_phase(regressor_dst, distill_loss, num_epochs, lr_sindy, sindy_reg_type, w_sindy_reg, st_freq, threshold,
                 log_interval, save_interval, save_dir, on_log_dst)


def train_SIGED(
    train_loader, test_loader, num_epochs, device, log_interval, save_interval, save_dir,
    autoencoder, discriminator, generator,
    lr_ae, lr_d, lr_g, w_recon, w_gan, w_reg_norm, w_reg_ortho, w_reg_closure,
    use_original_x, gan_st_freq, gan_st_thres, ae_arch,
    regressor, use_latent, lr_sindy, w_sindy_z, w_sindy_x, sindy_reg_type, w_sindy_reg, w_sym_reg, st_freq, threshold,
    int_t, int_dt, **kwargs
):
    """Adam loop of the reference (`train.py:382-614`), equation-discovery part (the GAN/autoencoder updates are
    commented out there too). Data-space branch: MSE + w_sym_reg·symmreg_i + L1; thresholding every st_freq epochs.
    With w_sym_reg == 0 and an unconstrained regressor on a specialised library the loop runs fused, one launch per
    iteration (`fused_step=False` keeps the operator-by-operator path, which also logs the unweighted sym-reg value)."""
    if sindy_reg_type != 'l1':
        raise ValueError(f'Unknown regularization type: {sindy_reg_type}')
    if use_latent:
        raise NotImplementedError('train_SIGED(use_latent=True): the latent Adam branch of the reference raises '
                                  'TypeError as shipped (train.py:505); use train_SIGED_lbfgs')
    if kwargs.get('fused_step', True) and w_sym_reg == 0.0 and _fusable(regressor):
        # the regulariser has weight 0: it cannot move the parameters (the reference still evaluates it for its log
        # line). Every batch iteration — forward, MSE + L1, backward, Adam update — is ONE kernel launch (sb_fit_step).
        return _train_adam_fused(train_loader, num_epochs, device, log_interval, save_interval, save_dir, regressor,
                                 lr_sindy, w_sindy_x, w_sindy_reg, st_freq, threshold, kwargs.get('print_eq'))
    optimizer = torch.optim.Adam(regressor.parameters(), lr=lr_sindy)
    symm_loss = make_symmreg_pttrain(autoencoder, generator)
    for epoch in range(num_epochs):
        running = {'loss_sindy_x': [], 'loss_sym_reg': [], 'loss_sindy_reg': []}
        regressor.train()   # the autoencoder / generator stay in eval mode (the reference's .train() calls are commented out)
        for x, dx in train_loader:
            x, dx = x.to(device), dx.to(device)
            loss_x = regressor.mse_loss(x, dx)

            def forward_step(q):
                return odeint(regressor, q, int_t, int_dt)

            x_fx = torch.stack([x, forward_step(x)], dim=1)
            loss_sym = symm_loss(x_fx, f=forward_step)
            loss_reg = _l1(regressor)
            loss = w_sindy_x * loss_x + w_sym_reg * loss_sym + w_sindy_reg * loss_reg
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            for k, v in (('loss_sindy_x', loss_x), ('loss_sym_reg', loss_sym), ('loss_sindy_reg', loss_reg)):
                running[k].append(float(v))
        if st_freq > 0 and (epoch + 1) % st_freq == 0:
            regressor.set_threshold(threshold)
        metrics = {k: float(np.mean(v)) for k, v in running.items() if v}
        if (epoch + 1) % log_interval == 0:
            print(', '.join([f'Epoch {epoch}'] + [f'{k}: {v:.4f}' for k, v in metrics.items()]))
            if kwargs.get('print_eq'):
                regressor.print()
        _log(metrics)
        if (epoch + 1) % save_interval == 0:
            _save(regressor, save_dir, f'regressor_{epoch}.pt')


def _fusable(regressor):
    """Unconstrained regressor on a CUDA device whose library has a specialised fused kernel."""
    from sindy_b200 import native
    if getattr(regressor, 'constraint', False) or not isinstance(getattr(regressor, 'Xi', None), torch.nn.Parameter):
        return False
    if not regressor.Xi.is_cuda:
        return False
    return native.train_step_variant(regressor.library).startswith('fused')


def _train_adam_fused(train_loader, num_epochs, device, log_interval, save_interval, save_dir, regressor, lr_sindy,
                      w_sindy_x, w_sindy_reg, st_freq, threshold, print_eq):
    """Adam loop of `train.py:491-540` (data-space branch, w_sym_reg = 0) with one launch per iteration. The kernel
    updates `regressor.Xi` in place with torch's Adam arithmetic and leaves Ξ⊙mask in the constant bank for the next
    batch; anything else that may load coefficients (thresholding, printing) clears that promise."""
    from sindy_b200 import native
    lib = regressor.library
    xi = regressor.Xi.data
    if xi.dtype != torch.float32 or not xi.is_contiguous():
        raise ValueError('regressor.Xi must be a contiguous float32 parameter')
    state = native.fit_state(lib, xi.device)
    resident, gen = False, -1
    for epoch in range(num_epochs):
        mse_terms, loss_terms = [], []
        regressor.train()
        for x, dx in train_loader:
            x, dx = x.to(device), dx.to(device)
            n = x.reshape(-1, lib.dim).shape[0]
            if native.slot_generation(xi.device) != gen:      # another fit has loaded the resident slot since
                resident = False
            loss, _, packed = native.fit_step(x, dx, xi, regressor.mask, lib, 'adam', lr_sindy, w_mse=w_sindy_x,
                                              w_l1=w_sindy_reg, state=state, w_resident=resident)
            resident, gen = True, native.slot_generation(xi.device)
            mse_terms.append(packed[0] / (n * lib.dim))     # device scalars: no host sync inside the epoch
            loss_terms.append(loss.clone())
        if st_freq > 0 and (epoch + 1) % st_freq == 0:
            regressor.set_threshold(threshold)
            resident = False
        mse = torch.stack(mse_terms).double()
        tot = torch.stack(loss_terms).double()
        metrics = {'loss_sindy_x': float(mse.mean())}
        if w_sindy_reg != 0.0:
            metrics['loss_sindy_reg'] = float(((tot - w_sindy_x * mse) / w_sindy_reg).mean())
        if (epoch + 1) % log_interval == 0:
            print(', '.join([f'Epoch {epoch}'] + [f'{k}: {v:.4f}' for k, v in metrics.items()]))
            if print_eq:
                regressor.print()
                resident = False
        _log(metrics)
        if (epoch + 1) % save_interval == 0:
            _save(regressor, save_dir, f'regressor_{epoch}.pt')


def _solve_loop(step, printer, num_epochs, log_interval):
    for epoch in range(num_epochs):
        residual, completed = step()
        if (epoch + 1) % log_interval == 0:
            print(f'Iteration {epoch}, loss: {residual:.4f}')
            printer()
        if completed:
            print(f'Final convergence reached at iteration {epoch}; exit training.')
            break


def train_WSINDy(wrapper, train_x, num_epochs, device, log_interval, save_interval, save_dir, w_sindy_reg, threshold,
                 **kwargs):
    train_x = train_x.to(device)
    _solve_loop(lambda: wrapper.solve(train_x, w_sindy_reg, threshold), wrapper.regressor.print, num_epochs,
                log_interval)


def train_SINDy(regressor, x, dx, num_epochs, device, log_interval, save_interval, save_dir, w_sindy_reg, threshold,
                **kwargs):
    x, dx = x.to(device), dx.to(device)
    _solve_loop(lambda: solve_SINDy_one_step(regressor, x, dx, w_sindy_reg, threshold), regressor.print, num_epochs,
                log_interval)
