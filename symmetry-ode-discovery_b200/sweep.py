"""Seed sweeps in ONE process.

The reference repeats every experiment for seeds 0..49 by starting `python main.py --seed i --config ...` fifty times
(`run_scripts/*.sh`; each start pays ≈ 18 s of imports and data loading for a fit that takes a fraction of that) and
then aggregates the per-seed `.npz` files with `evaluation/eval_eq.py:38-85`. Here the seeds of a sweep run back to
back in one process on data that is already on the device — the LBFGS fit of every seed on the closed-form objective of
its own subsample (one data pass per seed, `train_SIGED_lbfgs(cached_gram=True)`) — and, under `torchrun`, are dealt
round-robin to the ranks (seeds are independent: no collective on the data path, one gather of the results).

`evaluate` and `aggregate` restate the reference's metrics (`eval_eq.py:7-34`, `:38-85`): per-equation "correct form"
(mask equals the truth's support), MSE over the truth's non-zero coefficients, success counts and RMSE mean (std).
What a seed controls follows `main.py:25-77`: the shuffled subsample of the training set that LBFGS sees
(`lbfgs_subsample`) and the random initial coefficients. The draws themselves are this module's own (the reference's
draw order runs through its autoencoder / GAN constructors, which are not part of this repository), so sweeps are
statistically, not seed-for-seed, comparable with the reference's.
"""
from __future__ import annotations

import io
from contextlib import redirect_stdout
from typing import Callable, Dict, Iterable, List, Optional

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["evaluate", "aggregate", "run_seed_sweep", "lbfgs_fit"]


def evaluate(regressor, truth) -> Dict[str, np.ndarray]:
    """`eval_sindy_regressor` (`eval_eq.py:7-34`): masked coefficients, per-equation correct form and MSE over the
    truth's support, and both over all equations."""
    with torch.no_grad():
        coef = regressor.get_Xi() if regressor.constraint else regressor.Xi
        coef = coef.detach().cpu().numpy()
        mask = regressor.mask.bool().cpu().numpy()
    truth = np.asarray(truth, dtype=np.float64)
    coef = np.where(mask, coef, 0.0)
    truth_mask = truth != 0
    correct = np.array([float(np.all(mask[i] == truth_mask[i])) for i in range(coef.shape[0])])
    mse = np.array([np.mean((coef[i, truth_mask[i]] - truth[i, truth_mask[i]]) ** 2) for i in range(coef.shape[0])])
    return {"coefficients": coef, "correct_form": correct, "mse": mse, "correct_form_all": bool(np.all(correct)),
            "mse_all": float(np.mean(mse))}


def aggregate(results: Iterable[Dict], mse_multiplier: float = 1.0, verbose: bool = True) -> Dict:
    """`aggregate_results` (`eval_eq.py:38-85`) on in-memory results: success counts per equation and jointly, RMSE
    mean and std over the successful runs and over all runs (the reference prints these lines; so does this)."""
    results = list(results)
    cf = np.stack([r["correct_form"] for r in results])
    cf_all = np.array([bool(r["correct_form_all"]) for r in results])
    rmse = np.sqrt(np.stack([r["mse"] for r in results]))
    rmse_all = np.sqrt(np.array([r["mse_all"] for r in results]))
    out = {"runs": len(results), "success": cf.sum(0).astype(int), "success_all": int(cf_all.sum()),
           "rmse": [], "rmse_any": []}

    def stats(v):
        return (float(np.mean(v) * mse_multiplier), float(np.std(v) * mse_multiplier)) if len(v) else (float("nan"),) * 2

    for i in range(cf.shape[1]):
        out["rmse"].append(stats(rmse[cf[:, i] > 0, i]))
        out["rmse_any"].append(stats(rmse[:, i]))
    out["rmse_all"] = stats(rmse_all[cf_all])
    out["rmse_all_any"] = stats(rmse_all)
    if verbose:
        print(f'Loaded results from {out["runs"]} runs.')
        for i, s in enumerate(out["success"]):
            print(f'Equation {i} success rate = {s}/{out["runs"]}')
        print(f'Joint success rate = {out["success_all"]}/{out["runs"]}')
        for i in range(cf.shape[1]):
            print(f'Equation {i} RMSE = {out["rmse"][i][0]:.4f} ({out["rmse"][i][1]:.4f})')
            print(f'Equation {i} RMSE (any) = {out["rmse_any"][i][0]:.4f} ({out["rmse_any"][i][1]:.4f})')
        print(f'All equations RMSE = {out["rmse_all"][0]:.4f} ({out["rmse_all"][1]:.4f})')
        print(f'All equations RMSE (any) = {out["rmse_all_any"][0]:.4f} ({out["rmse_all_any"][1]:.4f})')
    return out


def lbfgs_fit(regressor, x, dx, lr_sindy=0.1, w_sindy_reg=0.0, st_freq=50, threshold=0.05, num_epochs=200, **kwargs):
    """The fit of `main.py` with `--sindy_optimizer lbfgs` and no sym-reg (`train_SIGED_lbfgs`, data-space branch) on
    the closed-form objective of this batch: one data pass, then K×K algebra per closure."""
    import train
    ident = torch.nn.Identity()
    train.train_SIGED_lbfgs(
        train_loader=[(x, dx)], test_loader=[(x, dx)], num_epochs=num_epochs, device=x.device, log_interval=10 ** 9,
        save_interval=10 ** 9, save_dir=None, autoencoder=ident, generator=ident, regressor=regressor,
        regressor_dst=None, use_latent=False, distill_latent=False, lr_sindy=lr_sindy, w_sindy_z=0.0, w_sindy_x=1.0,
        sindy_reg_type='l1' if w_sindy_reg > 0 else 'none', w_sindy_reg=w_sindy_reg, sym_reg_type='i', w_sym_reg=0.0,
        st_freq=st_freq, threshold=threshold, int_t=0.1, int_dt=0.01, print_eq=False, cached_gram=True)


def run_seed_sweep(x: torch.Tensor, dx: torch.Tensor, truth, seeds: Iterable[int], make_regressor: Callable[[], object],
                   fit: Callable = lbfgs_fit, subsample: float = 0.5, group=None, quiet: bool = True,
                   **fit_kwargs) -> Optional[List[Dict]]:
    """Fit one regressor per seed on a shuffled `subsample` of (x, dx) (`main.py:33-37`: the single LBFGS batch) and
    evaluate it against `truth`. x, dx: (N, d) CUDA tensors. Under torch.distributed the seeds are dealt round-robin to
    the ranks and the per-seed results are gathered on every rank in seed order. Returns the list of result dicts
    (each with its 'seed')."""
    seeds = list(seeds)
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    xf, dxf = x.reshape(-1, x.shape[-1]), dx.reshape(-1, dx.shape[-1])
    n = xf.shape[0]
    take = max(1, int(n * subsample))
    mine = []
    for s in seeds[rank::world]:
        torch.manual_seed(int(s))
        np.random.seed(int(s))
        idx = torch.randperm(n, device=xf.device)[:take]
        xb, dxb = xf[idx].contiguous(), dxf[idx].contiguous()
        regressor = make_regressor()
        sink = io.StringIO()
        with redirect_stdout(sink) if quiet else _nullcontext():
            fit(regressor, xb, dxb, **fit_kwargs)
        res = evaluate(regressor, truth)
        res["seed"] = int(s)
        mine.append(res)
    if not distributed or world == 1:
        return mine
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    return sorted((r for part in gathered for r in part), key=lambda r: seeds.index(r["seed"]))


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False
