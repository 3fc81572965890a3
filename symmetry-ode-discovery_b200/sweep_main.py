"""Seed sweep of one reference cfg in ONE process — the replacement for `run_scripts/*.sh` (50 × `python main.py --seed i
--config ...`, each paying the interpreter start, imports and data loading).

    python symmetry-ode-discovery_b200/sindy_b200/run.py --reference <reference> \
        symmetry-ode-discovery_b200/sweep_main.py --config dosc/noise20_sindy.cfg --seeds 0-49 [--gpu 0]

Run through the launcher: the argument parser, data set, autoencoder / discriminator / generator classes and the truth
tables come from the reference checkout, the equation model and the fit from this repository. For every seed the script
replays the random draws of `main.py:25-77` IN THE REFERENCE'S ORDER — `torch.manual_seed`, the constructors of
AutoEncoder, Discriminator and LieGenerator (they consume the generator even when unused), the regressor's `randn`, and
only then the shuffled LBFGS batch (`train.py:626-627`: the DataLoader draws its permutation at the first `iter`) — so
seed k of the sweep starts exactly where `python main.py --seed k` starts. All seeds are then fitted as ONE batched
LBFGS problem on the closed-form objective of their own subsample (`sweep.batched_lbfgs_fits`), evaluated like
`main.py:120-138` (`eval_results/<save_dir>/seed<k>.npz`) and aggregated like `evaluation/eval_eq.py:38-85`.

Supported: the LBFGS cfgs without a symmetry regulariser and without an autoencoder in the loss (`dosc/noise20_sindy`,
`dosc/noise20_esindy`, `growth/noise05_sindy`, `growth/noise05_esindy`); anything else exits with a message.
"""
import json
import os
import sys

import numpy as np
import torch


def parse_seeds(spec):
    out = []
    for part in spec.split(","):
        lo, _, hi = part.partition("-")
        out.extend(range(int(lo), int(hi or lo) + 1))
    return out


def main():
    argv = sys.argv[1:]
    seeds, json_out = [0], None
    for flag in ("--seeds", "--json"):
        if flag in argv:
            i = argv.index(flag)
            val = argv[i + 1]
            del argv[i:i + 2]
            if flag == "--seeds":
                seeds = parse_seeds(val)
            else:
                json_out = val
    sys.argv = [sys.argv[0]] + argv
    from torch.utils.data import DataLoader, TensorDataset
    from parser_utils import get_args          # reference
    from dataset import get_dataset            # reference
    from autoencoder import AutoEncoder        # reference
    from gan import Discriminator, LieGenerator
    from evaluation.eval_eq import sindy_truth
    import sindy
    import sweep

    args = vars(get_args())
    if args['sindy_optimizer'] != 'lbfgs' or args['w_sym_reg'] > 0.0 or args['use_latent'] or args['mt_data']:
        sys.exit("sweep_main: only LBFGS cfgs without sym-reg / latent space are batched; run main.py per seed instead")
    train_dataset, _, args = get_dataset(args)
    dev = args['device']
    x_all, dx_all = train_dataset.x.to(dev), train_dataset.dx.to(dev)
    n = len(train_dataset)
    take = int(n * args['lbfgs_subsample'])
    index_set = TensorDataset(torch.arange(n))

    def draw(seed, n_, take_):
        torch.manual_seed(seed)                                     # main.py:25-27
        np.random.seed(seed)
        loader = DataLoader(index_set, batch_size=take, shuffle=True)   # main.py:33-37 (no draw yet)
        AutoEncoder(**args); Discriminator(**args)                  # main.py:42-44: consume the generator like main.py
        generator = LieGenerator(**args)
        a = dict(args)
        if a['eq_constraint']:                                      # main.py:72-76
            L_list = generator.get_full_basis_list()
            repr_dim = L_list[0].shape[-1] // a['n_comps']
            a['L_list'] = [L[:repr_dim, :repr_dim].detach().cpu() for L in L_list]
        regressor = sindy.SINDyRegression(**a).to(dev)              # main.py:77
        (idx,) = next(iter(loader))                                 # train.py:626-627: the permutation is drawn HERE
        return idx.to(dev), regressor

    truth = sindy_truth[args['task']]
    results = sweep.run_seed_sweep_batched(x_all, dx_all, truth, seeds, None, subsample=args['lbfgs_subsample'],
                                           lr_sindy=args['lr_sindy'], st_freq=args['st_freq'],
                                           threshold=args['threshold'], num_epochs=args['num_epochs'], draw=draw)
    out_dir = f'eval_results/{args["save_dir"]}'
    os.makedirs(out_dir, exist_ok=True)
    for r in results:                                               # main.py:128-138
        np.savez(f'{out_dir}/seed{r["seed"]}.npz', coefficients=r["coefficients"], correct_form=r["correct_form"],
                 mse=r["mse"], correct_form_all=r["correct_form_all"], mse_all=r["mse_all"])
    sweep.aggregate(results)
    if json_out:
        with open(json_out, "w") as f:
            json.dump([{k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in r.items()} for r in results], f)


if __name__ == "__main__":
    main()
