"""Goldens of BASELINE configs C1-C4: the reference's UNMODIFIED `main.py` / `main_wsindy.py` run with the cfg files
`run_configs/{dosc/noise20_sindy,growth/noise05_esindy,lv/noise99_eq_isymreg,selkov/noise20_eq_wsindy}.cfg` on the CPU
(`--gpu -1`), on the data sets of tests/golden/data (oracle/gen_config_data.py), seed 0.

TEST INFRASTRUCTURE. Build container only (needs /root/reference):

    python oracle/gen_config_golden.py [C1 C2 C3 C4 ...]      -> tests/golden/configs.npz  (+ configs_log.txt)

Stored per config: the masked coefficient matrix, `correct_form` and `mse` of `evaluation/eval_eq.py:7-34`, and the raw
parameters of `saved_models/<save_dir>/regressor.pt`. `tests/test_gpu_configs.py` runs the same entry points through
the drop-in launcher on the B200 and compares.
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import config_runs  # noqa: E402

REF = os.environ.get("SINDY_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "configs.npz")
LOG = os.path.join(ROOT, "tests", "golden", "configs_log.txt")


def wsindy_f64_sequence(key, seed=0):
    """The weak-form STLSQ sequence of `main_wsindy.py:25-46` + `train.py:855-869` for the cfg's data set with the
    least-squares solves in float64 (oracle.wsindy_one_step(solve_dtype=float64)): same seed, same random trajectory
    and 80 % window (`main_wsindy.py:36-39`), same threshold / ridge / iteration cap as the cfg."""
    import torch
    from oracle import sindy_oracle as O
    cfg = config_runs.CONFIGS[key]
    toks = open(os.path.join(REF, "run_configs", cfg["cfg"])).read().split()
    opt = {toks[i].lstrip("-"): toks[i + 1] for i in range(len(toks) - 1) if toks[i].startswith("--") and not toks[i + 1].startswith("--")}
    poly, thr, w, iters = int(opt.get("poly_order", 2)), float(opt["threshold"]), float(opt["w_sindy_reg"]), int(opt["num_epochs"])
    dt = {"lv": 0.002, "selkov": 0.002, "dosc": 0.2, "growth": 0.02}[cfg["ode"]]          # dataset.py:161-167
    noise = float(opt["noise"])
    x = torch.load(os.path.join(config_runs.DATA, f"{cfg['ode']}-train-noise{int(100 * noise):02d}-gp-x.pt")).float()
    n_ics, n_steps, _ = x.shape
    torch.manual_seed(seed)
    np.random.seed(seed)
    start = np.random.randint(0, n_steps - int(0.8 * n_steps))
    traj = np.random.randint(0, n_ics)
    win = x[traj, start:start + int(0.8 * n_steps)].numpy()
    T = int(0.8 * n_steps)
    mask = np.ones((2, O.term_count(2, poly)), dtype=np.float32)
    xis, masks = [], []
    for _ in range(iters):
        Xi, mask, conv = O.wsindy_one_step(win, mask, w, thr, dt, T * dt, poly, solve_dtype=torch.float64)
        xis.append(Xi)
        masks.append(mask.copy())
        if conv:
            break
    return np.stack(xis), np.stack(masks), (Xi * mask)


def main(keys):
    store = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    logs = []
    for key in keys:
        with tempfile.TemporaryDirectory() as work:
            t0 = time.time()
            res, stdout = config_runs.run_entry(key, work, REF, dropin=False, gpu=-1)
            dt = time.time() - t0
        for k, v in res.items():
            store[f"{key}_{k}"] = v
        if config_runs.CONFIGS[key]["script"] == "main_wsindy.py":
            xis, masks, final = wsindy_f64_sequence(key)
            store[f"{key}_f64_xis"], store[f"{key}_f64_masks"], store[f"{key}_f64_coefficients"] = xis, masks, final
            if key == "C4":   # evidence that the reference's own fp32 result is not reproducible for this cfg
                for th in ("1", "8"):
                    with tempfile.TemporaryDirectory() as w2:
                        r2, _ = config_runs.run_entry(key, w2, REF, dropin=False, gpu=-1,
                                                      env={"OMP_NUM_THREADS": th, "MKL_NUM_THREADS": th})
                    store[f"{key}_reference_threads{th}_coefficients"] = r2["coefficients"]
        tail = "\n".join(stdout.strip().splitlines()[-14:])
        logs.append(f"=== {key}: {config_runs.CONFIGS[key]['cfg']} (reference on CPU, {dt:.0f} s) ===\n{tail}\n")
        print(logs[-1], flush=True)
    np.savez_compressed(OUT, **store)
    with open(LOG, "a") as f:
        f.write("\n".join(logs))
    print(f"wrote {OUT}")


if __name__ == "__main__":
    main(sys.argv[1:] or ["C1", "C2", "C3", "C4", "C1e", "C1w", "C2s"])
