"""Generates the data sets of BASELINE configs C1-C4 with the UNMODIFIED reference generators and stores them as
the `.pt` files `dataset.ODEDataset` loads (`dataset.py:170-200`), under tests/golden/data/.

TEST INFRASTRUCTURE. Run in the build container only (needs /root/reference):

    python oracle/gen_config_data.py [dosc growth lv selkov]

The reference's generators (`python -m data_utils.<ode>`, README "Option 2") never seed NumPy
(`damped_oscillator.py:28-44`), so every file is produced by `runpy.run_module(..., run_name='__main__')` of the
reference module after `np.random.seed(seed)`: same code path, reproducible draws. Sizes:

  * dosc   (C1): README sizes, train (50, 100, 2) / val (10, 100, 2) — 10^4 RK4 steps, GP-smoothed, subsampled by 100
  * growth (C2): README sizes, train (100, 100, 2) / val (20, 100, 2)
  * lv     (C3): REDUCED from (200, 10^4, 2) = 16 MB per file to (16, 2000, 2) so that it can live in git; same
                 generator, noise 0.99, GP smoothing
  * selkov (C4): REDUCED from 10 trajectories to 4 (main_wsindy.py uses ONE random 80 % window of one trajectory);
                 the full 10^4 steps are kept because T = 8000 is what the config's weak form integrates over
"""
import os
import runpy
import sys
import time

REF = os.environ.get("SINDY_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "data")
sys.path.insert(0, REF)

import numpy as np  # noqa: E402

JOBS = {
    # name: (module, [(argv, seed), ...])
    "dosc": ("data_utils.damped_oscillator", [
        (["--n_ics", "50", "--noise", "0.2", "--smoothing", "gp"], 1001),
        (["--n_ics", "10", "--noise", "0.2", "--smoothing", "gp", "--save_name", "val"], 1002)]),
    "growth": ("data_utils.growth", [
        (["--n_ics", "100", "--noise", "0.05", "--smoothing", "gp"], 2001),
        (["--n_ics", "20", "--noise", "0.05", "--smoothing", "gp", "--save_name", "val"], 2002)]),
    "lv": ("data_utils.lotka", [
        (["--n_ics", "16", "--num_steps", "2000", "--noise", "0.99", "--smoothing", "gp"], 3001),
        (["--n_ics", "4", "--num_steps", "2000", "--noise", "0.99", "--smoothing", "gp", "--save_name", "val"], 3002)]),
    "selkov": ("data_utils.selkov", [
        (["--n_ics", "4", "--noise", "0.2", "--smoothing", "gp"], 4001),
        (["--n_ics", "2", "--noise", "0.2", "--smoothing", "gp", "--save_name", "val"], 4002)]),
}


def main(names):
    os.makedirs(OUT, exist_ok=True)
    for name in names:
        module, runs = JOBS[name]
        for argv, seed in runs:
            t0 = time.time()
            np.random.seed(seed)
            sys.argv = [module] + argv + ["--save_dir", OUT]
            mod = runpy.run_module(module, run_name="__main__")
            assert mod["__file__"].startswith(REF), mod["__file__"]
            print(f"{module} {' '.join(argv)} seed={seed}: {time.time() - t0:.0f} s", flush=True)
    for f in sorted(os.listdir(OUT)):
        print(f"{f}: {os.path.getsize(os.path.join(OUT, f)) / 1024:.0f} KiB")


if __name__ == "__main__":
    main(sys.argv[1:] or list(JOBS))
