"""`run.py --seeds 0-2 main.py --config <cfg>`: three seeds of a BASELINE cfg in ONE process against three separate
processes — wall clock, and that every seed's `eval_results/.../seed<k>.npz` is identical either way."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import config_runs  # noqa: E402

REF = config_runs.find_reference()
key = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = config_runs.CONFIGS[key]
env = config_runs._env({"SINDY_B200_INIT_RNG": "cpu"})
script = os.path.join(REF, cfg["script"])
base = ["--config", cfg["cfg"], "--gpu", "0"] + cfg["extra"]
with tempfile.TemporaryDirectory() as wa, tempfile.TemporaryDirectory() as wb:
    config_runs.prepare_workdir(wa, REF, key)
    config_runs.prepare_workdir(wb, REF, key)
    t0 = time.time()
    r = subprocess.run([sys.executable, config_runs.LAUNCHER, "--reference", REF, "--seeds", "0-2", script] + base,
                       cwd=wa, env=env, capture_output=True, text=True)
    t_loop = time.time() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    t0 = time.time()
    for k in range(3):
        r = subprocess.run([sys.executable, config_runs.LAUNCHER, "--reference", REF, script] + base + ["--seed", str(k)],
                           cwd=wb, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    t_sep = time.time() - t0
    same = []
    for k in range(3):
        a = np.load(os.path.join(wa, "eval_results", cfg["save_dir"], f"seed{k}.npz"))
        b = np.load(os.path.join(wb, "eval_results", cfg["save_dir"], f"seed{k}.npz"))
        same.append(bool(np.array_equal(a["coefficients"], b["coefficients"])))
        print(k, a["coefficients"].ravel()[:4], b["coefficients"].ravel()[:4])
print(f"{key}: 3 seeds in one process {t_loop:.1f} s, in three processes {t_sep:.1f} s; identical coefficients per seed: {same}")
