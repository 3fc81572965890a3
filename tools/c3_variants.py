"""Config 3 (`lv/noise99_eq_isymreg.cfg`) end to end in several arithmetic variants on the same GPU, against the golden of
the reference's CPU run: how far do runs of the SAME algorithm land from each other when only fp32 summation order
changes? (LBFGS + a ratio-of-means regulariser through a random frozen ReLU network amplify rounding differences.)
    python tools/c3_variants.py"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import config_runs  # noqa: E402

REF = config_runs.find_reference()
want = np.load(os.path.join(ROOT, "tests", "golden", "configs.npz"))["C3_coefficients"]
runs = {}
for name, dropin, env in (("dropin, autoencoder on cuBLAS fp32", True, {"SINDY_B200_AE_MLP": "0"}),
                          ("dropin, autoencoder on the tensor cores (3xTF32, split accumulators)", True,
                           {"SINDY_B200_AE_MLP": "1", "SB_MLP_SPLIT_ACC": "1"}),
                          ("dropin, autoencoder on the tensor cores (3xTF32, one accumulator)", True,
                           {"SINDY_B200_AE_MLP": "1", "SB_MLP_SPLIT_ACC": "0"}),
                          ("reference alone on the GPU (initial parameters from the CUDA generator)", False, {}))[:int(os.environ.get("C3_VARIANTS", "3"))]:
    env = dict(env, SINDY_B200_INIT_RNG="cpu")
    with tempfile.TemporaryDirectory() as work:
        t0 = time.time()
        try:
            res, _ = config_runs.run_entry("C3", work, REF, dropin=dropin, gpu=0, env=env, timeout=3000)
        except Exception as e:  # noqa: BLE001
            print(name, "FAILED", str(e)[-500:])
            continue
        runs[name] = res["coefficients"]
        print(f"{name}: {time.time() - t0:.1f} s wall (process start, data, fit, evaluation)", flush=True)
scale = np.abs(want).max()
for a, ca in runs.items():
    print(f"{a}: vs CPU golden {np.abs(ca - want).max() / scale:.2e}, mask equal {np.array_equal(ca != 0, want != 0)}")
names = list(runs)
for i in range(len(names)):
    for j in range(i + 1, len(names)):
        print(f"  {names[i]}  <->  {names[j]}: {np.abs(runs[names[i]] - runs[names[j]]).max() / scale:.2e}")
print("golden\n", want)
for a, ca in runs.items():
    print(a, "\n", ca)
