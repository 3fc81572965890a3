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

import numpy as np  # noqa: E402

import config_runs  # noqa: E402

REF = os.environ.get("SINDY_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "configs.npz")
LOG = os.path.join(ROOT, "tests", "golden", "configs_log.txt")


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
        tail = "\n".join(stdout.strip().splitlines()[-14:])
        logs.append(f"=== {key}: {config_runs.CONFIGS[key]['cfg']} (reference on CPU, {dt:.0f} s) ===\n{tail}\n")
        print(logs[-1], flush=True)
    np.savez_compressed(OUT, **store)
    with open(LOG, "a") as f:
        f.write("\n".join(logs))
    print(f"wrote {OUT}")


if __name__ == "__main__":
    main(sys.argv[1:] or ["C1", "C2", "C3", "C4", "C1e", "C1w", "C2s"])
