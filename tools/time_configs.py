"""Wall-clock of the four BASELINE configs' entry points on the test fixtures (tests/golden/data): the reference alone on
the host CPU (`--gpu -1`), the reference alone on cuda:0, and the same unmodified script through the drop-in launcher on
cuda:0. Whole process each time (interpreter start, imports, data, fit, evaluation)."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import config_runs  # noqa: E402

REF = config_runs.find_reference()
keys = sys.argv[1:] or ["C1", "C2", "C3", "C4"]
for key in keys:
    row = []
    for name, dropin, gpu in (("reference, CPU", False, -1), ("reference, cuda:0", False, 0), ("drop-in, cuda:0", True, 0)):
        with tempfile.TemporaryDirectory() as work:
            config_runs.prepare_workdir(work, REF, key)        # fixtures / stand-in checkpoint outside the timed part
            t0 = time.time()
            try:
                config_runs.run_entry(key, work, REF, dropin=dropin, gpu=gpu, timeout=1500)
                row.append(f"{name} {time.time() - t0:.1f} s")
            except Exception as e:  # noqa: BLE001
                row.append(f"{name} FAILED ({str(e)[-120:]!r})")
    print(f"{key} ({config_runs.CONFIGS[key]['cfg']}): " + " | ".join(row), flush=True)
