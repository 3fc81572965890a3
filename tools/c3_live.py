"""Config 3 live: the reference alone on cuda:0 and the drop-in on cuda:0 from the same seed (same CUDA-generator initial
parameters), the drop-in with the autoencoder on the tensor cores and on cuBLAS fp32."""
import os, sys, tempfile, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import config_runs
REF = config_runs.find_reference()
with tempfile.TemporaryDirectory() as w:
    a, _ = config_runs.run_entry("C3", os.path.join(w, "ref"), REF, dropin=False, gpu=0, timeout=3000)
    b, _ = config_runs.run_entry("C3", os.path.join(w, "b200"), REF, dropin=True, gpu=0, timeout=3000)
    c, _ = config_runs.run_entry("C3", os.path.join(w, "b200c"), REF, dropin=True, gpu=0, timeout=3000, env={"SINDY_B200_AE_MLP": "0"})
s = np.abs(a["coefficients"]).max()
print("mask equal", np.array_equal(a["coefficients"] != 0, b["coefficients"] != 0), np.array_equal(a["coefficients"] != 0, c["coefficients"] != 0))
print("ref-gpu vs dropin(tc AE): %.2e   ref-gpu vs dropin(cuBLAS AE): %.2e   dropin tc vs cublas: %.2e" % (
    np.abs(a["coefficients"] - b["coefficients"]).max() / s, np.abs(a["coefficients"] - c["coefficients"]).max() / s,
    np.abs(b["coefficients"] - c["coefficients"]).max() / s))
print(a["coefficients"]); print(b["coefficients"]); print(c["coefficients"])
