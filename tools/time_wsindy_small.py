import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
from sindy_b200 import native
for (d, p) in ((2, 3), (3, 5), (2, 2), (3, 3)):
    lib = native.Library(d, p)
    for ntr in (8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
        x = torch.rand(ntr, 8000, d, device="cuda") * 0.8 + 0.2
        row = []
        for mode in ("1", "0", "s"):
            os.environ["SB_WSINDY_TC"] = mode
            native.wsindy_integrals(x, lib, 0.002, 16.0, 50); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): native.wsindy_integrals(x, lib, 0.002, 16.0, 50)
            b.record(); torch.cuda.synchronize()
            row.append(a.elapsed_time(b) / 5)
        print(f"K={lib.K} n_traj={ntr}: tensor-core {row[0]:.3f} ms | CUDA-core batched {row[1]:.3f} ms | per-test-function {row[2]:.3f} ms", flush=True)
