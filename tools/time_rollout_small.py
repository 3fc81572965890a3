"""Both fp32 RK4 rollout kernels (two / one initial condition per thread) at the shard sizes of a 1e6-IC rollout over 8, 4, 2, 1 GPUs."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
from sindy_b200 import native
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
g = torch.Generator(device="cuda").manual_seed(2)
lib = native.Library(3, 5); K = lib.K
Xi = torch.zeros(3, K, device="cuda")
Xi[0, 1], Xi[0, 2], Xi[1, 1], Xi[1, 2], Xi[1, 6], Xi[2, 5], Xi[2, 3] = -10, 10, 2.8, -1, -1, 1, -8 / 3
for n in (125000, 250000, 500000, 1000000):
    x0 = torch.rand(n, 3, device="cuda", generator=g) * 2 - 1
    out = []
    for multi in ("1", "0"):
        os.environ["SB_ROLLOUT_MULTI"] = multi
        out.append(timeit(lambda: native.rollout(x0, Xi, lib, 0.002, 2000, 10, "rk4")))
    print(f"n_ics={n}: two-IC kernel {out[0]:.3f} ms, one-IC kernel {out[1]:.3f} ms; ideal from 1e6: {58.37*n/1e6:.3f}")
