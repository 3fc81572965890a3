"""Rollout timing (BASELINE config: 1e6 ICs x 2000 RK4 steps, every 10th state stored), fp32 and fp64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
import torch
from sindy_b200 import native

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]

g = torch.Generator(device="cuda").manual_seed(2)
x0 = torch.rand(10**6, 3, device="cuda", generator=g) * 2 - 1
for (d, p) in ((3, 5), (3, 3), (2, 3)):
    lib = native.Library(d, p)
    K = lib.K
    Xi = torch.zeros(d, K, device="cuda")
    if d == 3:
        Xi[0, 1], Xi[0, 2], Xi[1, 1], Xi[1, 2], Xi[1, 6], Xi[2, 5], Xi[2, 3] = -10, 10, 2.8, -1, -1, 1, -8 / 3
    else:
        Xi[0, 1], Xi[0, 2], Xi[1, 1], Xi[1, 2] = -0.1, -1.0, 1.0, -0.1
    flops = 4 * ((K - 1 - d) + 2 * K * d) + 12 * d
    xs = x0[:, :d].contiguous()
    ms = timeit(lambda: native.rollout(xs, Xi, lib, 0.002, 2000, 10, "rk4"))
    print(f"d={d} p={p} f32 1e6 ICs x 2000 steps: {ms:.1f} ms -> {2e9/ms/1e6:.2f} G IC-steps/s, {flops*2e9/ms/1e9:.1f} TFLOP/s")
    xd, Xd = xs[:10**5].double(), Xi.double()
    ms = timeit(lambda: native.rollout(xd, Xd, lib, 0.002, 2000, 10, "rk4", record_dx=True))
    print(f"d={d} p={p} f64 1e5 ICs x 2000 steps (dx recorded): {ms:.1f} ms -> {2e8/ms/1e6:.2f} G IC-steps/s")
