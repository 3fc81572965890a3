"""Scratch GPU probe: FP32 peak variants and fused train-step timings (not the bench contract)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
import torch
from sindy_b200 import native

torch.cuda.set_device(0)
print(torch.cuda.get_device_name(0))
for v, name in ((0, "FFMA"), (1, "FFMA2"), (2, "FFMA2+const")):
    print(f"fp32 peak {name}: {native.fp32_peak(v, 8192):.1f} TFLOP/s")

def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]

for (d, p, n) in ((3, 5, 10**7), (3, 5, 10**8), (3, 3, 10**8), (2, 2, 10**8), (2, 3, 10**8)):
    lib = native.Library(d, p)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n, d, device="cuda", generator=g) * 2 - 1
    dx = torch.randn(n, d, device="cuda", generator=g)
    W = torch.randn(d, lib.K, device="cuda", generator=g)
    out = torch.empty(lib.step_out_len(3), dtype=torch.float64, device="cuda")
    med, best = timeit(lambda: native.train_step(x, dx, W, lib, 3, out=out))
    print(f"train_step d={d} p={p} K={lib.K} n={n:.0e} [{native.train_step_variant(lib, 3)}]: median {med:.3f} ms best {best:.3f} ms "
          f"-> {n / med / 1e6:.2f} Gsamples/s, {8 * d * n / med / 1e6:.0f} GB/s")
    if (d, p) == (3, 5) and n == 10**7:
        fb = 12
        outb = torch.empty(lib.step_out_len(fb), dtype=torch.float64, device="cuda")
        med, best = timeit(lambda: native.train_step(x, dx, None, lib, fb, out=outb), reps=3)
        print(f"  generic GRAM|B: median {med:.3f} ms")
    del x, dx
# rollout
lib = native.Library(3, 5)
g = torch.Generator(device="cuda").manual_seed(2)
n_ics = 10**6
x0 = torch.rand(n_ics, 3, device="cuda", generator=g) * 2 - 1
Xi = torch.zeros(3, 56, device="cuda")
Xi[0, 1], Xi[0, 2] = -10, 10
Xi[1, 1], Xi[1, 2], Xi[1, 6] = 2.8, -1, -1
Xi[2, 5], Xi[2, 3] = 1, -8 / 3
for dt_, steps in ((torch.float32, 2000), (torch.float64, 200)):
    med, best = timeit(lambda: native.rollout(x0.to(dt_), Xi.to(dt_), lib, 0.002, steps, 10, "rk4", record_dx=False), reps=3)
    print(f"rollout {dt_} 1e6 ICs x {steps} steps: {med:.1f} ms -> {n_ics * steps / med / 1e6:.2f} G IC-steps/s")
