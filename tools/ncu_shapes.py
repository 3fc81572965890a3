"""Launches one kernel family on 1e8 (or argv[2]) samples for an ncu capture (tools, not product):
    python tools/ncu_shapes.py fused33 | fused23 | fused22 | fused35 | moments35 | forward35 | rollout35 | stlsq35"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
from sindy_b200 import native
what = sys.argv[1]
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10 ** 8
d, p = int(what[-2]), int(what[-1])
lib = native.Library(d, p)
g = torch.Generator(device="cuda").manual_seed(0)
reps = 3
if what.startswith("rollout"):
    x0 = torch.rand(10 ** 6, d, device="cuda", generator=g) * 2 - 1
    W = 0.1 * torch.randn(d, lib.K, device="cuda", generator=g)
    for _ in range(reps):
        native.rollout(x0, W, lib, 0.002, 200, 10, "rk4")
else:
    x = torch.rand(n, d, device="cuda", generator=g) * 2 - 1
    W = torch.randn(d, lib.K, device="cuda", generator=g)
    if what.startswith("fused"):
        dx = torch.randn(n, d, device="cuda", generator=g)
        out = torch.empty(lib.step_out_len(3), dtype=torch.float64, device="cuda")
        for _ in range(reps):
            native.train_step(x, dx, W, lib, 3, out=out)
    elif what.startswith("moments"):
        from sindy_b200 import symreg
        for _ in range(reps):
            symreg.gram(x, lib)
    elif what.startswith("stlsq"):
        dx = torch.randn(n, d, device="cuda", generator=g)
        out = torch.empty(lib.step_out_len(12), dtype=torch.float64, device="cuda")
        for _ in range(reps):
            native.train_step(x, dx, None, lib, 12, out=out)
    elif what.startswith("forward"):
        for _ in range(reps):
            native.forward(x, W, lib)
torch.cuda.synchronize()
print("ok", what)
