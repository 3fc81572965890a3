"""A/B timing of the small-library fused kernels (SB_FUSED_VARIANT_SMALL), one subprocess per variant."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, torch
sys.path[:0] = [%r, %r]
from sindy_b200 import native
n = 10**8
shapes = [(2, 2, 0), (2, 3, 0), (3, 3, 0)] if not os.environ.get('AB_MORE') else [(3, 2, 0), (2, 2, 1)]
for (d, p, ex) in shapes:
    lib = native.Library(d, p, False, bool(ex))
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n, d, device="cuda", generator=g) * 2 - 1
    dx = torch.randn(n, d, device="cuda", generator=g)
    W = torch.randn(d, lib.K, device="cuda", generator=g)
    out = torch.empty(lib.step_out_len(3), dtype=torch.float64, device="cuda")
    for _ in range(3): native.train_step(x, dx, W, lib, 3, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(15):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); native.train_step(x, dx, W, lib, 3, out=out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print("variant", os.environ.get("SB_FUSED_VARIANT_SMALL"), (d, p, ex), "median %%.4f ms best %%.4f ms -> %%.0f GB/s" %% (ts[7], ts[0], 8 * d * n / ts[7] / 1e6))
    del x, dx
''' % (ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200"))
for v in sys.argv[1:] or ["5", "13", "37", "45"]:
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SB_FUSED_VARIANT_SMALL=v))
