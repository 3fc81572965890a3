"""Forward h(x) = Θ(x)Wᵀ at N = 1e8: bulk-copy ring kernel against the grid-stride kernel (SB_FORWARD_RING)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
from sindy_b200 import native
n = 10**8
for (d, p) in ((3, 5), (3, 3), (2, 3)):
    lib = native.Library(d, p)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n, d, device="cuda", generator=g) * 2 - 1
    W = torch.randn(d, lib.K, device="cuda", generator=g)
    for ring in ("1", "0"):
        os.environ["SB_FORWARD_RING"] = ring
        for _ in range(3): native.forward(x, W, lib)
        torch.cuda.synchronize(); ts = []
        for _ in range(11):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); native.forward(x, W, lib); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ts.sort()
        fl = (lib.K - 1 - d) + 2 * lib.K * d
        print(f"forward ({d},{p}) ring={ring}: {ts[5]:.4f} ms  {8*d*n/ts[5]/1e6:.0f} GB/s  {fl*n/ts[5]/1e9:.1f} TFLOP/s", flush=True)
    del x
