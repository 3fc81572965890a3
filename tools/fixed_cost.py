"""Fixed cost of the fused <3,5> launch: time vs number of tiles per CTA (n = 148*1024*t)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
import torch
from sindy_b200 import native
lib = native.Library(3, 5)
g = torch.Generator(device="cuda").manual_seed(1)
nmax = 148 * 1024 * 128
x = torch.rand(nmax, 3, device="cuda", generator=g) * 2 - 1
dx = torch.randn(nmax, 3, device="cuda", generator=g)
W = torch.randn(3, 56, device="cuda", generator=g)
mask = torch.ones_like(W)
pk = torch.empty(170, dtype=torch.float64, device="cuda"); ls = torch.empty((), device="cuda"); gr = torch.empty(3, 56, device="cuda")
state = native.fit_state(lib, "cuda")
xi = W.clone()
mode = sys.argv[1] if len(sys.argv) > 1 else "closure"
for t in (1, 82, 128):
    n = 148 * 1024 * t
    if mode == "fit":   # one launch per iteration: closure + Adam + next W
        native.load_w(xi, mask, lib)
        f = lambda: native.fit_step(x[:n], dx[:n], xi, mask, lib, "adam", 1e-4, state=state, w_resident=True, packed=pk, loss=ls, grad=gr)
    else:
        f = lambda: native.closure(x[:n], dx[:n], W, mask, lib, 0.0, packed=pk, loss=ls, grad=gr)
    for _ in range(3): f()
    torch.cuda.synchronize()
    gph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side): f()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(gph):
        for _ in range(20): f()
    gph.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): gph.replay()
    b.record(); torch.cuda.synchronize()
    us = 1e3 * a.elapsed_time(b) / 100
    print(f"tiles/CTA {t:4d} n={n:10d}: {us:8.2f} us per {mode} (graph of 20, back to back)  per-tile {us/t:6.3f} us")
