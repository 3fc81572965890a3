"""Scratch timing: moment-Gram kernel, STLSQ data pass and the sym-reg step at N = 1e8."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
import torch
from sindy_b200 import native
from sindy_b200.dist import ShardedTrainStep

def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]

n = 10**8
g = torch.Generator(device="cuda").manual_seed(1)
for (d, p) in ((3, 5), (3, 3), (2, 3)):
    lib = native.Library(d, p)
    x = torch.rand(n, d, device="cuda", generator=g) * 2 - 1
    dx = torch.randn(n, d, device="cuda", generator=g)
    W = torch.randn(d, lib.K, device="cuda", generator=g)
    for flags, name in ((4, "GRAM"), (8, "B"), (12, "GRAM|B (STLSQ pass)"), (3, "LOSS|GRAD"), (15, "all")):
        out = torch.empty(lib.step_out_len(flags), dtype=torch.float64, device="cuda")
        ms = timeit(lambda: native.train_step(x, dx, W, lib, flags, out=out))
        print(f"d={d} p={p} K={lib.K} {name:22s} [{native.train_step_variant(lib, flags)}]: {ms:.3f} ms -> {n/ms/1e6:.1f} Gsamples/s")
    if (d, p) == (3, 5):
        so3 = torch.zeros(3, 3, 3)
        k = 0
        for i in range(3):
            for j in range(i):
                so3[k, i, j], so3[k, j, i] = 1, -1; k += 1
        mask = torch.ones_like(W)
        for graph in (False, True):
            st = ShardedTrainStep(lib, x, dx, sym_gens=list(so3), w_sym=0.1, use_graph=graph, sgd_lr=1e-3)
            st.step(W, mask, 0.0)
            ms = timeit(lambda: st.step(None, None, 0.0))
            print(f"  closure + so(3) sym-reg step (graph={graph}): {ms:.3f} ms -> {n/ms/1e6:.1f} Gsamples/s")
    del x, dx
