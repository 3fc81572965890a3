"""Timeline of the fused <3,5> kernel's phases from in-kernel %globaltimer stamps (sb_debug_trace): launch gap
between back-to-back launches, time to the first tile, sample loop, CTA reduction + ticket, last-block reduction,
epilogue. Usage: python tools/trace_fused.py [fit|closure] [n_samples]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
import torch
from sindy_b200 import native
mode = sys.argv[1] if len(sys.argv) > 1 else "fit"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 12_500_000
lib = native.Library(3, 5)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(n, 3, device="cuda", generator=g) * 2 - 1
dx = torch.randn(n, 3, device="cuda", generator=g)
xi = torch.randn(3, 56, device="cuda", generator=g); mask = torch.ones_like(xi)
pk = torch.empty(170, dtype=torch.float64, device="cuda"); ls = torch.empty((), device="cuda"); gr = torch.empty(3, 56, device="cuda")
state = native.fit_state(lib, "cuda")
R = 6
bufs = [torch.zeros(16 * 592, dtype=torch.int64, device="cuda") for _ in range(R)]
native.load_w(xi, mask, lib)
def f():
    if mode == "fit":
        native.fit_step(x, dx, xi, mask, lib, "adam", 1e-4, state=state, w_resident=True, packed=pk, loss=ls, grad=gr)
    else:
        native.closure(x, dx, xi, mask, lib, 0.0, packed=pk, loss=ls, grad=gr)
for _ in range(3): f()
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
gph = torch.cuda.CUDAGraph()
with torch.cuda.stream(side): f()
torch.cuda.current_stream().wait_stream(side)
with torch.cuda.graph(gph):
    for r in range(R):
        native.load().sb_debug_trace(bufs[r].data_ptr())
        f()
native.load().sb_debug_trace(None)
gph.replay(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); gph.replay(); b.record(); torch.cuda.synchronize()
print(f"{mode} n={n}: {1e3 * a.elapsed_time(b) / R:.2f} us per iteration (graph of {R})")
T = [t.view(-1, 16).cpu() for t in bufs]
prev_end = None
for r in range(1, R):
    t = T[r]; live = t[:, 0] > 0; t = t[live]
    t0 = int(t[:, 0].min())
    last = int(t[:, 6].argmax())
    us = lambda v: (int(v) - t0) / 1e3
    gap = (t0 - prev_end) / 1e3 if prev_end else float("nan")
    print(f"launch {r}: ctas={t.shape[0]} gap_from_prev_end={gap:6.2f} | entry max {us(t[:,0].max()):5.2f} | first tile min/max {us(t[:,1].min()):5.2f}/{us(t[:,1].max()):5.2f}"
          f" | loop end min/max {us(t[:,2].min()):7.2f}/{us(t[:,2].max()):7.2f} | ticket max {us(t[:,3].max()):7.2f}"
          f" | rows staged {us(t[last,8]):7.2f} | summed {us(t[last,9]):7.2f} | totals {us(t[last,4]):7.2f} | peer {us(t[last,5]):7.2f} | end {us(t[last,6]):7.2f}")
    prev_end = int(t[last, 6])
t = T[R - 1]; t = t[t[:, 0] > 0]; t0 = int(t[:, 0].min())
le = (t[:, 2] - t0).double() / 1e3
order = torch.argsort(le)
print("loop-end (us) by CTA, sorted: " + " ".join(f"{le[i]:.1f}@sm{int(t[i,7])}" for i in order.tolist()))
