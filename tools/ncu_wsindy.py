"""One launch of the batched WSINDy kernel for an `ncu --set full` capture (tools, not product)."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
from sindy_b200 import native
d, p, ntr = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (2, 3, 8192)))
lib = native.Library(d, p)
x = torch.rand(ntr, 8000, d, device="cuda") * 0.8 + 0.2
for _ in range(2):
    native.wsindy_integrals(x, lib, 0.002, 16.0, 50)
torch.cuda.synchronize()
