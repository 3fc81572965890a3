import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
import numpy as np, torch
from sindy_b200 import native
g = np.load(os.path.join(ROOT, "tests/golden/wsindy.npz"))
traj = torch.as_tensor(g["traj"]).cuda(); dt = float(g["dt"]); t_max = float(g["t_max"])
lib = native.Library(2, 3)
def rel(a, b): return float((a.cpu().numpy() - b).__abs__().max() / np.abs(b).max())
G0, b0 = native.wsindy_integrals(traj, lib, dt, t_max, 50)
print("fresh  ", rel(G0, g["G"]), rel(b0, g["b"]))
which = sys.argv[1] if len(sys.argv) > 1 else "moments"
lib35 = native.Library(3, 5)
n = 3_000_017
x = torch.rand(n, 3, device="cuda") * 2 - 1; dx = torch.randn(n, 3, device="cuda"); W = torch.randn(3, 56, device="cuda")
if which == "moments":
    native.train_step(x, dx, None, lib35, 12)
elif which == "fused":
    native.train_step(x, dx, W, lib35, 3)
elif which == "generic":
    native.train_step(x[:5000], dx[:5000], None, lib35, 12)
elif which == "rollout":
    native.rollout(x[:100].double(), W.double() * 0.01, lib35, 0.01, 10, 1, "rk4", record_dx=True)
torch.cuda.synchronize()
G1, b1 = native.wsindy_integrals(traj, lib, dt, t_max, 50)
print("after", which, rel(G1, g["G"]), rel(b1, g["b"]), "bitwise same:", torch.equal(G0, G1))
