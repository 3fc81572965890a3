"""Scratch timing: rollout, specialised forward/backward at N = 1e8, reference-style closure through autograd."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
import torch
from sindy_b200 import native
import sindy

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]

lib = native.Library(3, 5)
g = torch.Generator(device="cuda").manual_seed(2)
x0 = torch.rand(10**6, 3, device="cuda", generator=g) * 2 - 1
Xi = torch.zeros(3, 56, device="cuda")
Xi[0, 1], Xi[0, 2] = -10, 10
Xi[1, 1], Xi[1, 2], Xi[1, 6] = 2.8, -1, -1
Xi[2, 5], Xi[2, 3] = 1, -8 / 3
ms = timeit(lambda: native.rollout(x0, Xi, lib, 0.002, 2000, 10, "rk4"), 3)
print(f"rollout f32 1e6 ICs x 2000 steps: {ms:.1f} ms -> {2e9/ms/1e6:.2f} G IC-steps/s, {1588*2e9/ms/1e9:.1f} TFLOP/s")
n = 10**8
x = torch.rand(n, 3, device="cuda", generator=g) * 2 - 1
W = torch.randn(3, 56, device="cuda", generator=g)
ms = timeit(lambda: native.forward(x, W, lib))
print(f"forward  n=1e8: {ms:.3f} ms -> {n/ms/1e6:.1f} Gsamples/s, {24*n/ms/1e6:.0f} GB/s")
gy = torch.randn(n, 3, device="cuda", generator=g)
ms = timeit(lambda: native.backward(x, gy, W, lib, True, False))
print(f"backward gw n=1e8: {ms:.3f} ms -> {n/ms/1e6:.1f} Gsamples/s")
del gy
# the reference's closure as written (regressor(x), MSELoss, backward) on the drop-in module
reg = sindy.SINDyRegression(3, 5, False, False, threshold=0.1, device="cuda", constrain_constant=True)
dx = native.forward(x, Xi, lib)
def ref_style():
    reg.zero_grad()
    loss = torch.nn.MSELoss()(reg(x), dx)
    loss.backward()
ms = timeit(ref_style, 3)
print(f"reference-style closure (forward + MSELoss + backward through autograd) n=1e8: {ms:.2f} ms -> {n/ms/1e6:.2f} Gsamples/s")
def fused_style():
    reg.zero_grad()
    loss = reg.mse_loss(x, dx)
    loss.backward()
ms = timeit(fused_style, 3)
print(f"fused closure via module (mse_loss + backward) n=1e8: {ms:.2f} ms -> {n/ms/1e6:.2f} Gsamples/s")
