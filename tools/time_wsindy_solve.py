"""WSINDyWrapper: constructor and `solve` (one thresholding step) for one Sel'kov-length trajectory (T = 8000, d = 2,
cubic library, 50 test functions) — the reference's CPU numbers for the same calls are in bench.py's extras."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
import sindy
T, dt = 8000, 0.002
t = torch.arange(T, dtype=torch.float32) * dt
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(T, 2, device="cuda", generator=g) * 0.8 + 0.2
for w in (0.05, 0.0):
    reg = sindy.SINDyRegression(2, 3, False, False, threshold=0.1, device="cuda", constrain_constant=True)
    for _ in range(2):
        ws = sindy.WSINDyWrapper(reg, t, T * dt, 50, device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        ws = sindy.WSINDyWrapper(reg, t, T * dt, 50, device="cuda")
    torch.cuda.synchronize(); ctor = (time.perf_counter() - t0) / 10 * 1e3
    for _ in range(3):
        reg.reset_mask(); ws.solve(x, w, 0.1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        reg.reset_mask(); ws.solve(x, w, 0.1)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 20 * 1e3
    t0 = time.perf_counter()
    for _ in range(20):
        ws.integrals(x)
    torch.cuda.synchronize(); ms_i = (time.perf_counter() - t0) / 20 * 1e3
    print(f"w_sindy_reg={w}: ctor {ctor:.2f} ms, solve {ms:.2f} ms (integrals kernel {ms_i:.3f} ms)", flush=True)
