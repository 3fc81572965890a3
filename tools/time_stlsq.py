"""solve_SINDy (mask reset + up to 5 STLSQ iterations) on Lorenz-form data: data pass vs solves."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
import sindy
from sindy_b200 import native
for n in (20000, 4_000_000, 100_000_000):
    lib = native.Library(3, 5)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(n, 3, device="cuda", generator=g) * 2 - 1
    Xi = torch.zeros(3, lib.K, device="cuda")
    Xi[0, 1], Xi[0, 2], Xi[1, 1], Xi[1, 2], Xi[1, 6], Xi[2, 5], Xi[2, 3] = -10, 10, 2.8, -1, -1, 1, -8 / 3
    dx = native.forward(x, Xi, lib) + 0.01 * torch.randn(n, 3, device="cuda", generator=g)
    reg = sindy.SINDyRegression(3, 5, False, False, threshold=0.1, device="cuda", constrain_constant=True)
    for _ in range(2): sindy.solve_SINDy(reg, x, dx, 0.0, 0.1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): sindy.solve_SINDy(reg, x, dx, 0.0, 0.1)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    stats = sindy.stlsq_statistics(reg, x, dx); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): sindy.stlsq_statistics(reg, x, dx)
    torch.cuda.synchronize(); ms_d = (time.perf_counter() - t0) / 5 * 1e3
    ok = bool(torch.equal(reg.mask.bool(), Xi != 0))
    print(f"n={n}: solve_SINDy {ms:.2f} ms (data pass {ms_d:.2f} ms), support recovered {ok}, "
          f"coef err {float(((reg.Xi.detach()*reg.mask) - Xi).abs().max() / 10):.1e}", flush=True)
    del x, dx
