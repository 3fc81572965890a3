"""Tensor-core WSINDy kernel (SB_WSINDY_TC=1) against the CUDA-core batched kernel and the CPU oracle, then timing."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, numpy as np, torch
sys.path[:0] = [%r, %r]
from sindy_b200 import native
from oracle import sindy_oracle as O
mode = os.environ.get("SB_WSINDY_TC", "0")
for (d, p, ntr, T, nt) in ((2, 3, 37, 1000, 50), (3, 5, 13, 1500, 50), (2, 2, 16, 333, 50), (3, 3, 24, 2049, 64)):
    lib = native.Library(d, p)
    rng = np.random.default_rng(100 * d + p)
    x = rng.uniform(0.2, 1.2, (ntr, T, d)).astype(np.float32)
    G, b = native.wsindy_integrals(torch.from_numpy(x).cuda(), lib, 0.002, T * 0.002, nt)
    torch.cuda.synchronize()
    errs = []
    for r in (0, ntr // 2, ntr - 1):
        Go, bo = O.wsindy_integrals(x[r], 0.002, T * 0.002, p, n_test=nt)
        errs.append((np.abs(G[r].cpu().numpy() - Go).max() / np.abs(Go).max(), np.abs(b[r].cpu().numpy() - bo).max() / np.abs(bo).max()))
    print("tc=" + mode, (d, p, ntr, T, nt), "max rel err G/b:", max(e[0] for e in errs), max(e[1] for e in errs), flush=True)
for (d, p, ntr) in ((2, 3, 18944), (2, 3, 56832), (3, 5, 4736), (2, 2, 18944), (3, 3, 9472)):
    lib = native.Library(d, p)
    x = torch.rand(ntr, 8000, d, device="cuda") * 0.8 + 0.2
    native.wsindy_integrals(x, lib, 0.002, 16.0, 50); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): native.wsindy_integrals(x, lib, 0.002, 16.0, 50)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    fl = (lib.K - 1 - d) + 100 * (lib.K + d)
    print("tc=" + mode, f"d={d} p={p} K={lib.K} ntraj={ntr}: {ms:.3f} ms  {ntr*8000/ms/1e6:.2f} Gsamples/s  {fl*ntr*8000/ms/1e9:.1f} TFLOP/s (algorithmic)", flush=True)
    del x
''' % (ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200"))
for v in sys.argv[1:] or ["1", "0"]:
    try:
        subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SB_WSINDY_TC=v), timeout=240)
    except subprocess.TimeoutExpired:
        print("TIMEOUT with SB_WSINDY_TC=" + v)
