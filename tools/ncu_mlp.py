"""Three launches of the tensor-core MLP layer (sb_mlp_gemm, 40 000 x 512 x 512, ReLU epilogue) for an `ncu --set full`
capture (tools, not product)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
from sindy_b200 import mlp, native  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
f = 512
lib = native.load()
a = torch.randn(m, f, device="cuda")
w = torch.randn(f, f, device="cuda") / f ** 0.5
b = torch.randn(f, device="cuda")
pa, pc = mlp._Panel.from_rows(a), mlp._Panel(m, f, a.device)
pk = torch.empty(2 * f * f, device="cuda")
s = native._stream(a.device)
native._check(lib.sb_mlp_pack_weights(w.data_ptr(), f, f, 0, pk.data_ptr(), s), "pack")
for _ in range(3):
    native._check(lib.sb_mlp_gemm(pa.ptr(), m, f, pk.data_ptr(), f, b.data_ptr(), None, 1, pc.ptr(), s), "gemm")
torch.cuda.synchronize()
