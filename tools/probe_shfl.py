"""Design probe: shuffle throughput next to the FMA pipe (sb_fp32_peak variants 3 and 4)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
from sindy_b200 import native
import torch
torch.cuda.init()
for v, name in ((2, "ffma2_const"), (3, "shfl only"), (4, "ffma2_const + 1 shfl per 4")):
    t = native.fp32_peak(v, 4096)
    lanes_per_clk_sm = t * 1e12 / 2 / (148 * 1.965e9)
    print(f"variant {v} ({name}): {t:.2f} T(2 x lane-op)/s = {lanes_per_clk_sm:.1f} lane-ops/clk/SM")
