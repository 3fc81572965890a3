"""Timing of the tensor-core MLP chain (sb_mlp_gemm, `sindy_b200/mlp.py`) against cuBLAS through PyTorch on the same
GPU: one 512-wide layer, and the decoder value + JVP + transpose chain of config 3's autoencoder.
    python tools/time_mlp.py [rows ...]         (default 40000 1000000)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "symmetry-ode-discovery_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from sindy_b200 import mlp, native  # noqa: E402
from test_gpu_mlp import make_ae  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    rows = [int(float(v)) for v in sys.argv[1:]] or [40000, 1000000]
    lib = native.load()
    f = 512
    dev = torch.device("cuda")
    s = native._stream(dev)
    for m in rows:
        a = torch.randn(m, f, device=dev)
        w = torch.randn(f, f, device=dev) / f ** 0.5
        b = torch.randn(f, device=dev)
        pa, pc = mlp._Panel.from_rows(a), mlp._Panel(m, f, dev)
        pk = torch.empty(2 * f * f, device=dev)
        native._check(lib.sb_mlp_pack_weights(w.data_ptr(), f, f, 0, pk.data_ptr(), s), "pack")
        t_tc = timed(lambda: native._check(lib.sb_mlp_gemm(pa.ptr(), m, f, pk.data_ptr(), f, b.data_ptr(), None, 1,
                                                           pc.ptr(), s), "gemm"))
        torch.backends.cuda.matmul.allow_tf32 = False
        t_f32 = timed(lambda: torch.relu(torch.nn.functional.linear(a, w, b)))
        torch.backends.cuda.matmul.allow_tf32 = True
        t_tf32 = timed(lambda: torch.relu(torch.nn.functional.linear(a, w, b)))
        torch.backends.cuda.matmul.allow_tf32 = False
        ref = torch.relu(a.double() @ w.double().t() + b.double())
        e_tc = float((pc.to_rows().double() - ref).abs().max() / ref.abs().max())
        e_f32 = float((torch.relu(torch.nn.functional.linear(a, w, b)).double() - ref).abs().max() / ref.abs().max())
        fl = 2.0 * m * f * f
        print(f"layer m={m}: sb_mlp_gemm {t_tc:.4f} ms = {fl / t_tc / 1e9:.1f} TFLOP/s fp32-equivalent "
              f"({3 * fl / t_tc / 1e9:.0f} TF/s tf32 issued), err {e_tc:.1e} | cuBLAS fp32 {t_f32:.4f} ms "
              f"({fl / t_f32 / 1e9:.1f} TF/s, err {e_f32:.1e}) | cuBLAS tf32x1 {t_tf32:.4f} ms", flush=True)
        del a, pa, pc, ref
        # decoder: value + JVP + gradient w.r.t. the tangent, as symmreg_i uses it
        batch = m // 2
        ae = make_ae(seed=0)
        fast = mlp.accelerate(ae)
        z = torch.randn(batch, 2, 2, device=dev)
        v = torch.randn(batch, 2, 2, device=dev)
        c = torch.randn(batch, 2, 2, device=dev)

        def ours():
            vg = v.clone().requires_grad_(True)
            (fast[1].value_and_jvp(z, vg)[1] * c).sum().backward()
            return vg.grad

        def theirs():
            vg = v.clone().requires_grad_(True)
            jt = torch.autograd.functional.jvp(ae.decoder, z, vg, create_graph=True)[1]
            (jt * c).sum().backward()
            return vg.grad

        if m <= 200000:
            t_o, t_t = timed(ours, 10, 2), timed(theirs, 10, 2)
            d = (ours() - theirs()).abs().reshape(batch, -1).max(dim=1).values / theirs().abs().max()
            print(f"decoder JVP + backward, batch {batch} x 2 comps: tensor-core chain {t_o:.3f} ms | PyTorch double-vjp "
                  f"{t_t:.3f} ms | {t_t / t_o:.2f}x | rel diff: median {float(d.median()):.1e}, rows above 1e-5: "
                  f"{int((d > 1e-5).sum())} (ReLU units within rounding of zero flip)", flush=True)


if __name__ == "__main__":
    main()
