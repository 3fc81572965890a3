"""GPU parity of the tensor-core MLP chain (`sindy_b200/mlp.py`, `csrc/sb_mlp.cu`) against plain PyTorch
fp32 / fp64 evaluations of the same frozen network — the reference's `AutoEncoder` encoder / decoder structure
(`autoencoder.py:38-66`: Linear + BatchNorm1d(eval) + ReLU blocks, orthogonal last encoder layer, 512 wide, 5 layers).
Tolerance: with both halves of the split rounded to nearest the products carry ~22 mantissa bits; what remains is the
tensor core's fp32 accumulator, which truncates on each of the 192 accumulations per output: measured 3.3e-6 of the
largest output for one 512-deep layer against fp64 (cuBLAS fp32: 8e-7; one-pass TF32: 1e-3). Bounds: 1e-5 per layer,
2e-5 through the five-layer networks and their derivatives."""
import numpy as np
import pytest
import torch
import torch.nn as nn
from torch.nn.utils.parametrizations import orthogonal

pytestmark = pytest.mark.gpu


class Reshape(nn.Module):          # same name and behaviour as the reference's `model.py:8-14`
    def __init__(self, *args):
        super().__init__()
        self.shape = args

    def forward(self, x):
        return x.reshape(self.shape)


class RefShapedAutoEncoder(nn.Module):
    """Module tree of `autoencoder.py:38-66` (ae_arch 'mlp', batch_norm, ortho_ae, activation ReLU)."""

    def __init__(self, input_dim=2, hidden=512, latent=2, n_layers=5, n_comps=2):
        super().__init__()
        bn = lambda f: [Reshape(-1, f), nn.BatchNorm1d(f), Reshape(-1, n_comps, f)]
        self.encoder = nn.Sequential(
            nn.Linear(input_dim, hidden), *bn(hidden), nn.ReLU(),
            *[nn.Sequential(nn.Linear(hidden, hidden), *bn(hidden), nn.ReLU()) for _ in range(n_layers - 1)],
            orthogonal(nn.Linear(hidden, latent)), *bn(latent))
        self.decoder = nn.Sequential(
            nn.Linear(latent, hidden), nn.ReLU(),
            *[nn.Sequential(nn.Linear(hidden, hidden), nn.ReLU()) for _ in range(n_layers - 1)],
            nn.Linear(hidden, input_dim))

    def encode(self, x):
        return self.encoder(x)

    def decode(self, z):
        return self.decoder(z)

    @property
    def device(self):
        return next(self.parameters()).device


def make_ae(seed=0, **kw):
    torch.manual_seed(seed)
    ae = RefShapedAutoEncoder(**kw)
    with torch.no_grad():
        for m in ae.modules():
            if isinstance(m, nn.BatchNorm1d):
                m.running_mean.copy_(0.1 * torch.randn_like(m.running_mean))
                m.running_var.copy_(0.5 + torch.rand_like(m.running_var))
                m.weight.copy_(1.0 + 0.2 * torch.randn_like(m.weight))
                m.bias.copy_(0.1 * torch.randn_like(m.bias))
    ae.eval()
    for p in ae.parameters():
        p.requires_grad_(False)
    return ae.cuda()


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_rows(a, b, outliers=1e-2):
    """Error of a DERIVATIVE of a ReLU network: a unit whose pre-activation is within rounding of zero flips its mask
    and changes that row's Jacobian by O(1/sqrt(width)) in any fp32 evaluation (cuBLAS against fp64 as well). Rows are
    therefore judged individually: the given fraction may be off, the rest is returned as a max-norm relative error."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = (a - b).abs().reshape(a.shape[0], -1).max(dim=1).values / b.abs().max().clamp_min(1e-30)
    k = 0 if err.numel() < 50 else max(2, int(outliers * err.numel()))
    return float(err.sort().values[err.numel() - 1 - k])


@pytest.fixture(scope="module")
def mlp():
    from sindy_b200 import mlp as M
    assert torch.cuda.is_available()
    return M


@pytest.mark.parametrize("m", [1, 127, 128, 300, 4097])
def test_panel_format_round_trip_is_exact(mlp, m):
    x = torch.randn(m, 512, device="cuda") * torch.logspace(-6, 6, 512, device="cuda")
    p = mlp._Panel.from_rows(x)
    assert bool(((p.to_rows() - x).abs() <= x.abs() * 2.0 ** -23).all())   # hi + lo: two tf32 roundings, ≤ 2⁻²⁴ relative


@pytest.mark.parametrize("variant", ["default", "no_narrow", "pair", "one_accumulator"])
@pytest.mark.parametrize("m,mode", [(128, 0), (300, 1), (1000, 2), (2500, 1), (5100, 0), (40000, 1)])
def test_wide_layer_kernel_vs_fp64(mlp, m, mode, variant, monkeypatch):
    """One sb_mlp_gemm launch: C = epilogue(A·Wᵀ (+ b)) for the three epilogues, ragged last tile included. The row
    counts exercise the last partial round as 4 narrow pieces per tile (128, 300, 1000, 40000), as 2 (2500) and whole
    (5100); variants: all rounds whole tiles, the CTA-pair kernel (cta_group::2), one accumulator (A/B record)."""
    from sindy_b200 import native
    if variant == "no_narrow":
        monkeypatch.setenv("SB_MLP_NARROW", "0")
    elif variant == "pair":
        monkeypatch.setenv("SB_MLP_PAIR", "1")
    elif variant == "one_accumulator":
        if m != 40000:
            pytest.skip("A/B variant: one size")
        monkeypatch.setenv("SB_MLP_SPLIT_ACC", "0")
    g = torch.Generator(device="cuda").manual_seed(m + mode)
    f = 512
    a = torch.randn(m, f, device="cuda", generator=g)
    w = torch.randn(f, f, device="cuda", generator=g) / f ** 0.5
    b = torch.randn(f, device="cuda", generator=g)
    r = torch.randn(m, f, device="cuda", generator=g)
    lib = native.load()
    pa, pr, pc = mlp._Panel.from_rows(a), mlp._Panel.from_rows(r), mlp._Panel(m, f, a.device)
    pk = torch.empty(2 * f * f, device="cuda")
    s = native._stream(a.device)
    native._check(lib.sb_mlp_pack_weights(w.data_ptr(), f, f, 0, pk.data_ptr(), s), "pack")
    native._check(lib.sb_mlp_gemm(pa.ptr(), m, f, pk.data_ptr(), f, b.data_ptr() if mode != 2 else None,
                                  pr.ptr() if mode == 2 else None, mode, pc.ptr(), s), "gemm")
    ref = a.double() @ w.double().t()
    if mode != 2:
        ref = ref + b.double()
    if mode == 1:
        ref = ref.clamp_min(0)
    if mode == 2:
        ref = ref * (r > 0)
    got = pc.to_rows()
    assert rel(got, ref) < 1e-5
    # the transposed packing: C = A·W
    native._check(lib.sb_mlp_pack_weights(w.data_ptr(), f, f, 1, pk.data_ptr(), s), "pack")
    native._check(lib.sb_mlp_gemm(pa.ptr(), m, f, pk.data_ptr(), f, None, None, 0, pc.ptr(), s), "gemm")
    assert rel(pc.to_rows(), a.double() @ w.double()) < 1e-5


@pytest.mark.parametrize("batch", [1, 77, 20000])
def test_frozen_mlp_value_jvp_and_gradients_vs_torch(mlp, batch):
    ae = make_ae(seed=3)
    pair = mlp.accelerate(ae)
    assert pair is not None, "the reference-shaped autoencoder must be recognised"
    enc, dec = pair
    assert mlp.accelerate(ae) is pair             # cached on the module
    g = torch.Generator(device="cuda").manual_seed(batch)
    x = torch.randn(batch, 2, 2, device="cuda", generator=g)
    t = torch.randn(batch, 2, 2, device="cuda", generator=g)
    ae64 = make_ae(seed=3).double()
    for fast, mod, mod64 in ((enc, ae.encoder, ae64.encoder), (dec, ae.decoder, ae64.decoder)):
        y = fast.value(x)
        assert y.shape == mod(x).shape
        assert rel(y, mod64(x.double())) < 2e-5
        y2, jt = fast.value_and_jvp(x, t)
        _, jt64 = torch.autograd.functional.jvp(mod64, x.double(), t.double())
        assert torch.equal(y2, y) and rel_rows(jt, jt64) < 2e-5
        # gradients with respect to x and t of a scalar of (value, tangent): the transpose chains
        xg, tg = x.clone().requires_grad_(True), t.clone().requires_grad_(True)
        cy, cj = torch.randn_like(y), torch.randn_like(jt)
        yy, jj = fast.value_and_jvp(xg, tg)
        ((yy * cy).sum() + (jj * cj).sum()).backward()
        x64, t64 = x.double().requires_grad_(True), t.double().requires_grad_(True)
        y64, j64 = torch.autograd.functional.jvp(mod64, x64, t64, create_graph=True)
        ((y64 * cy.double()).sum() + (j64 * cj.double()).sum()).backward()
        assert rel_rows(xg.grad, x64.grad) < 2e-5 and rel_rows(tg.grad, t64.grad) < 2e-5
        xg = x.clone().requires_grad_(True)
        (fast.value(xg) * cy).sum().backward()
        x64 = x.double().requires_grad_(True)
        (mod64(x64) * cy.double()).sum().backward()
        assert rel_rows(xg.grad, x64.grad) < 2e-5


def test_modules_outside_the_supported_form_keep_the_pytorch_path(mlp):
    ae = make_ae(seed=1)
    ae.train()
    assert mlp.accelerate(ae) is None                               # BatchNorm would use batch statistics
    ae.eval()
    next(ae.decoder.parameters()).requires_grad_(True)
    assert mlp.accelerate(ae) is None                               # trainable: autograd must see the parameters
    next(ae.decoder.parameters()).requires_grad_(False)
    assert mlp.accelerate(ae) is not None
    ae.decoder[1] = nn.Tanh()
    assert mlp.accelerate(ae) is None                               # not piecewise linear
    small = make_ae(seed=1, hidden=96)
    assert mlp.accelerate(small) is None                            # width not a multiple of 256


def test_weight_update_invalidates_the_cached_chain(mlp):
    ae = make_ae(seed=2)
    x = torch.randn(64, 2, 2, device="cuda")
    y0 = mlp.accelerate(ae)[1].value(x)
    with torch.no_grad():
        ae.decoder[0].weight.mul_(1.5)
    y1 = mlp.accelerate(ae)[1].value(x)
    assert rel(y1, ae.decoder(x)) < 2e-5 and rel(y0, y1) > 1e-2


def test_symmetry_regularisers_through_the_tensor_core_chain(mlp, monkeypatch):
    """symmreg_i / symmreg_f / symmreg_r / group_action_and_jacobian with the autoencoder on the tensor cores against
    the same functions on the PyTorch modules (SINDY_B200_AE_MLP=0): loss values and dL/dΞ."""
    import model_utils
    import sindy
    import standins
    ae = make_ae(seed=5)
    _, gen = standins.make_standins(seed=0, input_dim=2, n_comps=2, hidden=32, device="cuda")
    torch.manual_seed(0)
    x = 0.5 * torch.randn(4000, 2, device="cuda")
    reg = sindy.SINDyRegression(2, 2, False, True, threshold=0.05, device="cuda", constrain_constant=True)
    reg.Xi.data = 0.3 * torch.randn_like(reg.Xi)

    def run():
        out = {}
        flow = model_utils.EulerFlowMap(reg, 0.1, 0.01)
        x_fx = torch.stack([x, flow(x)], dim=1)
        li = model_utils.symmreg_i(x_fx, ae, gen, f=flow, require_grad=True)
        reg.zero_grad(); li.backward()
        out["i"] = (float(li), reg.Xi.grad.clone())
        z_x = model_utils.encode_constant_component(ae, x)            # None when the PyTorch modules are in use
        x_fx = torch.stack([x, flow(x)], dim=1)
        lz = model_utils.symmreg_i(x_fx, ae, gen, f=flow, require_grad=True, z_x=z_x)
        reg.zero_grad(); lz.backward()
        out["z"] = (float(lz), reg.Xi.grad.clone(), z_x is not None)
        x_fx = torch.stack([x, flow(x)], dim=1)
        lf = model_utils.symmreg_f(x_fx, ae, gen, f=flow, require_grad=True)
        reg.zero_grad(); lf.backward()
        out["f"] = (float(lf), reg.Xi.grad.clone())
        lr = model_utils.symmreg_r(x, ae, gen, h=reg, require_grad=True)
        reg.zero_grad(); lr.backward()
        out["r"] = (float(lr), reg.Xi.grad.clone())
        lb = model_utils.symmreg_r(x[:512], ae, gen, h=reg, normalize='in_batch', z_mean=ae.encoder[-2].bias,
                                    require_grad=True)                  # double-vjp branch (PyTorch modules)
        reg.zero_grad(); lb.backward()
        out["b"] = (float(lb), reg.Xi.grad.clone())
        gx, jgx = model_utils.group_action_and_jacobian(x, ae, gen)
        out["g"] = (torch.stack(gx), torch.stack(jgx))
        # the reference's own precompute (vmap(jacfwd(...)), `model_utils.py:172-211`) runs under functorch transforms:
        # the tensor-core chain steps aside there and the result is the PyTorch one whatever the switch says
        gx_p, jgx_p = model_utils.precompute_symmreg_r(x[:64], ae, gen)
        out["p"] = (torch.stack(gx_p), torch.stack(jgx_p))
        return out

    fast = run()
    assert ae.__dict__.get("_sb_frozen_mlps", (None, None))[1] is not None
    monkeypatch.setenv("SINDY_B200_AE_MLP", "0")
    slow = run()
    # the encoder output of the constant data half computed once: same loss and gradient as encoding [x, f(x)] together
    assert fast["z"][2] and not slow["z"][2]
    assert abs(fast["z"][0] - fast["i"][0]) <= 1e-6 * abs(fast["i"][0]) and rel(fast["z"][1], fast["i"][1]) < 1e-5
    for k in "ifrb":
        assert abs(fast[k][0] - slow[k][0]) < 1e-4 * abs(slow[k][0]), k
        assert rel(fast[k][1], slow[k][1]) < 5e-4, k
    assert rel(fast["p"][0], slow["p"][0]) < 1e-5 and rel(fast["p"][1], slow["p"][1]) < 1e-6
    assert rel(fast["g"][0][0], slow["g"][0][0]) < 1e-5 and rel_rows(fast["g"][1][0], slow["g"][1][0]) < 1e-4


@pytest.mark.parametrize("variant", ["default", "no_narrow", "pair"])
def test_layer_kernels_write_inside_their_buffers(mlp, variant, monkeypatch):
    """Guard bands around every output buffer of the layer kernels (compute-sanitizer is not available on the GPU
    pool): panel outputs of sb_mlp_pack_rows / thin_in / gemm and the row-major outputs of thin_out / unpack_rows sit
    between sentinel regions that must come back untouched, for a ragged row count (300 = 2 tiles + 44 rows)."""
    from sindy_b200 import native
    if variant == "no_narrow":
        monkeypatch.setenv("SB_MLP_NARROW", "0")
    elif variant == "pair":
        monkeypatch.setenv("SB_MLP_PAIR", "1")
    lib = native.load()
    dev = torch.device("cuda")
    s = native._stream(dev)
    m, f, guard = 300, 512, 1 << 14                       # guard: floats on either side
    sentinel = -12345.678

    def guarded(n_floats):
        big = torch.full((n_floats + 2 * guard,), sentinel, device=dev)
        return big, big[guard:guard + n_floats]

    def intact(big, n_floats):
        return bool((big[:guard] == sentinel).all()) and bool((big[guard + n_floats:] == sentinel).all())

    n_panel = int(lib.sb_mlp_panel_bytes(m, f)) // 4
    g = torch.Generator(device=dev).manual_seed(0)
    x2 = torch.randn(m, 2, device=dev, generator=g)
    w_in, b_in = torch.randn(f, 2, device=dev, generator=g), torch.randn(f, device=dev, generator=g)
    w = torch.randn(f, f, device=dev, generator=g) / f ** 0.5
    w_out = torch.randn(2, f, device=dev, generator=g)
    big_a, pa = guarded(n_panel)
    big_c, pc = guarded(n_panel)
    big_y, y = guarded(m * 2)
    big_r, rows = guarded(m * f)
    pk = torch.empty(2 * f * f, device=dev)
    native._check(lib.sb_mlp_pack_weights(w.data_ptr(), f, f, 0, pk.data_ptr(), s), "pack")
    native._check(lib.sb_mlp_thin_in(x2.data_ptr(), m, 2, w_in.data_ptr(), b_in.data_ptr(), None, f, 1, pa.data_ptr(), s),
                  "thin_in")
    for mode, bias, mask in ((1, b_in, None), (2, None, pa), (0, None, None)):
        native._check(lib.sb_mlp_gemm(pa.data_ptr(), m, f, pk.data_ptr(), f, bias.data_ptr() if bias is not None else None,
                                      mask.data_ptr() if mask is not None else None, mode, pc.data_ptr(), s), "gemm")
    native._check(lib.sb_mlp_thin_out(pc.data_ptr(), m, f, w_out.data_ptr(), None, 2, y.data_ptr(), s), "thin_out")
    native._check(lib.sb_mlp_unpack_rows(pc.data_ptr(), m, f, rows.data_ptr(), s), "unpack")
    torch.cuda.synchronize()
    assert intact(big_a, n_panel) and intact(big_c, n_panel) and intact(big_y, m * 2) and intact(big_r, m * f)
    h1 = torch.relu(x2.double() @ w_in.double().t() + b_in.double())
    assert rel(rows.view(m, f), h1 @ w.double().t()) < 1e-5          # the last launch was the plain product
    assert rel(y.view(m, 2), (h1 @ w.double().t()) @ w_out.double().t()) < 1e-5
    big_p, pp = guarded(n_panel)
    native._check(lib.sb_mlp_pack_rows(rows.data_ptr(), m, f, pp.data_ptr(), s), "pack_rows")
    torch.cuda.synchronize()
    assert intact(big_p, n_panel)


@pytest.mark.parametrize("in_dim,hidden,out_dim,n_layers", [(1, 256, 1, 2), (3, 256, 5, 3), (8, 1024, 8, 3), (2, 512, 2, 1)])
def test_other_network_shapes_and_input_forms(mlp, in_dim, hidden, out_dim, n_layers):
    """Thin dimensions 1..8, hidden widths 256 / 512 / 1024, one to three hidden layers (n_layers = 1: no wide layer at
    all, only the two thin kernels); empty batches, non-contiguous and float64 inputs, no leading batch structure."""
    torch.manual_seed(in_dim + hidden)
    mods = [nn.Linear(in_dim, hidden), nn.ReLU()]
    for _ in range(n_layers - 1):
        mods += [nn.Linear(hidden, hidden), nn.ReLU()]
    mods += [nn.Linear(hidden, out_dim)]
    net = nn.Sequential(*mods).cuda().eval()
    for p in net.parameters():
        p.requires_grad_(False)
    fast = mlp.FrozenMLP.from_module(net)
    assert fast is not None
    net64 = nn.Sequential(*mods).double()
    x = torch.randn(333, in_dim, device="cuda")
    t = torch.randn(333, in_dim, device="cuda")
    assert rel(fast.value(x), net64(x.double())) < 2e-5
    _, jt64 = torch.autograd.functional.jvp(net64, x.double(), t.double())
    assert rel_rows(fast.value_and_jvp(x, t)[1], jt64) < 2e-5
    # empty batch, 1-D sample, non-contiguous view, float64 input
    assert fast.value(x[:0]).shape == (0, out_dim) and fast.value_and_jvp(x[:0], t[:0])[1].shape == (0, out_dim)
    assert fast.value(x[5]).shape == (out_dim,) and rel(fast.value(x[5]), net64(x[5].double())) < 2e-5
    wide = torch.randn(333, 2 * in_dim, device="cuda")
    assert rel(fast.value(wide[:, ::2]), net64(wide[:, ::2].double())) < 2e-5
    assert rel(fast.value(x.double()), net64(x.double())) < 2e-5
    xg = x.clone().requires_grad_(True)
    c = torch.randn(333, out_dim, device="cuda")
    (fast.value(xg) * c).sum().backward()
    x64 = x.double().requires_grad_(True)
    (net64(x64) * c.double()).sum().backward()
    assert rel_rows(xg.grad, x64.grad) < 2e-5
    with pytest.raises(ValueError):
        fast.value(torch.randn(4, in_dim + 1, device="cuda"))


@pytest.mark.parametrize("k,n", [(256, 512), (512, 256), (1024, 256), (256, 1024)])
def test_wide_layer_kernel_rectangular(mlp, k, n):
    """sb_mlp_gemm with different input and output widths (C (m × n) = A (m × k)·Bᵀ, B (n × k)), both weight packings."""
    from sindy_b200 import native
    lib = native.load()
    g = torch.Generator(device="cuda").manual_seed(k + n)
    m = 777
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=g)
    s = native._stream(a.device)
    pa, pc = mlp._Panel.from_rows(a), mlp._Panel(m, n, a.device)
    pk = torch.empty(2 * n * k, device="cuda")
    native._check(lib.sb_mlp_pack_weights(w.data_ptr(), n, k, 0, pk.data_ptr(), s), "pack")
    native._check(lib.sb_mlp_gemm(pa.ptr(), m, k, pk.data_ptr(), n, b.data_ptr(), None, 1, pc.ptr(), s), "gemm")
    assert rel(pc.to_rows(), torch.relu(a.double() @ w.double().t() + b.double())) < 1e-5
    wt = w.t().contiguous()                                   # (k × n): the same B from its transpose
    native._check(lib.sb_mlp_pack_weights(wt.data_ptr(), n, k, 1, pk.data_ptr(), s), "pack")
    native._check(lib.sb_mlp_gemm(pa.ptr(), m, k, pk.data_ptr(), n, None, None, 0, pc.ptr(), s), "gemm")
    assert rel(pc.to_rows(), a.double() @ w.double().t()) < 1e-5


@pytest.mark.parametrize("m,out_dim", [(300, 2), (2500, 3), (40000, 2), (1000, 8)])
def test_fused_thin_output_layer(mlp, m, out_dim, monkeypatch):
    """sb_mlp_gemm_out: y = relu(A·Wᵀ + b)·W_outᵀ + b_out from the wide layer's epilogue, with and without the wide
    activations written; y is bitwise independent of the tile schedule (narrow last round or not) because the partial
    products are formed per 64-column granule and added in granule order."""
    from sindy_b200 import native
    lib = native.load()
    g = torch.Generator(device="cuda").manual_seed(m + out_dim)
    f = 512
    a = torch.randn(m, f, device="cuda", generator=g)
    w = torch.randn(f, f, device="cuda", generator=g) / f ** 0.5
    b = torch.randn(f, device="cuda", generator=g)
    w_out = torch.randn(out_dim, f, device="cuda", generator=g) / f ** 0.5
    b_out = torch.randn(out_dim, device="cuda", generator=g)
    s = native._stream(a.device)
    pa, pc = mlp._Panel.from_rows(a), mlp._Panel(m, f, a.device)
    pk = torch.empty(2 * f * f, device="cuda")
    native._check(lib.sb_mlp_pack_weights(w.data_ptr(), f, f, 0, pk.data_ptr(), s), "pack")
    scratch = torch.empty(int(lib.sb_mlp_partials_bytes(m, f, out_dim)) // 4, device="cuda")
    ys = []
    for narrow, with_panel in (("1", True), ("1", False), ("0", False)):
        monkeypatch.setenv("SB_MLP_NARROW", narrow)
        y = torch.full((m, out_dim), float("nan"), device="cuda")
        scratch.fill_(float("nan"))
        native._check(lib.sb_mlp_gemm_out(pa.ptr(), m, f, pk.data_ptr(), f, b.data_ptr(), None, 1,
                                          pc.ptr() if with_panel else None, w_out.data_ptr(), b_out.data_ptr(), out_dim,
                                          scratch.data_ptr(), y.data_ptr(), s), "gemm_out")
        ys.append(y)
    h = torch.relu(a.double() @ w.double().t() + b.double())
    assert rel(ys[0], h @ w_out.double().t() + b_out.double()) < 1e-5
    assert rel(pc.to_rows(), h) < 1e-5
    assert torch.equal(ys[0], ys[1]) and torch.equal(ys[0], ys[2])
