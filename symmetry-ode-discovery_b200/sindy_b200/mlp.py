"""The frozen autoencoder's MLPs on the tensor cores.

`FrozenMLP` evaluates an encoder / decoder of the reference's `AutoEncoder` (`autoencoder.py:38-66`, ae_arch 'mlp':
Linear [+ BatchNorm1d in eval mode] + ReLU blocks between a thin input and a thin output) through the `sb_mlp_*` entry
points: one tcgen05 kernel launch per 512-wide layer, activations in the tensor core's panel format between layers.

    value(x)            = module(x)                                  (what `autoencoder.encode / decode` return)
    value_and_jvp(x, t) = (module(x), J_module(x)·t)                 (`jvp(autoencoder.decoder, z, v)`,
                                                                      `model_utils.py:32,38-53`, `autoencoder.py:110-132`)

Both are differentiable with respect to x and t (one transpose chain per cotangent, same kernels with Wᵀ): that is all
`loss.backward()` asks of a FROZEN network with piecewise-linear activations — the reference's double-vjp graph
differentiates ReLU twice and gets exactly the zeros this module never computes. Weights are read once
(`FrozenMLP.from_module`), BatchNorm statistics folded in; a module that is not of this form, is in training mode or
has trainable parameters returns None and the caller keeps the PyTorch path.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch.autograd import Function

from . import native

LIN, RELU, MASK = 0, 1, 2
_THIN_MAX = 8
_CHUNK_ROWS = 1 << 18          # rows per pass when no graph is recorded (1 GiB of panel per 512-wide activation)


def _check(status, what):
    native._check(status, what)


def _stream(dev):
    return native._stream(dev)


class _Panel:
    """An (m × f) activation tensor in the panel format (device buffer + logical shape)."""
    __slots__ = ("buf", "m", "f")

    def __init__(self, m: int, f: int, dev: torch.device):
        nbytes = int(native.load().sb_mlp_panel_bytes(m, f))
        self.buf = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
        self.m, self.f = m, f

    def ptr(self):
        return self.buf.data_ptr()

    def to_rows(self) -> torch.Tensor:
        out = torch.empty(self.m, self.f, dtype=torch.float32, device=self.buf.device)
        with torch.cuda.device(out.device):
            _check(native.load().sb_mlp_unpack_rows(self.ptr(), self.m, self.f, out.data_ptr(), _stream(out.device)),
                   "sb_mlp_unpack_rows")
        return out

    @staticmethod
    def from_rows(x: torch.Tensor) -> "_Panel":
        x = native._f32c(x, "x")
        p = _Panel(x.shape[0], x.shape[1], x.device)
        with torch.cuda.device(x.device):
            _check(native.load().sb_mlp_pack_rows(x.data_ptr(), x.shape[0], x.shape[1], p.ptr(), _stream(x.device)),
                   "sb_mlp_pack_rows")
        return p


def _flatten(seq) -> list:
    out = []
    for m in seq:
        if isinstance(m, torch.nn.Sequential):
            out.extend(_flatten(m))
        else:
            out.append(m)
    return out


def fold_layers(module):
    """Parse an `nn.Sequential` of Linear / BatchNorm1d (eval) / ReLU / Reshape / Identity (`autoencoder.py:38-66`,
    `model.py:17-58`) into ([(W, b), ...] in float64 with the BatchNorm statistics folded in, ReLU between all layers,
    output shape of the last Reshape or None). Returns None if the module has another form, is training, or has
    trainable parameters."""
    if not isinstance(module, torch.nn.Module):
        return None
    seq = getattr(module, "layers", module)          # EncoderMLP / DecoderMLP wrap a Sequential
    if not isinstance(seq, torch.nn.Sequential):
        return None
    if any(p.requires_grad for p in module.parameters()):
        return None
    layers, out_shape = [], None
    pending_relu_ok = False
    for m in _flatten(seq):
        if isinstance(m, torch.nn.Linear):
            if layers and not layers[-1][2]:
                return None                           # two affine maps without an activation between them
            w = m.weight.detach().double()            # a parametrised (orthogonal) weight is materialised here
            b = m.bias.detach().double() if m.bias is not None else torch.zeros(w.shape[0], dtype=torch.float64,
                                                                                 device=w.device)
            layers.append([w, b, False])
            pending_relu_ok = True
        elif isinstance(m, torch.nn.BatchNorm1d):
            if m.training or not m.track_running_stats or not layers or layers[-1][2]:
                return None
            scale = 1.0 / torch.sqrt(m.running_var.double() + m.eps)
            shift = -m.running_mean.double() * scale
            if m.affine:
                scale = scale * m.weight.detach().double()
                shift = shift * m.weight.detach().double() + m.bias.detach().double()
            layers[-1][0] = layers[-1][0] * scale[:, None]
            layers[-1][1] = layers[-1][1] * scale + shift
        elif isinstance(m, torch.nn.ReLU):
            if not pending_relu_ok:
                return None
            layers[-1][2] = True
            pending_relu_ok = False
        elif isinstance(m, torch.nn.Identity):
            continue
        elif type(m).__name__ == "Reshape" and hasattr(m, "shape"):
            out_shape = tuple(m.shape)
        else:
            return None
    if len(layers) < 2 or layers[-1][2] or not all(l[2] for l in layers[:-1]):
        return None
    return [(w, b) for w, b, _ in layers], (out_shape if out_shape and len(out_shape) > 2 else None)


class FrozenMLP:
    """layers: [(W (out × in), b (out))...] with ReLU after every layer but the last."""

    def __init__(self, layers: List[Tuple[torch.Tensor, torch.Tensor]], out_shape: Optional[tuple] = None):
        if len(layers) < 2:
            raise ValueError("FrozenMLP needs a thin input layer and a thin output layer")
        dev = layers[0][0].device
        if dev.type != "cuda":
            raise ValueError("FrozenMLP lives on a CUDA device (there is no CPU path)")
        self.device = dev
        self.in_dim = layers[0][0].shape[1]
        self.out_dim = layers[-1][0].shape[0]
        self.width = layers[0][0].shape[0]
        self.out_shape = out_shape
        f = self.width
        if self.in_dim > _THIN_MAX or self.out_dim > _THIN_MAX or f % 256 or f > 2048:
            raise ValueError(f"unsupported MLP shape {self.in_dim} -> {f} -> {self.out_dim}")
        for w, _ in layers[1:-1]:
            if tuple(w.shape) != (f, f):
                raise ValueError("hidden layers must be square")
        if layers[-1][0].shape[1] != f:
            raise ValueError("output layer does not match the hidden width")
        c = lambda t: t.detach().to(dev, torch.float32).contiguous()
        self.w_in, self.b_in = c(layers[0][0]), c(layers[0][1])                  # (f × in), (f)
        self.w_in_t = self.w_in.t().contiguous()                                 # (in × f): transpose chain's last step
        self.w_out, self.b_out = c(layers[-1][0]), c(layers[-1][1])              # (out × f), (out)
        self.w_out_t = self.w_out.t().contiguous()                               # (f × out): transpose chain's first step
        self.hidden = []                                                         # [(packed W, packed Wᵀ, bias)]
        lib = native.load()
        with torch.cuda.device(dev):
            for w, b in layers[1:-1]:
                w = c(w)
                pk = torch.empty(2 * f * f, dtype=torch.float32, device=dev)
                pk_t = torch.empty(2 * f * f, dtype=torch.float32, device=dev)
                _check(lib.sb_mlp_pack_weights(w.data_ptr(), f, f, 0, pk.data_ptr(), _stream(dev)), "sb_mlp_pack_weights")
                # cotangents: G_in = G_out · W, i.e. B[n][k] = W[k][n]
                _check(lib.sb_mlp_pack_weights(w.data_ptr(), f, f, 1, pk_t.data_ptr(), _stream(dev)),
                       "sb_mlp_pack_weights")
                self.hidden.append((pk, pk_t, c(b)))

    # ---- construction from the reference's modules ----
    @classmethod
    def from_module(cls, module) -> Optional["FrozenMLP"]:
        """An encoder / decoder of the reference's AutoEncoder as a FrozenMLP, or None (see `fold_layers`)."""
        folded = fold_layers(module)
        if folded is None or folded[0][0][0].device.type != "cuda":
            return None
        try:
            return cls(folded[0], folded[1])
        except ValueError:
            return None

    # ---- chains (rows: (m × in_dim) fp32 contiguous) ----
    def _chain(self, first, m, wide, out_w, out_b, out_dim, keep_all, keep_last):
        """thin input layer -> wide layers -> thin output layer. `first(panel)` launches the thin input kernel into the
        panel; wide = [(packed weights, bias or None, mode, mask panel or None)]; the LAST wide layer runs with the thin
        output layer fused into its epilogue (`sb_mlp_gemm_out`) and writes its own activations only if keep_last.
        Returns (kept panels, y (m × out_dim))."""
        lib, dev, f = native.load(), self.device, self.width
        kept = []
        y = torch.empty(m, out_dim, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            cur = _Panel(m, f, dev)
            first(cur)
            if keep_all or (keep_last and not wide):
                kept.append(cur)
            for i, (pk, bias, mode, mask) in enumerate(wide):
                last = i + 1 == len(wide)
                want_panel = keep_all or not last or keep_last
                nxt = _Panel(m, f, dev) if want_panel else None
                args = (cur.ptr(), m, f, pk.data_ptr(), f, native._ptr(bias), mask.ptr() if mask is not None else None, mode,
                        nxt.ptr() if nxt is not None else None)
                if last:
                    scratch = torch.empty(int(lib.sb_mlp_partials_bytes(m, f, out_dim)) // 4, dtype=torch.float32, device=dev)
                    _check(lib.sb_mlp_gemm_out(*args, out_w.data_ptr(), native._ptr(out_b), out_dim, scratch.data_ptr(),
                                               y.data_ptr(), _stream(dev)), "sb_mlp_gemm_out")
                else:
                    _check(lib.sb_mlp_gemm(*args, _stream(dev)), "sb_mlp_gemm")
                if nxt is not None:
                    cur = nxt
                    if keep_all or (last and keep_last):
                        kept.append(cur)
            if not wide:        # no wide layer at all: the thin output kernel reads the first panel
                _check(lib.sb_mlp_thin_out(cur.ptr(), m, f, out_w.data_ptr(), native._ptr(out_b), out_dim, y.data_ptr(),
                                           _stream(dev)), "sb_mlp_thin_out")
        return kept, y

    def _value_chain(self, x: torch.Tensor, keep: bool):
        """Hidden activations H_1..H_L (panel format; none if not keep) and y (m × out)."""
        lib, dev, f, m = native.load(), self.device, self.width, x.shape[0]

        def first(panel):
            _check(lib.sb_mlp_thin_in(x.data_ptr(), m, self.in_dim, self.w_in.data_ptr(), self.b_in.data_ptr(), None, f,
                                      RELU, panel.ptr(), _stream(dev)), "sb_mlp_thin_in")
        wide = [(pk, b, RELU, None) for pk, _, b in self.hidden]
        return self._chain(first, m, wide, self.w_out, self.b_out, self.out_dim, keep_all=keep, keep_last=keep)

    def _tangent_chain(self, t: torch.Tensor, hs) -> torch.Tensor:
        """J(x)·t with the ReLU masks of the value chain's activations hs."""
        lib, dev, f, m = native.load(), self.device, self.width, t.shape[0]

        def first(panel):
            _check(lib.sb_mlp_thin_in(t.data_ptr(), m, self.in_dim, self.w_in.data_ptr(), None, hs[0].ptr(), f, MASK,
                                      panel.ptr(), _stream(dev)), "sb_mlp_thin_in")
        wide = [(pk, None, MASK, h) for (pk, _, _), h in zip(self.hidden, hs[1:])]
        return self._chain(first, m, wide, self.w_out, None, self.out_dim, keep_all=False, keep_last=False)[1]

    def _transpose_chain(self, g: torch.Tensor, hs) -> torch.Tensor:
        """J(x)ᵀ·g (m × in) for a cotangent g (m × out)."""
        lib, dev, f, m = native.load(), self.device, self.width, g.shape[0]

        def first(panel):
            _check(lib.sb_mlp_thin_in(g.data_ptr(), m, self.out_dim, self.w_out_t.data_ptr(), None, hs[-1].ptr(), f, MASK,
                                      panel.ptr(), _stream(dev)), "sb_mlp_thin_in")
        wide = [(pk_t, None, MASK, h) for (_, pk_t, _), h in zip(reversed(self.hidden), reversed(hs[:-1]))]
        return self._chain(first, m, wide, self.w_in_t, None, self.in_dim, keep_all=False, keep_last=False)[1]

    # ---- public operators ----
    def _rows(self, x: torch.Tensor, name: str) -> torch.Tensor:
        if x.shape[-1] != self.in_dim:
            raise ValueError(f"`{name}` has last dimension {x.shape[-1]}, the network expects {self.in_dim}")
        return native._f32c(x, name).reshape(-1, self.in_dim)

    def _shape_out(self, y: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
        if self.out_shape is not None:
            return y.reshape(self.out_shape)
        return y.reshape(*like.shape[:-1], self.out_dim)

    def value(self, x: torch.Tensor) -> torch.Tensor:
        rows = self._rows(x, "x")
        if rows.shape[0] == 0:
            return self._shape_out(rows.new_zeros(0, self.out_dim), x)
        if torch.is_grad_enabled() and x.requires_grad:
            y = _Value.apply(rows, self)
        else:
            ys = [self._value_chain(rows[i:i + _CHUNK_ROWS], keep=False)[1] for i in range(0, rows.shape[0], _CHUNK_ROWS)]
            y = torch.cat(ys) if ys else rows.new_zeros(0, self.out_dim)
        return self._shape_out(y, x)

    __call__ = value

    def value_rows(self, x: torch.Tensor) -> torch.Tensor:
        """module applied to the rows of a 2-D tensor, (rows × out) without the module's own output reshape."""
        saved, self.out_shape = self.out_shape, None
        try:
            return self.value(x.reshape(-1, self.in_dim))
        finally:
            self.out_shape = saved

    def value_and_jvp(self, x: torch.Tensor, t: torch.Tensor):
        rows, trows = self._rows(x, "x"), self._rows(t, "t")
        if trows.shape[0] != rows.shape[0]:
            raise ValueError("x and t must describe the same rows")
        if rows.shape[0] == 0:
            z = rows.new_zeros(0, self.out_dim)
            return self._shape_out(z, x), self._shape_out(z.clone(), x)
        if torch.is_grad_enabled() and (x.requires_grad or t.requires_grad):
            y, jt = _ValueJvp.apply(rows, trows, self)
        else:
            ys, jts = [], []
            for i in range(0, rows.shape[0], _CHUNK_ROWS):
                hs, y = self._value_chain(rows[i:i + _CHUNK_ROWS], keep=True)
                ys.append(y)
                jts.append(self._tangent_chain(trows[i:i + _CHUNK_ROWS], hs))
            y = torch.cat(ys) if ys else rows.new_zeros(0, self.out_dim)
            jt = torch.cat(jts) if jts else rows.new_zeros(0, self.out_dim)
        return self._shape_out(y, x), self._shape_out(jt, x)


class _Value(Function):
    @staticmethod
    def forward(ctx, rows, mlp: FrozenMLP):
        hs, y = mlp._value_chain(rows, keep=True)
        ctx.mlp, ctx.hs = mlp, hs
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        return ctx.mlp._transpose_chain(native._f32c(gy, "gy"), ctx.hs), None


class _ValueJvp(Function):
    @staticmethod
    def forward(ctx, rows, trows, mlp: FrozenMLP):
        hs, y = mlp._value_chain(rows, keep=True)
        jt = mlp._tangent_chain(trows, hs)
        ctx.mlp, ctx.hs = mlp, hs
        ctx.set_materialize_grads(False)      # an unused output (symmreg_i takes only the tangent) costs no chain
        return y, jt

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy, gjt):
        # d(J(x)·t)/dx vanishes almost everywhere for ReLU networks (the reference's autograd returns the same zeros)
        gx = gt = None
        if ctx.needs_input_grad[0] and gy is not None:
            gx = ctx.mlp._transpose_chain(native._f32c(gy, "gy"), ctx.hs)
        if ctx.needs_input_grad[1] and gjt is not None:
            gt = ctx.mlp._transpose_chain(native._f32c(gjt, "gjt"), ctx.hs)
        return gx, gt, None


def accelerate(autoencoder):
    """(encoder, decoder) as FrozenMLPs for a frozen, eval-mode reference AutoEncoder on a CUDA device, else None.
    Cached on the module and rebuilt when a parameter or buffer changed (version counters)."""
    if not isinstance(autoencoder, torch.nn.Module) or autoencoder.training:
        return None
    enc, dec = getattr(autoencoder, "encoder", None), getattr(autoencoder, "decoder", None)
    if not isinstance(enc, torch.nn.Module) or not isinstance(dec, torch.nn.Module):
        return None
    stamp = (tuple((t.data_ptr(), t._version, t.requires_grad)
                   for t in list(autoencoder.parameters()) + list(autoencoder.buffers())),
             tuple(type(m).__name__ for m in autoencoder.modules()))
    cached = autoencoder.__dict__.get("_sb_frozen_mlps")
    if cached is not None and cached[0] == stamp:
        return cached[1]
    pair = None
    if not any(p.requires_grad for p in autoencoder.parameters()):
        e, d = FrozenMLP.from_module(enc), FrozenMLP.from_module(dec)
        if e is not None and d is not None:
            pair = (e, d)
    autoencoder.__dict__["_sb_frozen_mlps"] = (stamp, pair)
    return pair
