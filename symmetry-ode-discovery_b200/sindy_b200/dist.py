"""Sample-sharded train step: one process per GPU, each rank owns its trajectories, ONE all-reduce per step.

The reference is single-device (`parser_utils.py:118`); the shard/all-reduce layer is new. Every rank runs the
fused kernel over its own samples and obtains the packed fp64 SUMS `[Σr², n, Σ r_iΘ_k (d×K), (Gram), (ΘᵀẊ)]`
(see include/sindy_b200.h). Sums — not per-rank means — are all-reduced, then every rank divides by the GLOBAL
n·d, so the loss/gradient equal the single-process values for any (uneven) sharding and Ξ / mask / optimiser
state stay replicated without further communication. The message is d·K+2 doubles (1.4 KB for d=3, K=56):
latency-bound, so the whole step (W packing, kernel, all-reduce, epilogue) can be captured in a CUDA graph.

Launch counts per closure: single rank = 2 (pack Ξ⊙mask into the constant bank; fused kernel whose last block
writes loss and gradient). Multi rank = 3 + the all-reduce (mask multiply + pack, fused kernel, epilogue).

`HostStreamedStep` is the host-buffer entry point: x/dx live in pinned host memory and are streamed through
two device staging buffers on two copy streams while the kernel consumes the previous chunk.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import native
from .native import Library


def bind_to_gpu_numa_node(device) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs: the PCI device's `numa_node` and the
    node's `cpulist`), so that pinned host buffers allocated afterwards are node-local: with one process per GPU on a
    two-socket box, host->device streams otherwise cross the socket interconnect. Returns the node, or None if the
    topology is not exposed (then nothing is changed)."""
    try:
        p = torch.cuda.get_device_properties(device)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:   # no CUDA device / properties without PCI ids
        return None
    try:
        path = f"/sys/bus/pci/devices/{bdf.lower()}/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError):
        return None


def _world(group) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def mse_from_sums(packed: torch.Tensor, lib: Library, w: torch.Tensor, mask: Optional[torch.Tensor] = None,
                  w_l1: float = 0.0, params_l1: Optional[torch.Tensor] = None):
    """Epilogue on the (all-reduced) packed sums in plain torch ops: loss = Σr²/(n·d) [+ w_l1·‖Ξ‖₁] and dL/dΞ
    (`train.py:663-664,680-683,689`). Device-agnostic twin of sb_step_epilogue (used by the gloo tests)."""
    d, K = lib.dim, lib.K
    denom = packed[1] * d
    loss = packed[0] / denom
    grad = packed[2:2 + d * K].view(d, K) * (2.0 / denom)
    if mask is not None:
        grad = grad * mask
    if w_l1 != 0.0 and params_l1 is not None:
        loss = loss + w_l1 * params_l1.abs().sum()
        grad = grad + w_l1 * torch.sign(params_l1)
    return loss, grad.to(w.dtype)


class PeerExchange:
    """Symmetric (peer-mapped) buffers for the in-kernel all-reduce of sb_closure_peer, allocated through
    torch.distributed's symmetric memory (CUDA VMM + NVLink P2P). `available()` is False when the runtime cannot
    provide it; the sharded step then uses an NCCL all-reduce + epilogue launch instead."""

    def __init__(self, lib: Library, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD if group is None else group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        n_bytes = native.peer_buffer_bytes(lib, self.world)
        self.buf = symm_mem.empty((n_bytes + 7) // 8, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.epoch = torch.zeros(2, dtype=torch.int32, device=device)   # [epoch counter, sticky status word]
        torch.cuda.synchronize(device)
        dist.barrier(group)       # every rank's buffer is zeroed before anybody's kernel may write into it

    @staticmethod
    def create(lib: Library, device: torch.device, group=None):
        if os.environ.get("SB_NO_PEER", "0") == "1":
            return None
        try:
            return PeerExchange(lib, device, group)
        except Exception as exc:  # symmetric memory unsupported on this system / build
            if os.environ.get("SB_PEER_DEBUG"):
                print(f"[sindy_b200] peer exchange unavailable: {exc!r}")
            return None


class ShardedTrainStep:
    """loss, grad = step(Ξ, mask) over samples sharded across the ranks of `group` (or a single process).

    local_sums: callable (W) -> packed fp64 sums of THIS rank's shard. The default runs the CUDA kernel on
    (x, dx); tests inject a CPU stand-in to exercise the combine logic with the gloo backend.
    sgd_lr: if set, `Ξ -= sgd_lr·grad` is applied to the (static) parameters inside the step — a stand-in for the
    optimiser update so that consecutive benchmark steps see different coefficients.
    """

    def __init__(self, lib: Library, x: Optional[torch.Tensor] = None, dx: Optional[torch.Tensor] = None,
                 flags: int = native.SB_STEP_LOSS | native.SB_STEP_GRAD, group=None,
                 local_sums: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, use_graph: bool = False,
                 sgd_lr: Optional[float] = None, sym_gens=None, w_sym: float = 0.0, use_peer: bool = True):
        self.lib, self.flags, self.group = lib, flags, group
        # linear Lie-derivative regulariser (`train.py:503-507`) through the Gram matrix: see symreg.py
        self.w_sym = float(w_sym)
        self._sym = None
        if sym_gens is not None and len(sym_gens) > 0:
            from . import symreg
            dev = x.device if x is not None else torch.device("cpu")
            self._sym = (torch.stack([torch.as_tensor(v, dtype=torch.float64) for v in sym_gens]).to(dev),
                         torch.stack([symreg.lie_matrix(lib, v) for v in sym_gens]).to(dev))
        self.x, self.dx = x, dx
        self._custom = local_sums
        self._use_graph = use_graph and local_sums is None
        self._sgd_lr = sgd_lr
        self._graph = None
        self._bufs = None
        self.xi = None      # static parameters (graph mode / sgd mode)
        self.mask = None
        # in-kernel all-reduce over NVLink peer memory when several ranks share the step (no sym-reg Gram yet)
        self.peer = None
        if local_sums is None and x is not None and _world(group) > 1 and self._sym is None and use_peer:
            self.peer = PeerExchange.create(lib, x.device, group)

    # -- pieces ----------------------------------------------------------------------------------------
    def _buffers(self, dev):
        if self._bufs is None:
            d, K = self.lib.dim, self.lib.K
            n_step = self.lib.step_out_len(self.flags)
            n_gram = self.lib.step_out_len(native.SB_STEP_GRAM) if self._sym is not None else 0
            flat = torch.empty(n_step + n_gram, dtype=torch.float64, device=dev)  # ONE all-reduce covers both
            self._flat = flat
            self._gram = flat[n_step:] if n_gram else None
            self._bufs = (flat[:n_step], torch.empty((), dtype=torch.float32, device=dev),
                          torch.empty(d, K, dtype=torch.float32, device=dev))
        return self._bufs

    def _sym_terms(self, xi, mask):
        """(Σ_v tr(A_v G A_vᵀ), its gradient w.r.t. Ξ) from the (all-reduced) Gram; A_v = W M_v − v W."""
        K = self.lib.K
        vs, Ms = self._sym
        G = self._gram[2:].view(K, K)
        W = (xi if mask is None else xi * mask).double()
        A = W @ Ms - vs @ W                       # (V, d, K)
        AG = A @ G
        loss = (AG * A).sum()
        gA = 2.0 * AG
        gW = (gA @ Ms.transpose(1, 2)).sum(0) - (vs.transpose(1, 2) @ gA).sum(0)
        if mask is not None:
            gW = gW * mask
        return loss, gW

    def local_sums(self, w: torch.Tensor) -> torch.Tensor:
        if self._custom is not None:
            return self._custom(w)
        packed, _, _ = self._buffers(self.x.device)
        return native.train_step(self.x, self.dx, w, self.lib, self.flags, out=packed)

    def closure(self, xi: torch.Tensor, mask: Optional[torch.Tensor], w_l1: float):
        """(loss, dL/dΞ): the eager step."""
        world = _world(self.group)
        if self._custom is not None:
            wm = xi if mask is None else xi * mask
            packed = self._custom(wm)
            if world > 1:
                dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group)
            return mse_from_sums(packed, self.lib, xi, mask, w_l1, xi if w_l1 != 0.0 else None)
        packed, loss, grad = self._buffers(self.x.device)
        if self._sym is not None:   # second pass over x only: power sums -> Gram (data term and Gram share x in L2/HBM)
            native.train_step(self.x, None, None, self.lib, native.SB_STEP_GRAM, out=self._gram)
        if world == 1:
            native.closure(self.x, self.dx, xi, mask, self.lib, w_l1, packed=packed, loss=loss, grad=grad)
        elif self.peer is not None:
            native.closure_peer(self.x, self.dx, xi, mask, self.lib, self.peer.ptrs, self.peer.rank, self.peer.epoch,
                                w_l1, packed=packed, loss=loss, grad=grad)
        else:
            wm = xi if mask is None else xi * mask
            native.train_step(self.x, self.dx, wm, self.lib, self.flags, out=packed)
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
            native.step_epilogue(packed, xi, mask, self.lib, w_l1, loss=loss, grad=grad)
        if self._sym is not None:
            ls, gs = self._sym_terms(xi, mask)
            loss = loss + self.w_sym * ls.to(loss.dtype)
            grad = grad + (self.w_sym * gs).to(grad.dtype)
        return loss, grad

    def _body(self, w_l1):
        loss, grad = self.closure(self.xi, self.mask, w_l1)
        if self._sgd_lr is not None:
            self.xi.add_(grad, alpha=-self._sgd_lr)
        return loss, grad

    # -- public ----------------------------------------------------------------------------------------
    def step(self, xi: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None, w_l1: float = 0.0):
        """One closure evaluation. With use_graph or sgd_lr the parameters are static: pass `xi` to (re)load them,
        or None to keep the current ones (sgd mode advances them in place)."""
        if not self._use_graph and self._sgd_lr is None:
            return self.closure(xi, mask, w_l1)
        if self.xi is None:
            if xi is None:
                raise ValueError("the first call needs the parameters")
            self.xi = xi.detach().clone()
            self.mask = None if mask is None else mask.detach().clone()
        elif xi is not None:
            self.xi.copy_(xi)
            if mask is not None:
                self.mask.copy_(mask)
        if not self._use_graph:
            return self._body(w_l1)
        if self._graph is None:
            keep = self.xi.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside capture: allocations, NCCL channels, occupancy queries
                for _ in range(3):
                    self._body(w_l1)
            torch.cuda.current_stream().wait_stream(side)
            self.xi.copy_(keep)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._static_res = self._body(w_l1)
            self.xi.copy_(keep)
        self._graph.replay()
        return self._static_res


class FitStepper:
    """A whole training iteration (closure + optimiser update, `train.py:512-530` without sym-reg) as ONE kernel
    launch per rank, replayed as a CUDA graph: the fused kernel's last block all-reduces the sums over NVLink peer
    memory (several ranks), evaluates loss and gradient, applies Adam/SGD to Ξ in place and packs Ξ⊙mask into the
    constant bank for the next launch (sb_fit_step). Ξ, mask and the optimiser state are replicated: every rank
    applies the identical update. Needs a specialised library; raises otherwise (no fallback).

    step() -> loss tensor (static buffer; the value at the parameters BEFORE the update). `xi`, `grad` are the
    static parameter / gradient tensors. Call load(xi, mask) after changing parameters or mask from outside
    (thresholding): it also resets nothing else — the Adam state is kept, like torch's optimiser does.
    """

    def __init__(self, lib: Library, x: torch.Tensor, dx: torch.Tensor, kind: str = "adam", lr: float = 1e-3,
                 betas=(0.9, 0.999), eps: float = 1e-8, w_mse: float = 1.0, w_l1: float = 0.0, group=None,
                 use_graph: bool = True, use_peer: bool = True, sym_gens=None, w_sym: float = 0.0,
                 local_only: bool = False):
        self.lib, self.x, self.dx = lib, x, dx
        world = 1 if local_only else _world(group)   # local_only: this process's data alone, even under torchrun
        # linear Lie-derivative regulariser (`train.py:503-507`): the Gram matrix of the (fixed) data set is formed
        # once — one pass of the moment kernel per rank and one all-reduce — and turned into the quadratic form H;
        # every iteration then evaluates w_sym·wᵀHw and its gradient inside the kernel's epilogue
        self.sym_quad, self.w_sym = None, float(w_sym)
        if sym_gens is not None and len(sym_gens) > 0 and w_sym != 0.0:
            from . import symreg
            G = symreg.gram(x, lib).clone()
            if world > 1:
                dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
            self.sym_quad = symreg.quadratic_form(lib, sym_gens, G).to(torch.float32).contiguous()
        self.kind, self.lr, self.betas, self.eps, self.w_mse, self.w_l1 = kind, lr, betas, eps, w_mse, w_l1
        dev = x.device
        d, K = lib.dim, lib.K
        self.xi = torch.zeros(d, K, dtype=torch.float32, device=dev)
        self.mask = torch.ones(d, K, dtype=torch.float32, device=dev)
        self.state = native.fit_state(lib, dev)
        self.packed = torch.empty(2 + d * K, dtype=torch.float64, device=dev)
        self.loss = torch.empty((), dtype=torch.float32, device=dev)
        self.grad = torch.empty(d, K, dtype=torch.float32, device=dev)
        self.peer = None
        if world > 1:
            self.peer = PeerExchange.create(lib, dev, group) if use_peer else None
            if self.peer is None:
                raise RuntimeError("FitStepper over several ranks needs the peer exchange (symmetric memory); use "
                                   "ShardedTrainStep for the NCCL path")
        self.loss_hist = torch.zeros(64, dtype=torch.float32, device=dev)
        self._use_graph = use_graph
        self._graphs = {}
        self._warm = False
        self._loaded = False
        self._slot_gen = -1       # generation of the device's resident coefficient slot when this stepper last loaded it
        if self.peer is not None:
            dist.barrier(group)   # nobody launches an exchanging kernel before every rank has built its stepper

    def load(self, xi: torch.Tensor, mask: Optional[torch.Tensor] = None, reset_state: bool = False):
        self.xi.copy_(xi)
        if mask is not None:
            self.mask.copy_(mask)
        if reset_state:
            self.state.zero_()
        self._slot_gen = native.load_w(self.xi, self.mask, self.lib)
        self._loaded = True

    def _own_slot(self):
        """Before every launch / replay: if another fit has loaded the device's resident coefficient slot since this
        stepper did, re-pack — xi and mask live in this stepper's own memory. (Forward / closure / STLSQ calls use the
        scratch slot and never cause this.)"""
        if native.slot_generation(self.x.device) != self._slot_gen:
            self._slot_gen = native.load_w(self.xi, self.mask, self.lib)

    def check(self):
        """Raises if a peer was lost during an in-kernel exchange (sticky status word of sb_fit_step; reading it
        synchronises the device). After a loss the kernels leave Ξ, the optimiser state and the resident coefficients
        untouched; the ranks must be re-synchronised (load) before the word is cleared."""
        if self.peer is not None:
            st = int(self.peer.epoch[1])
            if st != 0:
                raise native.SindyB200Error(
                    f"rank {self.peer.rank}: a peer did not arrive at the in-kernel all-reduce of epoch {st} within "
                    f"$SB_PEER_TIMEOUT_MS; parameters were left at their last good values")

    def _launch(self, loss=None):
        p = self.peer
        native.fit_step(self.x, self.dx, self.xi, self.mask, self.lib, self.kind, self.lr, self.betas, self.eps,
                        self.w_mse, self.w_l1, state=self.state, w_resident=True, packed=self.packed,
                        loss=self.loss if loss is None else loss, grad=self.grad,
                        peer_ptrs=p.ptrs if p else None, rank=p.rank if p else 0, epoch=p.epoch if p else None,
                        sym_quad=self.sym_quad, w_sym=self.w_sym)

    def _capture(self, n_iters):
        """CUDA graph of n_iters consecutive iterations; iteration i writes its loss to loss_hist[i] (n_iters > 1)."""
        if not self._warm:
            # first-call attribute / occupancy queries and the workspace allocation must not happen under capture;
            # the warm-up iteration is undone (parameters, Adam state) — every rank does the same, so the peer epoch
            # stays consistent
            keep_xi, keep_state = self.xi.clone(), self.state.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._launch()
            torch.cuda.current_stream().wait_stream(side)
            self.xi.copy_(keep_xi)
            self.state.copy_(keep_state)
            self._slot_gen = native.load_w(self.xi, self.mask, self.lib)
            self._warm = True
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if n_iters == 1:
                self._launch()
            else:
                for i in range(n_iters):
                    self._launch(self.loss_hist[i])
        return g

    def step(self):
        """One iteration; returns the (static) loss tensor: the value at the parameters BEFORE the update."""
        if not self._loaded:
            raise ValueError("call load(xi, mask) first")
        self._own_slot()
        if not self._use_graph:
            self._launch()
            return self.loss
        if 1 not in self._graphs:
            self._graphs[1] = self._capture(1)
        self._graphs[1].replay()
        return self.loss

    def step_from_host(self, x_host: torch.Tensor, dx_host: torch.Tensor, chunk_samples: int = 1 << 22):
        """One iteration whose samples live in (pinned) HOST memory — the host-buffer entry point of the one-launch
        iteration: x and dx are copied host->device in chunks on the current stream into the stepper's resident
        buffers (asynchronously when the host tensors are pinned), then ONE sb_fit_step launch consumes them. The copy
        is PCIe-bound (2.4 GB at C5) and the kernel is ~3 % of it, so the chunks only bound the size of a single
        transfer. Returns the loss tensor; `h2d_bytes` holds the bytes copied."""
        if not self._loaded:
            raise ValueError("call load(xi, mask) first")
        if x_host.is_cuda or dx_host.is_cuda:
            raise ValueError("step_from_host takes host tensors")
        if x_host.shape != self.x.shape or dx_host.shape != self.dx.shape:
            raise ValueError(f"host tensors must have the shard's shape {tuple(self.x.shape)}")
        self._own_slot()
        n = x_host.shape[0]
        for start in range(0, n, int(chunk_samples)):
            stop = min(start + int(chunk_samples), n)
            self.x[start:stop].copy_(x_host[start:stop], non_blocking=True)
            self.dx[start:stop].copy_(dx_host[start:stop], non_blocking=True)
        self.h2d_bytes = 2 * n * self.lib.dim * 4
        self._launch()
        return self.loss

    def run(self, n_iters: int, unroll: int = 10):
        """n_iters iterations with no host involvement in between: graphs of `unroll` back-to-back launches (kernels
        inside one graph follow each other ~1 us apart, separate replays ~5 us). Returns the loss tensor of the last
        iteration; `loss_hist[:unroll]` holds the losses of the last full group."""
        if not self._loaded:
            raise ValueError("call load(xi, mask) first")
        self._own_slot()
        if not self._use_graph:
            for _ in range(n_iters):
                self._launch()
            return self.loss
        unroll = max(1, min(int(unroll), self.loss_hist.numel()))
        full, rest = divmod(int(n_iters), unroll)
        last = self.loss
        if full and unroll > 1:
            if unroll not in self._graphs:
                self._graphs[unroll] = self._capture(unroll)
            for _ in range(full):
                self._graphs[unroll].replay()
            last = self.loss_hist[unroll - 1]
        elif full:
            rest += full
        for _ in range(rest):
            last = self.step()
        return last


class HostStreamedStep:
    """Fused train step over (x, dx) that live in PINNED HOST memory: chunks are copied host->device on two
    alternating streams into two staging buffers and reduced by the kernel as they land; the packed sums of the
    chunks are added on the device. The timed region of bench.py's `e2e` is exactly one `__call__`."""

    def __init__(self, lib: Library, chunk_samples: int = 1 << 22, device: Optional[torch.device] = None,
                 flags: int = native.SB_STEP_LOSS | native.SB_STEP_GRAD):
        self.lib, self.flags = lib, flags
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.chunk = int(chunk_samples)
        d = lib.dim
        self._bufs = [(torch.empty(self.chunk, d, dtype=torch.float32, device=self.dev),
                       torch.empty(self.chunk, d, dtype=torch.float32, device=self.dev)) for _ in range(2)]
        self._streams = [torch.cuda.Stream(self.dev) for _ in range(2)]
        n_out = lib.step_out_len(flags)
        self._outs = [torch.empty(n_out, dtype=torch.float64, device=self.dev) for _ in range(2)]
        self._accs = [torch.zeros(n_out, dtype=torch.float64, device=self.dev) for _ in range(2)]  # one per stream
        self.h2d_bytes = 0

    def __call__(self, x_host: torch.Tensor, dx_host: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        if x_host.is_cuda or dx_host.is_cuda:
            raise ValueError("HostStreamedStep takes host tensors (pinned for asynchronous copies)")
        n = x_host.shape[0]
        main = torch.cuda.current_stream(self.dev)
        self.h2d_bytes = 0
        for s, acc in zip(self._streams, self._accs):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                acc.zero_()
        for ci, start in enumerate(range(0, n, self.chunk)):
            stop = min(start + self.chunk, n)
            m = stop - start
            slot = ci % 2
            bx, bdx = self._bufs[slot]
            with torch.cuda.stream(self._streams[slot]):
                bx[:m].copy_(x_host[start:stop], non_blocking=True)
                bdx[:m].copy_(dx_host[start:stop], non_blocking=True)
                native.train_step(bx[:m], bdx[:m], w, self.lib, self.flags, out=self._outs[slot])
                self._accs[slot] += self._outs[slot]
            self.h2d_bytes += 2 * m * self.lib.dim * 4
        for s in self._streams:
            main.wait_stream(s)
        return self._accs[0] + self._accs[1]
