#!/usr/bin/env python
"""bench.py — SINDy train-step throughput on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[4], SURVEY.md §8d "C5"): synthetic library fit, N = 1e8 samples in total,
d = 3, polynomial degree 5 (K = 56), X ~ U(-1,1)^3 generated on the device (seed 1234 + rank),
dX = Θ(X)Ξ*ᵀ + 0.01·randn with the Lorenz-form Ξ*, Ξ initialised randn(3,56) with torch.manual_seed(0).
A "step" is one iteration of the reference's Adam loop without sym-reg (`train.py:512-530`): loss = MSE + w·‖Ξ‖₁
and dL/dΞ over ALL samples, then the Adam update of Ξ — ONE launch of the fused kernel per rank (sb_fit_step): its
last block all-reduces the 170 packed fp64 sums over NVLink peer memory, writes loss and gradient, advances Ξ and
packs the next Ξ⊙mask into the constant bank. Samples are sharded over the ranks (total work fixed: strong scaling). Inputs (2.4 GB) exceed the 126 MB L2, so no flush is needed between steps.

Printed JSON (rank 0, one line): value = whole-job samples/s with inputs resident in HBM; e2e = the same step
through HostStreamedStep with x/dx in pinned HOST memory (H2D of every sample inside the timed region) and a
device->host read of the loss; roofline = the fused kernel against the FP32 FMA peak measured live by
sb_fp32_peak (not in MEASURED_PEAKS.json) and, as roofline_hbm, against the measured HBM copy bandwidth;
cpu_baseline = the oracle port of the reference closure (torch CPU, all host threads) on a bounded sample.
`--impl reference` times that CPU port alone (the reference is a Python package that cannot travel to the box).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "symmetry-ode-discovery_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

D, P, K = 3, 5, 56
FLOP_PER_SAMPLE = (K - 1 - D) + 2 * K * D + 3 * D + 2 * K * D   # 733: Θ, prediction, residual², r⊗Θ
BYTES_PER_SAMPLE = 8 * D                                        # x and dx, fp32, read once
METRIC = "SINDy train-step samples/s (fused Θ+symreg)"


def truth_xi(device):
    Xi = torch.zeros(D, K, device=device)
    Xi[0, 1], Xi[0, 2] = -10.0, 10.0
    Xi[1, 1], Xi[1, 2], Xi[1, 6] = 2.8, -1.0, -1.0
    Xi[2, 5], Xi[2, 3] = 1.0, -8.0 / 3.0
    return Xi


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (recipe in B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.file = None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.file.read().splitlines():
            parts = [q.strip() for q in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        self.file.close()
        os.unlink(self.file.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_closure_rate(n_cpu, budget_s=12.0, min_reps=2, max_reps=40):
    """The oracle port of the reference closure (regressor(x) + MSELoss + L1 + backward, `train.py:645-690`,
    with the cat-of-products Θ of `sindy.py:7-30` continued to degree 5) on all host threads."""
    from oracle import sindy_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(n_cpu, D, generator=g) * 2 - 1
    dx = O.torch_theta(x, P) @ truth_xi("cpu").T + 0.01 * torch.randn(n_cpu, D, generator=g)
    torch.manual_seed(0)
    Xi = torch.randn(D, K)
    mask = torch.ones(D, K)
    O.torch_closure(x, dx, Xi, mask, P)  # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < max_reps and (len(times) < min_reps or time.perf_counter() - t_start < budget_s):
        t0 = time.perf_counter()
        O.torch_closure(x, dx, Xi, mask, P)
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return n_cpu / med, torch.get_num_threads(), len(times), med


def run_reference(args, rank):
    """`--impl reference`: the reference's CPU path (oracle port: the reference is Python and only importable in
    the build container). Rank 0 alone works; other ranks exit."""
    if rank != 0:
        return
    from oracle import sindy_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n_ref = args.cpu_samples
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(n_ref, D, generator=g) * 2 - 1
    dx = O.torch_theta(x, P) @ truth_xi("cpu").T + 0.01 * torch.randn(n_ref, D, generator=g)
    torch.manual_seed(0)
    Xi = torch.randn(D, K)
    mask = torch.ones(D, K)
    for _ in range(args.warmup):
        _, grad = O.torch_closure(x, dx, Xi, mask, P)
        Xi = Xi - 1e-3 * grad
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, grad = O.torch_closure(x, dx, Xi, mask, P)
        Xi = Xi - 1e-3 * grad
    el = time.perf_counter() - t0
    value = n_ref * args.steps / el
    sample = f"{n_ref} of the 1e8 samples per step (Θ is materialised: N×56 fp32 plus autograd copies)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C5 synthetic library fit: d=3, poly degree 5 (K=56), MSE+L1 closure, loss+grad",
                   "samples_total": int(1e8), "samples_per_step_timed": n_ref},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=float, default=1e8, help="total samples over all ranks")
    ap.add_argument("--cpu-samples", type=int, default=1_000_000, help="bounded sample of the CPU baseline")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-graph", action="store_true", help="do not capture the multi-GPU step in a CUDA graph")
    ap.add_argument("--no-peer", action="store_true", help="multi-GPU: NCCL all-reduce instead of the in-kernel peer all-reduce")
    ap.add_argument("--skip-extras", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    from sindy_b200 import native
    from sindy_b200.dist import FitStepper, HostStreamedStep, ShardedTrainStep, bind_to_gpu_numa_node, mse_from_sums

    native.load()  # fails loudly if the CUDA library is missing: there is no fallback path
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = native.Library(D, P)
    assert lib.K == K

    # ---- synthetic shard, generated on the device ----
    n_total = int(args.samples)
    n_local = n_total // world + (1 if rank < n_total % world else 0)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.rand(n_local, D, device=dev, generator=gen) * 2 - 1
    dx = native.forward(x, truth_xi(dev), lib)
    dx.add_(torch.randn(n_local, D, device=dev, generator=gen), alpha=0.01)
    torch.manual_seed(0)
    Xi0 = torch.randn(D, K).to(dev)
    mask = torch.ones(D, K, device=dev)
    w_l1 = 0.0  # w_sindy_reg of the C5 config; the L1 term is still evaluated like the reference does
    flags = native.SB_STEP_LOSS | native.SB_STEP_GRAD
    variant = native.train_step_variant(lib, flags)

    # One step = one iteration of the reference's Adam loop without sym-reg (`train.py:512-530`): forward, MSE + L1,
    # backward, optimizer.step() — ONE launch of the fused kernel per rank (sb_fit_step): its last block all-reduces
    # the sums over NVLink peer memory (N > 1), evaluates loss and gradient, applies Adam to Ξ in place and packs Ξ⊙mask
    # into the constant bank for the next launch. Replayed as a CUDA graph. `--no-peer`: the 3-launch NCCL path.
    legacy = args.no_peer or args.no_graph
    kern_events = []
    fallback_note = None
    if not legacy:
        try:
            stepper = FitStepper(lib, x, dx, "adam", lr=1e-3, w_l1=w_l1, use_graph=True)
        except RuntimeError as exc:   # no peer-mapped symmetric memory on this box: the NCCL path still measures
            legacy, fallback_note = True, f"one-launch step unavailable ({exc}); NCCL path"
    if not legacy:
        stepper.load(Xi0, mask)
        collective = "none" if world == 1 else "in-kernel peer all-reduce over NVLink (sb_fit_step)"
        launches_per_step = 1
        use_graph = True

        def one_step(record=False):
            return stepper.step()
    else:
        use_graph = (world > 1) and not args.no_graph
        stepper = ShardedTrainStep(lib, x, dx, flags=flags, use_graph=use_graph, sgd_lr=1e-3, use_peer=not args.no_peer)
        collective = "none" if world == 1 else ("in-kernel peer all-reduce over NVLink (sb_closure_peer)"
                                                if stepper.peer is not None else "NCCL all-reduce + epilogue launch")
        stepper.step(Xi0, mask, w_l1)  # loads the static parameters (and captures the graph when enabled)
        stepper.xi.copy_(Xi0)
        launches_per_step = 2 if (world == 1 or stepper.peer is not None) else 3

        def one_step(record=False):
            if record and not use_graph:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                packed, loss_t, grad_t = stepper._buffers(dev)
                if world == 1:
                    native.closure(x, dx, stepper.xi, stepper.mask, lib, w_l1, packed=packed, loss=loss_t, grad=grad_t)
                    e1.record()
                else:
                    native.train_step(x, dx, stepper.xi * stepper.mask, lib, flags, out=packed)
                    e1.record()
                    dist.all_reduce(packed)
                    native.step_epilogue(packed, stepper.xi, stepper.mask, lib, w_l1, loss=loss_t, grad=grad_t)
                kern_events.append((e0, e1))
                stepper.xi.add_(grad_t, alpha=-1e-3)
                return loss_t
            loss_t, _ = stepper.step(None, None, w_l1)
            return loss_t

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the one-launch step runs in groups of `unroll` iterations per graph replay (no host work between iterations)
    unroll = next(u for u in (25, 20, 16, 10, 8, 5, 4, 2, 1) if args.steps % u == 0)

    def run_steps(k):
        if not legacy:
            return stepper.run(k, unroll)
        for _ in range(k):
            out = one_step(record=True)
        return out

    if not legacy:
        stepper.run(unroll, unroll)      # capture outside the timed region
        stepper.load(Xi0, mask, reset_state=True)
    for _ in range(args.warmup):
        one_step()
    barrier()
    launches0 = native.kernel_launches()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        # the ranks leave the host barrier up to a few hundred us apart (a step is 0.2 ms at 8 GPUs) and the first
        # exchange would charge that skew to the timed steps: one more untimed step lets the in-kernel exchange align
        # the GPUs on the device; ev0 is then recorded on every rank at the end of the same exchange
        one_step()
    ev0.record()
    loss = run_steps(args.steps)
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = native.kernel_launches() - launches0
    if use_graph:   # kernels launched by the replayed graph are not seen by the library's launch counter
        launches = launches_per_step * args.steps
    elapsed_ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_ms = float(elapsed_ms)
    value = n_total * args.steps / (elapsed_ms * 1e-3)
    final_loss = float(loss)

    # ---- kernel-only duration (roofline numerator) ----
    if not legacy:
        kern_ms = elapsed_ms / args.steps
        kern_how = ("CUDA events around the K timed steps: a step IS one launch of the fused kernel (sample loop, "
                    "reductions, peer all-reduce, loss/gradient, Adam update), so launch gaps are included")
    elif kern_events:
        kern_ms = sum(a.elapsed_time(b) for a, b in kern_events) / len(kern_events)
        kern_how = "CUDA events around the library call (pack_w + fused kernel) inside every timed step"
    else:
        out = torch.empty(lib.step_out_len(flags), dtype=torch.float64, device=dev)
        wm = stepper.xi * mask
        native.train_step(x, dx, wm, lib, flags, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            native.train_step(x, dx, wm, lib, flags, out=out)
        e1.record()
        torch.cuda.synchronize()
        kern_ms = e0.elapsed_time(e1) / args.steps
        kern_how = "CUDA events over back-to-back launches right after the timed region (steps replay a CUDA graph)"
    kt = torch.tensor([kern_ms], device=dev)
    if world > 1:
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    kern_ms = float(kt)

    # ---- e2e: host buffers, H2D inside the timed region, loss read back ----
    numa_node = bind_to_gpu_numa_node(dev) if world > 1 else None   # node-local pinned buffers (one process per GPU)
    host_step = HostStreamedStep(lib, chunk_samples=1 << 22, device=dev, flags=flags)
    xh = torch.empty(n_local, D, dtype=torch.float32, pin_memory=True)
    dxh = torch.empty(n_local, D, dtype=torch.float32, pin_memory=True)
    xh.copy_(x)
    dxh.copy_(dx)
    Xi_e = Xi0.clone()

    def e2e_step():
        nonlocal Xi_e
        packed = host_step(xh, dxh, Xi_e * mask)
        if world > 1:
            dist.all_reduce(packed)
        loss_e, grad = mse_from_sums(packed, lib, Xi_e, mask)
        Xi_e = Xi_e - 1e-3 * grad
        return float(loss_e)  # device -> host read of the step's result

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_total * args.e2e_steps / float(e2e_s)
    h2d_bytes = host_step.h2d_bytes
    del xh, dxh

    # ---- sharded RK4 rollout of the same config (10^6 ICs x 2000 steps, every 10th state stored): the ICs split
    # over the ranks, no collective on the data path; time = max over ranks ----
    rollout_sharded = None
    if world > 1 and not args.skip_extras:
        n_ics = 10 ** 6 // world + (1 if rank < 10 ** 6 % world else 0)
        g2 = torch.Generator(device=dev).manual_seed(4321 + rank)
        x0 = torch.rand(n_ics, D, device=dev, generator=g2) * 2 - 1
        Xi_t = truth_xi(dev)
        native.rollout(x0, Xi_t, lib, 0.002, 2000, 10, "rk4")
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(3):
            native.rollout(x0, Xi_t, lib, 0.002, 2000, 10, "rk4")
        r1.record()
        torch.cuda.synchronize()
        rt = torch.tensor([r0.elapsed_time(r1) / 3], device=dev)
        dist.all_reduce(rt, op=dist.ReduceOp.MAX)
        rollout_sharded = {"ic_steps_per_s": 2e9 / (float(rt) * 1e-3), "ms": float(rt), "ics_total": 10 ** 6,
                           "steps": 2000, "dtype": "f32", "collective": "none (ICs are independent)"}

    if rank != 0:
        # every collective of this run has completed on all ranks (the last one is the e2e all-reduce); leave
        # without NCCL teardown: destroy_process_group() after CUDA-graph capture of collectives was seen to hang
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)

    # ---- roofline (rank 0) ----
    peaks, peak_src = measured_peaks()
    fp32_peaks = {name: native.fp32_peak(v, 4096) for v, name in ((0, "ffma"), (1, "ffma2"), (2, "ffma2_const"))}
    fp32_peak = max(fp32_peaks.values())
    n_kernel = n_local  # rank 0's shard (largest)
    ach_tflops = FLOP_PER_SAMPLE * n_kernel / (kern_ms * 1e-3) / 1e12
    ach_gbs = BYTES_PER_SAMPLE * n_kernel / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if int(tj.get("samples_per_launch", -1)) == n_kernel:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "fp32", "achieved": ach_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": ach_tflops / fp32_peak, "traffic": traffic, "kernel": variant, "kernel_ms": kern_ms,
                "flop_per_sample": FLOP_PER_SAMPLE, "samples_per_launch": n_kernel,
                "peak_source": "measured live: sb_fp32_peak, best of " + json.dumps(fp32_peaks),
                "peak_nominal": 148 * 128 * 2 * 1.965e9 / 1e12, "timing": kern_how}
    roofline_hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach_gbs / peaks["hbm_gbs"], "traffic": traffic, "bytes_per_sample": BYTES_PER_SAMPLE,
                    "peak_source": peak_src}

    result = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C5 synthetic library fit: d=3, poly degree 5 (K=56), MSE+L1 closure, loss+grad",
                   "step": ("one Adam iteration (train.py:512-530): loss, gradient, update in ONE launch (sb_fit_step)"
                            if not legacy else "closure + SGD update (pack_w, fused kernel[, all-reduce, epilogue], axpy)"),
                   "samples_total": n_total, "samples_per_gpu": n_local, "symreg": "none",
                   "l2": "inputs (24 B/sample, %.2f GB per GPU) larger than L2; no flush" % (24 * n_local / 1e9),
                   "parallelism": f"sample-sharded x{world}, one all-reduce of {2 + D * K} fp64 sums per step",
                   "collective": collective, "fallback": fallback_note,
                   "cuda_graph": use_graph, "iterations_per_graph_replay": (unroll if not legacy else 1),
                   "final_loss": final_loss},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 8,
                "steps": args.e2e_steps,
                "how": "HostStreamedStep: pinned host x/dx -> 2 staging buffers on 2 streams -> fused kernel per "
                       "chunk; loss read back with .item()", "numa_node": numa_node},
        "roofline": roofline, "roofline_hbm": roofline_hbm,
    }

    if rollout_sharded is not None:
        result["extra"] = {"rk4_rollout_sharded": rollout_sharded}
    if world == 1:
        rate, cores, reps, med = cpu_closure_rate(args.cpu_samples)
        result["cpu_baseline"] = {
            "value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{args.cpu_samples} of the 1e8 samples, median of {reps} closures ({med * 1e3:.0f} ms each)"}
        if not args.skip_extras:
            try:
                result["extra"] = extras(native, dev, peaks, fp32_peak)
            except Exception as exc:   # secondary measurements must never cost the headline line
                result["extra"] = {"error": repr(exc)}
    print(json.dumps(result))
    sys.stdout.flush()
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


def extras(native, dev, peaks, fp32_peak):
    """Secondary measurements of the same hot path (not the headline): the HBM-bound d=2 library and the
    batched RK4 rollout of C5 (10^6 ICs x 2000 steps, every 10th state stored)."""
    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    out = {}
    gen = torch.Generator(device=dev).manual_seed(99)
    n = 10 ** 8
    for (d, p) in ((2, 2), (2, 3), (3, 3)):
        lib = native.Library(d, p)
        x = torch.rand(n, d, device=dev, generator=gen) * 2 - 1
        dx = torch.randn(n, d, device=dev, generator=gen)
        W = torch.randn(d, lib.K, device=dev, generator=gen)
        o = torch.empty(lib.step_out_len(3), dtype=torch.float64, device=dev)
        ms = timed(lambda: native.train_step(x, dx, W, lib, 3, out=o))
        gbs = 8 * d * n / (ms * 1e-3) / 1e9
        out[f"train_step_d{d}_p{p}_K{lib.K}"] = {"samples_per_s": n / (ms * 1e-3), "ms": ms, "hbm_gbs": gbs,
                                                 "hbm_frac": gbs / peaks["hbm_gbs"]}
        del x, dx
    # the C5 step with the linear so(3) Lie-derivative regulariser (weight 0.1, `train.py:503-507`): closure kernel +
    # power-sum Gram kernel (second pass over x) + K×K algebra, replayed as one CUDA graph; and the STLSQ data pass
    from sindy_b200.dist import ShardedTrainStep
    lib = native.Library(D, P)
    x = torch.rand(n, D, device=dev, generator=gen) * 2 - 1
    dx = native.forward(x, truth_xi(dev), lib)
    so3 = torch.zeros(3, 3, 3)
    q = 0
    for i in range(3):
        for j in range(i):
            so3[q, i, j], so3[q, j, i] = 1.0, -1.0
            q += 1
    W = torch.randn(D, K, device=dev, generator=gen)
    st = ShardedTrainStep(lib, x, dx, sym_gens=list(so3), w_sym=0.1, use_graph=True, sgd_lr=1e-6)
    st.step(W, torch.ones_like(W), 0.0)
    ms = timed(lambda: st.step(None, None, 0.0))
    flop_sym = FLOP_PER_SAMPLE + (9 + 65) + 2 * 286   # + leading powers, trailing table, one FMA per power sum
    out["train_step_C5_with_so3_symreg"] = {"samples_per_s": n / (ms * 1e-3), "ms": ms, "flop_per_sample": flop_sym,
                                            "fp32_tflops": flop_sym * n / (ms * 1e-3) / 1e12,
                                            "fp32_frac": flop_sym * n / (ms * 1e-3) / 1e12 / fp32_peak,
                                            "bytes_per_sample": 36}
    # the same objective with the Gram matrix of the fixed data set formed ONCE per fit (moment kernel, timed separately)
    # and the regulariser + gradient evaluated as a quadratic form in the epilogue of the one-launch iteration
    from sindy_b200.dist import FitStepper
    from sindy_b200 import symreg
    ms_gram = timed(lambda: symreg.gram(x, lib))
    fs = FitStepper(lib, x, dx, "adam", lr=1e-6, sym_gens=list(so3), w_sym=0.1)
    fs.load(W, torch.ones_like(W))
    fs.run(10, 10)
    ms = timed(lambda: fs.run(10, 10)) / 10
    out["fit_step_C5_with_so3_symreg_cached_gram"] = {
        "samples_per_s": n / (ms * 1e-3), "ms": ms, "gram_once_per_fit_ms": ms_gram,
        "note": "one launch per Adam iteration; w_sym*w'Hw and 2Hw evaluated by the kernel's last block, H from the Gram of the data set"}
    del fs
    o = torch.empty(lib.step_out_len(12), dtype=torch.float64, device=dev)
    ms = timed(lambda: native.train_step(x, dx, None, lib, 12, out=o))
    out["stlsq_data_pass_C5_gram_and_b"] = {"samples_per_s": n / (ms * 1e-3), "ms": ms,
                                            "variant": native.train_step_variant(lib, 12)}
    # the whole STLSQ solve of `sindy.py:318-324` (mask reset, up to 5 threshold iterations) on the Lorenz-form data:
    # ONE data pass (Gram by the moment kernel + ΘᵀẊ by the fused kernel + Σẋ²), then K×K fp64 solves per iteration
    import sindy
    reg = sindy.SINDyRegression(D, P, False, False, threshold=0.1, device=str(dev), constrain_constant=True)
    dxn = dx + 0.01 * torch.randn(n, D, device=dev, generator=gen)
    ms = timed(lambda: sindy.solve_SINDy(reg, x, dxn, 0.0, 0.1), reps=3)
    support_ok = bool(torch.equal(reg.mask.bool(), truth_xi(dev) != 0))
    coef_err = float(((reg.Xi.detach() * reg.mask) - truth_xi(dev)).abs().max() / truth_xi(dev).abs().max())
    out["stlsq_solve_SINDy_C5"] = {"samples_per_s": n / (ms * 1e-3), "ms": ms, "recovered_support_is_truth": support_ok,
                                   "max_coef_err_rel": coef_err}
    del x, dx, dxn
    x0 = torch.rand(10 ** 6, D, device=dev, generator=gen) * 2 - 1
    Xi = truth_xi(dev)
    ms = timed(lambda: native.rollout(x0, Xi, lib, 0.002, 2000, 10, "rk4"), reps=3)
    flops = 4 * ((K - 1 - D) + 2 * K * D) + 12 * D
    out["rk4_rollout_f32_1e6ics_2000steps"] = {"ic_steps_per_s": 2e9 / (ms * 1e-3), "ms": ms,
                                               "fp32_tflops": flops * 2e9 / (ms * 1e-3) / 1e12,
                                               "fp32_frac": flops * 2e9 / (ms * 1e-3) / 1e12 / fp32_peak}
    x0d, Xid = x0[: 10 ** 5].double(), Xi.double()
    ms = timed(lambda: native.rollout(x0d, Xid, lib, 0.002, 2000, 10, "rk4", record_dx=True), reps=3)
    out["rk4_rollout_f64_1e5ics_2000steps"] = {"ic_steps_per_s": 2e8 / (ms * 1e-3), "ms": ms}
    out["cpu_port_side_by_side"] = cpu_side_by_side()
    return out


def cpu_side_by_side():
    """The other two rows of the path on the host cores, bounded samples (part of the CPU-baseline reporting, SURVEY
    §8d): the oracle port of `solve_ode_batch` (float64 RK4, K = 56) and of one `solve_SINDy_one_step` (fp32 lstsq of
    the stacked [Θ; wI] like the reference). The port's full STLSQ does NOT recover the planted support at this size:
    LAPACK's default rank tolerance eps·rows (0.048 at 4e5 rows) exceeds σ_min/σ_max of Θ (0.019) — see DESIGN.md §2."""
    import numpy as np
    from oracle import sindy_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    Xi = truth_xi("cpu").double().numpy()
    rng = np.random.default_rng(0)
    x0 = rng.uniform(-1, 1, (10_000, D))
    t0 = time.perf_counter()
    O.solve_ode_batch(O.library_rhs(Xi, P), x0, dt=0.002, num_steps=50)
    rk = 10_000 * 49 / (time.perf_counter() - t0)
    n = 400_000
    x = rng.uniform(-1, 1, (n, D)).astype(np.float32)
    y = (O.theta(x, P) @ Xi.T + 0.01 * rng.standard_normal((n, D))).astype(np.float32)
    t0 = time.perf_counter()
    O.stlsq_one_step(x, y, np.ones((D, K), dtype=np.float32), 0.0, 0.1, P)
    st = n / (time.perf_counter() - t0)
    return {"rk4_f64_ic_steps_per_s": rk, "rk4_sample": "1e4 ICs x 50 steps", "stlsq_one_step_samples_per_s": st,
            "stlsq_sample": "4e5 samples, K = 56", "cores": torch.get_num_threads(), "kind": "port"}


if __name__ == "__main__":
    main()
