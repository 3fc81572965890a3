"""Per-rank timeline of the one-launch fit step over several GPUs (torchrun): when the rank's totals are ready, how
long the in-kernel peer exchange takes (push + flags + wait for the slowest rank), end of the epilogue.
%globaltimer is per GPU, so only differences within a rank are meaningful."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
import torch, torch.distributed as dist
from sindy_b200 import native
from sindy_b200.dist import FitStepper
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000 // world
lib = native.Library(3, 5)
g = torch.Generator(device=dev).manual_seed(1 + rank)
x = torch.rand(n, 3, device=dev, generator=g) * 2 - 1
dx = torch.randn(n, 3, device=dev, generator=g)
torch.manual_seed(0)
xi = torch.randn(3, 56).to(dev); mask = torch.ones_like(xi)
st = FitStepper(lib, x, dx, "adam", lr=1e-4, use_graph=False)
st.load(xi, mask)
for _ in range(5): st.step()
R = 8
bufs = [torch.zeros(16 * 592, dtype=torch.int64, device=dev) for _ in range(R)]
gph = torch.cuda.CUDAGraph()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side): st.step()
torch.cuda.current_stream().wait_stream(side)
with torch.cuda.graph(gph):
    for r in range(R):
        native.load().sb_debug_trace(bufs[r].data_ptr()); st.step()
native.load().sb_debug_trace(None)
dist.barrier(); torch.cuda.synchronize()
gph.replay(); torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); gph.replay(); b.record(); torch.cuda.synchronize()
per = 1e3 * a.elapsed_time(b) / R
lines = [f"rank {rank}: {per:.2f} us per iteration"]
prev_end = None
for r in range(2, R):
    t = bufs[r].view(-1, 16).cpu(); t = t[t[:, 0] > 0]
    t0 = int(t[:, 0].min()); last = int(t[:, 6].argmax())
    us = lambda v: (int(v) - t0) / 1e3
    gap = (t0 - prev_end) / 1e3 if prev_end else float("nan")
    lines.append(f"  launch {r}: gap {gap:5.2f} | loop end max {us(t[:,2].max()):7.2f} | totals {us(t[last,4]):7.2f} | exchange done {us(t[last,5]):7.2f}"
                 f" (+{us(t[last,5]) - us(t[last,4]):5.2f}) | end {us(t[last,6]):7.2f}")
    prev_end = int(t[last, 6])
out = [None] * world
dist.all_gather_object(out, "\n".join(lines))
if rank == 0:
    print("\n".join(out))
torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
