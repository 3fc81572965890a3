"""Multi-GPU check of the in-kernel peer all-reduce (run with torchrun, 2+ GPUs):
global loss/grad from sb_closure_peer == NCCL all-reduce path == single-process oracle sums, bitwise equal across
ranks, stable over repeated (and CUDA-graph replayed) steps, with uneven shards."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "symmetry-ode-discovery_b200")]
import numpy as np, torch, torch.distributed as dist
from sindy_b200 import native
from sindy_b200.dist import FitStepper, ShardedTrainStep

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ok = True
for (d, p) in ((3, 5), (2, 3)):
    lib = native.Library(d, p)
    rng = np.random.default_rng(0)
    n = 400_003
    x = rng.uniform(-1, 1, (n, d)).astype(np.float32); dx = rng.standard_normal((n, d)).astype(np.float32)
    Xi = rng.standard_normal((d, lib.K)).astype(np.float32); mask = (rng.random((d, lib.K)) > 0.2).astype(np.float32)
    cuts = np.linspace(0, n, world + 1).astype(int); cuts[1:-1] += 17 * np.arange(1, world)   # uneven shards
    lo, hi = cuts[rank], cuts[rank + 1]
    xs, dxs = torch.from_numpy(x[lo:hi]).to(dev), torch.from_numpy(dx[lo:hi]).to(dev)
    Xit, mt = torch.from_numpy(Xi).to(dev), torch.from_numpy(mask).to(dev)
    peer = ShardedTrainStep(lib, xs, dxs)
    nccl = ShardedTrainStep(lib, xs, dxs, use_peer=False)
    if peer.peer is None:
        print(f"rank {rank}: peer exchange unavailable"); ok = False; break
    l_p, g_p = peer.step(Xit, mt, 0.01); l_p = float(l_p); g_p = g_p.clone()
    l_n, g_n = nccl.step(Xit, mt, 0.01)
    full = native.closure(torch.from_numpy(x).to(dev), torch.from_numpy(dx).to(dev), Xit, mt, lib, 0.01)
    e1 = abs(l_p - float(l_n)) / abs(float(l_n)); e2 = float((g_p - g_n).abs().max() / g_n.abs().max())
    e3 = abs(l_p - float(full[0])) / abs(float(full[0])); e4 = float((g_p - full[1]).abs().max() / full[1].abs().max())
    gathered = [torch.zeros_like(g_p) for _ in range(world)]
    dist.all_gather(gathered, g_p)
    same = all(torch.equal(gathered[0], q) for q in gathered)
    # repeated steps (epoch/parity logic) and graph replay
    for _ in range(5):
        l_r, g_r = peer.step(Xit, mt, 0.01)
    rep = float(l_r) == l_p and torch.equal(g_r, g_p)
    gstep = ShardedTrainStep(lib, xs, dxs, use_graph=True)
    for _ in range(4):
        l_g, g_g = gstep.step(Xit, mt, 0.01)
    gr = float(l_g) == l_p and torch.equal(g_g, g_p)
    # one-launch Adam iterations (sb_fit_step) over the shards == the same iterations on the whole data on one GPU;
    # parameters bitwise equal across ranks after every step
    fit = FitStepper(lib, xs, dxs, "adam", lr=1e-2, w_l1=1e-3)
    fit.load(Xit, mt)
    losses = [float(fit.step()) for _ in range(6)]
    xg = [torch.zeros_like(fit.xi) for _ in range(world)]
    dist.all_gather(xg, fit.xi)
    fit_same = all(torch.equal(xg[0], q) for q in xg)
    xi1 = Xit.clone().contiguous(); st1 = native.fit_state(lib, dev); l1 = []
    xf, dxf = torch.from_numpy(x).to(dev), torch.from_numpy(dx).to(dev)
    for it in range(6):
        l, _, _ = native.fit_step(xf, dxf, xi1, mt, lib, "adam", 1e-2, w_l1=1e-3, state=st1, w_resident=False)
        l1.append(float(l))
    e5 = max(abs(a - b) / abs(b) for a, b in zip(losses, l1))
    e6 = float((fit.xi - xi1).abs().max() / xi1.abs().max())
    fit_ok = fit_same and e5 < 1e-5 and e6 < 1e-5
    good = e1 < 1e-6 and e2 < 1e-6 and e3 < 1e-5 and e4 < 1e-5 and same and rep and gr and fit_ok
    ok = ok and good
    if rank == 0:
        print(f"d={d} p={p}: vs nccl {e1:.1e}/{e2:.1e} vs single {e3:.1e}/{e4:.1e} ranks-bitwise-equal={same} repeat={rep} graph={gr} fit: ranks-equal={fit_same} loss {e5:.1e} xi {e6:.1e} -> {'OK' if good else 'FAIL'}")
t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
# a lost peer: rank 0 launches one more iteration that nobody else joins. The kernel must give up after
# $SB_PEER_TIMEOUT_MS, report NaN, leave the parameters / Adam state untouched and raise through FitStepper.check()
lost_ok = True
if rank == 0:
    os.environ["SB_PEER_TIMEOUT_MS"] = "300"
    xi_before, st_before = fit.xi.clone(), fit.state.clone()
    fit._use_graph = False
    l_lost = float(fit.step())
    raised = False
    try:
        fit.check()
    except native.SindyB200Error:
        raised = True
    lost_ok = (l_lost != l_lost) and torch.equal(fit.xi, xi_before) and torch.equal(fit.state, st_before) and raised
    print(f"lost peer: loss={l_lost} parameters-untouched={torch.equal(fit.xi, xi_before)} check-raised={raised} -> {'OK' if lost_ok else 'FAIL'}")
    print("PEER TEST", "PASSED" if (int(t) == 1 and lost_ok) else "FAILED")
    torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0 if (int(t) == 1 and lost_ok) else 1)
torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0 if int(t) == 1 else 1)
