"""Run under torchrun (one rank per GPU): data-parallel invariance of the CUDA path.
W ranks x (B/W) rows must match one rank x B rows (only summation order differs), for both gradient exchanges:
    (default)  tape + CUDA-graph path with the NCCL allreduce captured in the graph
    --peer     fused device step with the in-kernel NVLink peer-memory exchange (tp_xchg_*)
    --wide     a wide MLP (the tcgen05 kernel plan, step_wide.cu): NCCL allreduce of the gradient arena inside the plan, or with
               --peer the two-phase peer-memory exchange fused with the optimizer; Adam eps = 0.1 so that the plain 1e-4 bound
               applies to the parameters (default-eps Adam amplifies summation-order noise)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from taper_b200 import host
from taper_b200.dp import shard_permutation

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
host.set_device(local)
WIDE = "--wide" in sys.argv
spec = "linear:784:520,relu,linear:520:264,relu,linear:264:10" if WIDE else host.MLP_784_128_10
b, steps, n = (1000, 5, 16000) if WIDE else (64, 6, 4096)
EPS = 0.1 if WIDE else 1e-8
rng = np.random.default_rng(0)
X = rng.random((n, 784)).astype(np.float32); Y = rng.integers(0, 10, n).astype(np.float32)
perm = rng.permutation(n)

model = host.Model(spec, seed=0)                  # same seed on every rank; broadcast anyway
tr = host.Trainer(model, "adam", lr=1e-3, eps=EPS)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.frombuffer(bytearray(host.nccl_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(uid, 0)
tr.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
tr.broadcast_params(0)
PEER = "--peer" in sys.argv
if PEER:
    tr.peer_exchange_init(dist)
elif not WIDE:
    tr.set_use_fused(False)
tr.load_dataset(X, Y, np.resize(shard_permutation(perm, rank, world, b), n))    # one entry per sample (the shard's order, repeated)
local_losses = []
for s in range(steps):                            # eager, capture (with the allreduce inside the graph), replays
    tr.step_resident(b)
    local_losses.append(tr.fetch()[0])
params = [model.get_param(i) for i in range(model.num_params())]
assert tr.fused_steps() == (steps if (PEER or WIDE) else 0), tr.fused_steps()
assert tr.fused_kind() == (2 if WIDE else (1 if PEER else 0)), tr.fused_kind()

# single-replica reference on the global batch, same process, no communicator
ref_model = host.Model(spec, seed=0)
ref_tr = host.Trainer(ref_model, "adam", lr=1e-3, eps=EPS)
ref_tr.load_dataset(X, Y, perm.astype(np.uint32))
ref_losses = []
for s in range(steps):
    ref_tr.step_resident(b * world)
    ref_losses.append(ref_tr.fetch()[0])
t = torch.tensor(local_losses, dtype=torch.float64, device="cuda")
dist.all_reduce(t)
mean_losses = (t / world).cpu().numpy()
ok = True
for s in range(steps):
    if abs(mean_losses[s] - ref_losses[s]) > 1e-4 * abs(ref_losses[s]):
        ok = False; print(f"rank {rank}: step {s} loss {mean_losses[s]} vs {ref_losses[s]}")
for i, p in enumerate(params):
    r = ref_model.get_param(i)
    err = np.max(np.abs(p - r)); scale = max(np.max(np.abs(r)), 1e-6)
    if WIDE and r.ndim == 1 and i > 0:       # a (zero-initialised) bias is judged on its layer's weight scale, see bench.py dp_check
        scale = max(scale, np.max(np.abs(ref_model.get_param(i - 1))))
    if err > 1e-4 * scale + (0.0 if WIDE else 0.03 * 1e-3):
        ok = False; print(f"rank {rank}: param {i} err {err:.3e} scale {scale:.3e}")
# replicas identical across ranks
flat = torch.from_numpy(np.concatenate([p.reshape(-1) for p in params])).cuda()
ref0 = flat.clone(); dist.broadcast(ref0, 0)
if not torch.equal(flat, ref0):
    ok = False; print(f"rank {rank}: replica diverged from rank 0")
flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DP_CHECK", "OK" if flag.item() == 1 else "FAILED", f"world={world} graph_replays={tr.graph_replays()} fused_kind={tr.fused_kind()}", flush=True)
code = 0 if flag.item() == 1 else 1
dist.barrier()
dist.destroy_process_group()
sys.stdout.flush(); sys.stderr.flush()
os._exit(code)          # no interpreter finalisation: two NCCL instances + CUDA IPC mappings tearing down under it have ended a rank with SIGSEGV
