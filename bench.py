#!/usr/bin/env python
"""bench.py — MNIST samples/sec (fwd + bwd + optimizer step) of the tape-evaluation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg4|cfg5|cfg2|...] [--impl ours|reference]

A "step" is one iteration of the reference's train_epoch loop body (src/train.rs:106-138): Tape::reset, forward,
cross-entropy, accuracy, backward, [gradient allreduce], optimizer step, zero_grad — on one synthetic MNIST-shaped batch.
The headline workload is BASELINE.json configs[3] (MLP 784-1024-1024-10, batch 1024 per GPU, Adam): the configuration the
metric "samples/sec at 1/2/4/8 B200" is quoted on.  configs[4] (5-conv CNN, batch 1024/GPU, AdamW + StepLR) and configs[1]
(MLP 784-128-10, batch 512, Adam) are measured in the same run and reported under "sub_records".
One JSON line is printed by rank 0:

  value     whole-job samples/s with the dataset resident in HBM (60000 x 784 f32 = 188 MB > L2; every step gathers a fresh
            batch on the device), CUDA events on the launching stream, max over ranks
  e2e       the same metric through the reference-facing call a user makes, Trainer::train_epoch over a DataLoader
            (tp_trainer_train_epoch): per step the host-side batch gather (src/data/mnist.rs:276-309), the H2D copy of the
            batch and the D2H read of {loss, #correct} are all inside the timed region (wall clock, max over ranks)
  roofline  the dominant kernel of the step with its in-situ launch duration
  cpu_baseline  the CPU restatement of the reference (oracle/, NumPy + OpenBLAS) on this box's host cores
  dp_check  (N > 1) replicas bit-identical, N ranks x B match 1 rank x N*B and the oracle's first step to 1e-4

--impl reference times the oracle port instead (the reference is a Rust crate; no Rust toolchain exists in the image, see
DESIGN.md), same metric / config, all host threads.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F32 = np.float32

CONFIGS = {
    # name: (layer spec key, oracle builder, batch/GPU, optimizer, lr, weight decay, sample shape, BASELINE.json config text)
    "cfg2": ("MLP_784_128_10", ("mlp", [784, 128, 10]), 512, "adam", 1e-3, 0.0, (784,),
             "configs[1]: MLP 784-128-10, batch 512, Adam, fp32, tape fwd/bwd on CUDA"),
    "cfg1": ("MLP_784_128_10", ("mlp", [784, 128, 10]), 64, "sgd", 0.01, 0.0, (784,),
             "configs[0]: MLP 784-128-10, batch 64, SGD"),
    "example_mlp": ("MLP_EXAMPLE", ("mlp", [784, 128, 64, 10]), 256, "adam", 1e-3, 1e-4, (784,),
                    "examples/train_mnist.rs: MLP 784-128-64-10, batch 256, Adam(1e-3, wd 1e-4)"),
    "cfg4": ("MLP_784_1024_1024_10", ("mlp", [784, 1024, 1024, 10]), 1024, "adam", 1e-3, 0.0, (784,),
             "configs[3]: MLP 784-1024-1024-10, batch 1024/GPU, cross-entropy, Adam"),
    "cnn2": ("CNN2", ("cnn2", None), 256, "adam", 0.01, 1e-4, (1, 28, 28),
             "configs[2](i): Conv3x3-ReLU-MaxPool x2 + Linear, batch 256, Adam"),
    "cnn5": ("CNN5", ("cnn5", None), 256, "adam", 0.01, 1e-4, (1, 28, 28),
             "configs[2](ii): examples/train_mnist_cnn.rs 5-conv CNN, batch 256, Adam"),
    "cfg5": ("CNN5", ("cnn5", None), 1024, "adamw", 0.01, 1e-4, (1, 28, 28),
             "configs[4]: 5-conv CNN, batch 1024/GPU, AdamW + StepLR(5 epochs, 0.8) stepped every 59-step epoch, synthetic 28x28x1"),
}
PRIMARY = "cfg4"                    # BASELINE.json: the metric is quoted "at 1/2/4/8 B200" on configs[3]
SUBS = ["cfg5", "cfg2"]             # configs[4] (the >= 6x scaling target) and configs[1] (round 1's headline)
EPOCH_STEPS = 59                    # 60000 / 1024: cfg5 steps its LR scheduler (src/optim.rs:190-219) once per epoch-equivalent
DATASET_N = 60000                   # MNIST-sized


def workload_config(name, world):
    """The `config` object: identical in both arms (ours / reference) for the same workload."""
    spec_key, (kind, arg), batch, opt_kind, lr, wd, sample_shape, workload = CONFIGS[name]
    return {"workload": workload, "name": name, "batch_per_gpu": batch, "global_batch": batch * world, "optimizer": opt_kind,
            "lr": lr, "weight_decay": wd, "parallelism": f"dp{world}",
            "dataset": f"synthetic MNIST-shaped, {DATASET_N} x {int(np.prod(sample_shape))} u8 pixels (f32 = u8 / 255, src/data/mnist.rs:225) + "
                       "labels, per rank",
            "l2_policy": f"inputs larger than L2: every step takes a fresh batch out of the {DATASET_N}-sample dataset (188 MB as f32); "
                         "parameters and optimizer state are the step's own working set"}


def synthetic_u8(n, sample_shape, seed):
    rng = np.random.default_rng(seed)
    xu = rng.integers(0, 256, (n,) + tuple(sample_shape), dtype=np.uint8)      # MNIST pixels as on disk
    y = rng.integers(0, 10, n).astype(F32)                                      # labels stored as f32 (src/data/mnist.rs:268)
    return xu, y


def to_f32(xu):
    return (xu.astype(F32) / F32(255.0)).astype(F32)                            # src/data/mnist.rs:225


def mlp_flops(sizes, batch):
    """Algorithmic GEMM flops of one step: fwd + dW for every layer, dX for all but the first."""
    f = 0
    for i in range(len(sizes) - 1):
        g = 2 * batch * sizes[i] * sizes[i + 1]
        f += 2 * g + (g if i > 0 else 0)
    return f


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, windows):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(keep):
            sm, mx, reasons = [], [], set()
            for t, line in self.lines:
                if not keep(t):
                    continue
                f = [s.strip() for s in line.split(",")]
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except Exception:
                    continue
                for nme, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            return sm, mx, reasons

        sm, mx, reasons = collect(lambda t: any(a <= t <= b for a, b in windows))
        note = None
        if len(sm) < 3 and windows:
            # timed regions shorter than the 50 ms sampling period: use the samples within 0.5 s of them instead
            lo, hi = min(a for a, _ in windows) - 0.5, max(b for _, b in windows) + 0.5
            sm, mx, reasons = collect(lambda t: lo <= t <= hi)
            note = "timed regions shorter than the sampling period: samples within 0.5 s of them"
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


# ---- the reference's CPU implementation of the path: the oracle port --------------------------------------------------------
def build_oracle(kind, arg, seed):
    from oracle import taper_ref as R
    rng = np.random.default_rng(seed)
    if kind == "mlp":
        return R.build_mlp(arg, rng)
    return R.build_cnn2(rng) if kind == "cnn2" else R.build_cnn5(rng)


def blas_threads(n):
    """Pins the BLAS pools behind NumPy to n threads (torchrun exports OMP_NUM_THREADS=1: undo that explicitly)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def time_oracle(name, steps, warmup, threads, budget_s=None):
    """Times oracle train steps (the CPU restatement of the reference) at the config's own batch size with `threads` BLAS
    threads; returns (samples/s, steps run, seconds)."""
    from oracle import taper_ref as R
    spec_key, (kind, arg), batch, opt_kind, lr, wd, sample_shape, _ = CONFIGS[name]
    model = build_oracle(kind, arg, 0)
    params = model.parameters()
    opt = {"sgd": lambda: R.SGD(params, lr), "adam": lambda: R.Adam(params, lr, None, None, wd),
           "adamw": lambda: R.AdamW(params, lr, None, None, wd)}[opt_kind]()
    nb = 4
    xu, y = synthetic_u8(batch * nb, sample_shape, 1)
    x = to_f32(xu)

    def one(i):
        s = (i % nb) * batch
        R.train_step(model, opt, R.Tensor.new(x[s:s + batch], (batch,) + tuple(sample_shape)), R.Tensor.new(y[s:s + batch], (batch,)))

    with blas_threads(threads):
        t_w = time.perf_counter()
        for i in range(warmup):
            one(i)
            if budget_s is not None and time.perf_counter() - t_w > budget_s / 3:
                break
        t0 = time.perf_counter()
        done = 0
        for i in range(steps):
            one(i)
            done += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s and done >= 2:
                break
        dt = time.perf_counter() - t0
    return done * batch / dt, done, dt


def run_reference(args, name):
    """--impl reference: the reference's CPU implementation of the path = the oracle port, all host threads, the config's own
    batch size; a bounded sample (the whole run ends within a few minutes)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    budget = 90.0
    value, done, dt = time_oracle(name, args.steps, args.warmup, cores, budget_s=budget)
    sample = f"{done} oracle train steps at the config's batch size in {dt:.1f} s, {cores} BLAS threads"
    if done < args.steps:
        sample += f" (stopped at the {budget:.0f} s cap; {args.steps} requested)"
    v1, d1, t1 = time_oracle(name, 3, 1, 1, budget_s=15.0)
    out = {
        "impl": "reference", "metric": "MNIST samples/sec (fwd+bwd+step)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": dt / done * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, max(world, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": sample + "; NumPy/OpenBLAS restatement of the reference tape (oracle/taper_ref.py); "
                                            "the Rust reference cannot be built here (no cargo/rustc)",
                         "single_thread": {"value": v1, "unit": "samples/s", "sample": f"{d1} steps in {t1:.1f} s, 1 BLAS thread "
                                           "(the reference's default matrixmultiply build is single-threaded in GEMM)"}},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ---- our arm -----------------------------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed is plumbing only: rendezvous, the NCCL id / IPC handle exchange, max-over-ranks of the timings."""

    def __init__(self, world, local):
        self.world, self.pg = world, None
        if world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            self.pg, self.torch = dist, torch

    def barrier(self):
        from taper_b200 import host
        host.sync()
        if self.pg:
            self.pg.barrier()

    def max(self, v):
        if not self.pg:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.pg.all_reduce(t, op=self.pg.ReduceOp.MAX)
        return float(t.item())

    def mean(self, v):
        if not self.pg:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.pg.all_reduce(t)
        return float(t.item()) / self.world

    def all_ok(self, ok):
        if not self.pg:
            return ok
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device="cuda")
        self.pg.all_reduce(t, op=self.pg.ReduceOp.MIN)
        return int(t.item()) == 1

    def same_as_rank0(self, arr):
        if not self.pg:
            return True
        t = self.torch.from_numpy(np.ascontiguousarray(arr)).cuda()
        r0 = t.clone()
        self.pg.broadcast(r0, 0)
        return bool(self.torch.equal(t, r0))

    def nccl_id(self, rank):
        from taper_b200 import host
        uid = self.torch.zeros(128, dtype=self.torch.uint8, device="cuda")
        if rank == 0:
            uid = self.torch.frombuffer(bytearray(host.nccl_unique_id()), dtype=self.torch.uint8).cuda()
        self.pg.broadcast(uid, 0)
        return bytes(uid.cpu().numpy().tobytes())

    def close(self):
        if self.pg:
            self.pg.barrier()
            self.pg.destroy_process_group()


def make_trainer(name, args, dist, rank, world, eps=1e-8, nccl_only=False, seed=0):
    from taper_b200 import host
    spec_key, (kind, arg), batch, opt_kind, lr, wd, sample_shape, _ = CONFIGS[name]
    model = host.Model(getattr(host, spec_key), seed=seed)
    tr = host.Trainer(model, opt_kind, lr=lr, weight_decay=wd, eps=eps)
    if args.no_fused:
        tr.set_use_fused(False)
    if world > 1:
        tr.comm_init(rank, world, dist.nccl_id(rank))
        tr.broadcast_params(0)
        if not (args.nccl_only or nccl_only or args.no_fused):
            # small models: gradient exchange inside the persistent step kernel over NVLink peer memory.  Every rank must take
            # the same path, so a rank that cannot map its peers sends everybody back to the NCCL-in-graph path.  (The wide
            # plan of configs[3] sums its gradient arena with NCCL between its fold and optimizer kernels.)
            ok = True
            try:
                tr.peer_exchange_init(dist.pg)
            except Exception as e:
                ok = False
                print(f"rank {rank}: peer exchange unavailable ({e}); falling back to the NCCL allreduce", file=sys.stderr)
            if not dist.all_ok(ok):
                tr.set_use_fused(False)
    return model, tr


def dp_check(name, args, dist, rank, world, with_oracle):
    """Driver-visible data-parallel correctness: k steps with W ranks x B rows against (i) each other — replicas
    must stay bit-identical, (ii) one rank x W*B rows on this GPU, (iii) the oracle's first step on the global batch.
    Adam's eps is 0.1 here so that the update stays linear in the gradient and the plain 1e-4 bound applies (default-eps
    trajectories amplify summation-order noise; tests/test_step_gpu.py::close_after_adam)."""
    from taper_b200 import host
    spec_key, (kind, arg), batch, opt_kind, lr, wd, sample_shape, _ = CONFIGS[name]
    k = 2
    G = batch * world
    xu, y = synthetic_u8(G * k, sample_shape, 4242)                    # the same global batches on every rank
    x = to_f32(xu)
    # (ii) single replica, global batch, no communicator yet on this context
    ref_model = host.Model(getattr(host, spec_key), seed=0)
    ref_tr = host.Trainer(ref_model, opt_kind, lr=lr, weight_decay=wd, eps=0.1)
    if args.no_fused:
        ref_tr.set_use_fused(False)
    ref_losses = [ref_tr.step(x[s * G:(s + 1) * G], y[s * G:(s + 1) * G])[0] for s in range(k)]
    ref_params = [ref_model.get_param(i) for i in range(ref_model.num_params())]
    del ref_tr
    model, tr = make_trainer(name, args, dist, rank, world, eps=0.1)
    out = {"world": world, "steps": k, "ok": True, "adam_eps": 0.1}
    losses = []
    p_after1 = None
    for s in range(k):
        lo = s * G + rank * batch
        losses.append(tr.step(x[lo:lo + batch], y[lo:lo + batch])[0])
        if s == 0:
            p_after1 = [model.get_param(i) for i in range(model.num_params())]
    params = [model.get_param(i) for i in range(model.num_params())]
    flat = np.concatenate([p.reshape(-1) for p in params])
    identical = dist.all_ok(dist.same_as_rank0(flat))
    out["replicas_bit_identical"] = identical
    def param_err(got, ref):
        """max over tensors of max|got - ref| / scale; a bias is judged on its layer's weight scale (they add into the same
        pre-activation; a freshly initialised bias is ~0, so its own norm is one update — and one ReLU unit within summation
        noise of 0 legitimately moves a bias gradient by a sample's share, tests/test_step_gpu.py::relu_tie_slack)."""
        worst, per = 0.0, []
        for i, (p, r) in enumerate(zip(got, ref)):
            scale = max(float(np.max(np.abs(r))), 1e-6)
            if r.ndim == 1 and i > 0:
                scale = max(scale, float(np.max(np.abs(ref[i - 1]))))
            e = float(np.max(np.abs(np.asarray(p).reshape(-1) - np.asarray(r).reshape(-1)))) / scale
            per.append(e)
            worst = max(worst, e)
        return worst, per
    worst_loss = 0.0
    for s in range(k):
        worst_loss = max(worst_loss, abs(dist.mean(losses[s]) - ref_losses[s]) / abs(ref_losses[s]))
    worst_param, per = param_err(params, ref_params)
    out["vs_one_rank_global_batch"] = {"loss_rel": worst_loss, "param_rel_inf": worst_param, "per_tensor": per}
    ok = identical and worst_loss <= 1e-4 and worst_param <= 1e-4
    if with_oracle:
        from oracle import taper_ref as R
        o = build_oracle(kind, arg, 0)
        init = host.Model(getattr(host, spec_key), seed=0)         # the CUDA model's own seeded init, injected into the oracle
        for i, p in enumerate(o.parameters()):
            p._data[:] = init.get_param(i).reshape(-1)
        prm = o.parameters()
        oopt = {"sgd": lambda: R.SGD(prm, lr), "adam": lambda: R.Adam(prm, lr, None, 0.1, wd),
                "adamw": lambda: R.AdamW(prm, lr, None, 0.1, wd)}[opt_kind]()
        with blas_threads(os.cpu_count() or 1):
            l_or, _ = R.train_step(o, oopt, R.Tensor.new(x[:G], (G,) + tuple(sample_shape)), R.Tensor.new(y[:G], (G,)))
        lo_rel = abs(dist.mean(losses[0]) - l_or) / abs(l_or)
        po, per_o = param_err(p_after1, [np.asarray(b.data()).reshape(a.shape) for a, b in zip(p_after1, prm)])
        out["vs_oracle_step1"] = {"loss_rel": lo_rel, "param_rel_inf": po, "per_tensor": per_o}
        ok = ok and lo_rel <= 1e-4 and po <= 1e-4
    peer = not (args.nccl_only or args.no_fused) and tr.fused_kind() in (1, 2)
    out["exchange"] = ("two-phase NVLink peer-memory exchange fused with the optimizer (tp_xchg_*)" if tr.fused_kind() == 2 else
                       "in-kernel NVLink peer-memory exchange (tp_xchg_*)") if peer else "NCCL allreduce of the gradient arena inside the step"
    out["ok"] = bool(dist.all_ok(ok))
    del tr
    return out


def measure(name, args, dist, rank, world, local, windows, primary):
    """value + e2e of one workload.  Returns the record (rank 0 uses it) — every rank must call it."""
    from taper_b200 import host
    spec_key, (kind, arg), batch, opt_kind, lr, wd, sample_shape, _ = CONFIGS[name]
    cols = int(np.prod(sample_shape))
    model, tr = make_trainer(name, args, dist, rank, world)
    # each rank owns a shard: its own dataset (weak scaling: batch per GPU fixed)
    Xu, Y = synthetic_u8(DATASET_N, sample_shape, 1 + rank)
    X = to_f32(Xu)
    perm = np.random.default_rng(100 + rank).permutation(DATASET_N).astype(np.uint32)
    tr.load_dataset(X, Y, perm)                                        # f32 resident: 188 MB > L2
    W = max(args.warmup, 3)
    last = (0.0, 0.0)
    for _ in range(W):
        tr.step_resident(batch)
        last = tr.fetch()
    dist.barrier()
    e0, e1 = host.Event(), host.Event()
    l0 = host.launches()
    w0 = time.perf_counter()
    e0.record()
    sched_epoch = 0
    for i in range(args.steps):
        if tr.pending() >= 6:
            last = tr.fetch()
        tr.step_resident(batch)
        if name == "cfg5" and i % EPOCH_STEPS == EPOCH_STEPS - 1:     # StepLR(step 5, gamma 0.8) + optimizer.set_lr (src/train.rs:212-216)
            sched_epoch += 1
            tr.set_lr(lr * 0.8 ** (sched_epoch // 5))
    e1.record()
    while tr.pending():
        last = tr.fetch()
    host.sync()
    w1 = time.perf_counter()
    gpu_launches = host.launches() - l0
    ms = dist.max(e0.elapsed_ms(e1))
    dist.barrier()
    windows.append((w0, w1))
    value = world * batch * args.steps / (ms * 1e-3)
    fused_kind = tr.fused_kind()

    # ---- e2e: Trainer::train_epoch over a DataLoader (host gather into pinned memory, H2D, step, D2H of the results) ----------
    wide = fused_kind == 2
    ds = host.Dataset(Xu.reshape(DATASET_N, cols) if wide else X.reshape(DATASET_N, cols), Y)
    loader = host.Loader(ds, batch, shuffle=False, sample_shape=sample_shape)      # shuffle off: K steps are a fraction of an epoch
    tr.train_epoch(loader, max_batches=W)
    per_epoch = max(1, DATASET_N // batch)              # K steps may span several passes over the 60000-sample dataset (train_epoch
                                                        # resets the loader on entry; only full batches are timed)
    dist.barrier()
    t0 = time.perf_counter()
    done = 0
    while done < args.steps:
        k = min(args.steps - done, per_epoch)
        e2e_loss, e2e_acc = tr.train_epoch(loader, max_batches=k)
        done += k
    host.sync()
    t1 = time.perf_counter()
    e2e_s = dist.max(t1 - t0)
    dist.barrier()
    windows.append((t0, t1))
    rec = {
        "config": workload_config(name, world),
        "value": value, "unit": "samples/s", "ms_per_step": ms / args.steps, "steps": args.steps, "warmup": W,
        "e2e": {"value": world * batch * args.steps / e2e_s, "unit": "samples/s",
                "h2d_bytes_per_step": int(batch * (cols * (1 if wide else 4) + 4)), "d2h_bytes_per_step": 16,
                "ms_per_step": e2e_s / args.steps * 1e3, "input_dtype": "u8 pixels (divided by 255 on the device)" if wide else "f32",
                "call": "tp_trainer_train_epoch(trainer, loader, max_batches = steps): loader worker threads gather each batch into pinned "
                        "memory, H2D on the copy stream, results read back per step",
                "timing": "host wall clock around the call, device synchronised on both sides, max over ranks",
                "mean_loss": e2e_loss, "accuracy": e2e_acc},
        "gpu_launches": int(gpu_launches), "launches_per_step": gpu_launches / args.steps,
        "step_path": {2: "wide device tape: a plan of tcgen05 kernels chained with programmatic dependent launch (bf16x3 GEMMs on pre-split "
                         "operands, fused head / epilogues / optimizer)" + ("; NCCL allreduce of the gradient arena" if world > 1 else ""),
                      1: "device tape: one persistent kernel per step (grid barrier between phases; exact fp32 FFMA)"
                         + ("; gradient exchange in-kernel over NVLink peer memory" if world > 1 else ""),
                      0: "tape + CUDA graph, one kernel per op (3xTF32 tcgen05 GEMMs / implicit-GEMM convolutions)"
                         + ("; NCCL allreduce in the graph" if world > 1 else "")}[fused_kind],
        "gemm_math": {2: "bf16x3 on tcgen05 (kind::f16, 3 MMAs per product, ~1e-5 of |C|inf)", 1: "exact fp32 FFMA on the CUDA cores",
                      0: {0: "fp32 FFMA", 1: "3xTF32 tcgen05 (fp32-accurate)", 2: "1xTF32 tcgen05", 3: "bf16x3 tcgen05"}[args.gemm_mode]}[fused_kind],
        "conv_adjoint": ("full" if args.full_adjoint else "strict_reference (SURVEY A1)") if kind != "mlp" else None,
        "last_step": {"loss": last[0], "correct": last[1]},
    }
    del tr
    return rec


def plan_profile(name, peaks, tf32_peak):
    """In-situ duration of every kernel of the wide plan (one %globaltimer stamp per kernel, taken when its programmatic
    dependency wait is over, i.e. launch gaps are charged to the kernel before) on the config's shapes, with the roofline
    each kernel is bound by.  Drives the C ABI directly on a plan with random parameters."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import step_profile as SP
    import taper_b200
    from taper_b200 import capi
    lib = capi.lib
    spec_key, (kind, dims), batch, opt_kind, lr, wd, sample_shape, _ = CONFIGS[name]
    ctx = taper_b200.Ctx(0)
    d, step, keep = SP.build(ctx, dims, batch, opt_kind)
    if lib.tp_step_is_wide(step) != 1:
        capi.check(lib.tp_step_destroy(step))
        return None, []
    rng = np.random.default_rng(1)
    n = DATASET_N
    x = ctx.upload(rng.random((n, dims[0]), dtype=np.float32))
    y = ctx.upload(rng.integers(0, dims[-1], n).astype(np.float32))
    perm = ctx.upload(rng.permutation(n).astype(np.int32))
    cursor = ctx.upload(np.zeros(1, np.int32))
    run = lambda: capi.check(lib.tp_step_run(ctx.h, step, x.h, y.h, perm.h, cursor.h, n, -1, 0.01, 1.0, None, 0))
    for _ in range(10):
        run()
    capi.check(lib.tp_step_set_profile(step, 1))
    L = len(dims) - 1
    names = ["input"] + [f"fwd{l}" for l in range(L - 1)] + ["head"] + [f"dX{l}" for l in range(L - 2, 0, -1)] + ["dW_all", "fold", "optimizer"]
    nph = C.c_int()
    capi.check(lib.tp_step_info(step, C.byref(nph), None, None))
    folded = nph.value == len(names) - 1                   # the fold rode along on extra CTAs of the grouped dW launch
    if folded:
        names.remove("fold")
    acc = np.zeros(len(names))
    reps = 20
    for _ in range(reps):
        for _ in range(3):
            run()
        buf = np.zeros(32, np.int64)
        slots = C.c_int()
        capi.check(lib.tp_step_read_profile(step, buf.ctypes.data_as(C.POINTER(C.c_int64)), buf.size, C.byref(slots)))
        prev, lastv = buf[:16].astype(np.float64), buf[16:].astype(np.float64)
        tl = list(prev[:len(names)]) + [lastv[0]]
        acc += np.diff(tl) / 1e3
    us = acc / reps
    capi.check(lib.tp_step_destroy(step))
    hbm = peaks.get("hbm_gbs", 6650.0)
    bf16 = peaks.get("bf16_tflops", 1590.0)
    B = batch
    n_param = sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(L))
    out = []

    def tensor_entry(label, flops, t_us, note):
        ach = flops / (t_us * 1e-6) / 1e12
        e = {"kernel": label, "bound": "tensor", "achieved": ach, "peak": tf32_peak or bf16 / 2, "unit": "TFLOP/s",
             "frac": ach / (tf32_peak or bf16 / 2), "traffic": None, "launch_us": t_us,
             "peak_source": "measured cuBLAS TF32 8192^3 (this run): the fp32 tensor-core peak north_star names" if tf32_peak
             else "MEASURED_PEAKS.json bf16_tflops / 2",
             "tensor_pipe_frac": 3 * ach / bf16,
             "note": note + "; algorithmic flops 2*M*N*K; the kernel issues 3 bf16 MMAs per product, so its share of the measured bf16 "
                            f"peak ({bf16:.0f} TFLOP/s, MEASURED_PEAKS.json) is tensor_pipe_frac; launch_us is the in-situ slot "
                            "(launch gap to the next kernel included)"}
        out.append(e)
        return e

    def hbm_entry(label, nbytes, t_us, note):
        ach = nbytes / (t_us * 1e-6) / 1e9
        out.append({"kernel": label, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
                    "launch_us": t_us, "algorithmic_bytes": nbytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                    "note": note})

    i = 0
    hbm_entry("wide_input_kernel: gather + bf16 hi/lo planes", B * dims[0] * 8 + B * 8, us[i], "latency-bound at this size (3.2 MB)"); i += 1
    gemms = []
    for l in range(L - 1):
        gemms.append(tensor_entry(f"gemm_bx3_kernel fwd{l} {B}x{dims[l + 1]}x{dims[l]} (bias+ReLU+planes epilogue)", 2.0 * B * dims[l] * dims[l + 1],
                                  us[i], "N,T"))
        i += 1
    hbm_entry("wide_head_kernel: classifier + softmax-CE + dlogits + dZ planes", B * dims[L - 1] * 8 + B * 64, us[i], "latency-bound"); i += 1
    for l in range(L - 2, 0, -1):
        gemms.append(tensor_entry(f"gemm_bx3_kernel dX{l} {B}x{dims[l]}x{dims[l + 1]} (ReLU-mask + bias-gradient + planes epilogue)",
                                  2.0 * B * dims[l] * dims[l + 1], us[i], "N,N"))
        i += 1
    fl = sum(2.0 * B * dims[l] * dims[l + 1] for l in range(L))
    gemms.append(tensor_entry(f"gemm_bx3_kernel grouped dW (all {L} weight gradients in one launch" + (", bias-gradient fold on spare CTAs)" if folded else ")"),
                              fl, us[i], "T,N")); i += 1
    if not folded:
        hbm_entry("wide_fold_kernel: bias-gradient partials, results, Adam counters", 4 * 160 * sum(dims[1:]), us[i], "latency-bound"); i += 1
    hbm_entry("adam_dev_split_kernel: fused optimizer + parameter planes", (28 + 4) * n_param, us[i],
              "L2-resident state (52 MB < 126 MB L2): reported against HBM peak, so frac may exceed what DRAM alone allows"); i += 1
    roof = max(gemms, key=lambda e: e["launch_us"])
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        t = tj.get(name, {})
        if t.get("kernel") and t["kernel"] in roof["kernel"]:
            roof = dict(roof, traffic=t.get("dram_bytes_per_launch"), traffic_source=t.get("source"))
    except Exception:
        pass
    return roof, out


def measure_tf32_peak():
    """cuBLAS TF32 8192^3 through torch (library GEMM, used only as the roofline denominator)."""
    try:
        import torch
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(8192, 8192, device="cuda")
        b = torch.randn(8192, 8192, device="cuda")
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(8):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); a @ b; e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        del a, b
        torch.cuda.empty_cache()
        return 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def probe_kernels(peaks, tf32_peak):
    """Large-size evidence runs of single kernels through the C ABI (BASELINE metric: GEMM % TC-peak, elementwise % HBM)."""
    from taper_b200 import capi, host
    h = host.host_ctx()
    lib = capi.lib

    class HB:
        def __init__(self, n):
            self.h = C.c_void_p(); self.n = n
            capi.check(lib.tp_buf_alloc(h, n, C.byref(self.h)))
            capi.check(lib.tp_buf_fill(h, self.h, 0.5, n))

        def __del__(self):
            lib.tp_buf_release(self.h)

    def timeit(fn, reps):
        for _ in range(3):
            fn(0)
        e0, e1 = host.Event(), host.Event()
        host.sync()
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record(); e1.sync()
        return e0.elapsed_ms(e1) / reps * 1e-3

    out = []
    hbm = peaks.get("hbm_gbs", 6650.0)
    tc_peak = tf32_peak or peaks.get("bf16_tflops", 1590.0) / 2
    tc_note = "measured cuBLAS TF32 8192^3 (this run)" if tf32_peak else "bf16 measured / 2"
    try:
        nn_ = 8192
        ga, gb, gc = HB(nn_ * nn_), HB(nn_ * nn_), HB(nn_ * nn_)
        mode0 = C.c_int()
        capi.check(lib.tp_get_gemm_mode(h, C.byref(mode0)))
        for mode, label in ((2, "1xTF32, 128x256 tiles"), (1, "3xTF32 = fp32-accurate, algorithmic flops"),
                            (3, "bf16x3 incl. the two operand-split launches, algorithmic flops")):
            capi.check(lib.tp_set_gemm_mode(h, mode))
            t = timeit(lambda i: capi.check(lib.tp_sgemm_rowmajor(h, 0, 1, nn_, nn_, nn_, 1.0, ga.h, gb.h, 0.0, gc.h)), 5)
            fl = 2.0 * nn_ ** 3
            out.append({"kernel": f"tp_sgemm_rowmajor 8192^3 N,T tcgen05 ({label})", "bound": "tensor", "achieved": fl / t / 1e12,
                        "peak": tc_peak, "unit": "TFLOP/s", "frac": fl / t / 1e12 / tc_peak, "traffic": None, "launch_us": t * 1e6,
                        "peak_source": tc_note, "probe": True})
        capi.check(lib.tp_set_gemm_mode(h, mode0.value))
        del ga, gb, gc
    except Exception as e:                       # never let an evidence probe take the bench line down
        out.append({"kernel": "tp_sgemm_rowmajor 8192^3", "error": str(e)})
    n = 48 * 1024 * 1024
    p, g, m, v, hy = HB(n), HB(n), HB(n), HB(n), HB(8)
    capi.check(lib.tp_adam_hyper_init(h, hy.h, 1e-3, 0.9, 0.999, 1e-8, 0.0))
    capi.check(lib.tp_adam_advance(h, hy.h))
    t = timeit(lambda i: capi.check(lib.tp_adam_step_dev(h, p.h, g.h, m.h, v.h, hy.h, 1.0, 0, n)), 20)
    out.append({"kernel": f"adam_step {n} params (fused, flat arena)", "bound": "hbm", "achieved": 28 * n / t / 1e9, "peak": hbm,
                "unit": "GB/s", "frac": 28 * n / t / 1e9 / hbm, "traffic": None, "launch_us": t * 1e6, "probe": True,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650"})
    t = timeit(lambda i: capi.check(lib.tp_relu_bwd(h, p.h, g.h, m.h, n, 0)), 20)
    out.append({"kernel": f"relu_bwd {n} elements", "bound": "hbm", "achieved": 12 * n / t / 1e9, "peak": hbm, "unit": "GB/s",
                "frac": 12 * n / t / 1e9 / hbm, "traffic": None, "launch_us": t * 1e6, "probe": True,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650"})
    return out


def run_ours(args, name):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    import taper_b200                      # raises if libtaper_b200.so is missing: there is no fallback
    from taper_b200 import host
    host.set_device(local)
    host.config(conv_full_adjoint=1 if args.full_adjoint else 0, gemm_mode=args.gemm_mode)
    dist = Dist(world, local)
    sampler = ClockSampler(local) if rank == 0 else None
    windows = []

    check = None
    if world > 1 and not args.no_dp_check:
        check = dp_check(name, args, dist, rank, world, with_oracle=True)
    rec = measure(name, args, dist, rank, world, local, windows, True)
    subs = []
    if not args.no_subs and name == PRIMARY:
        for sub in SUBS:
            sw = []
            r = measure(sub, args, dist, rank, world, local, sw, False)
            if world > 1 and not args.no_dp_check:
                r["dp_check"] = dp_check(sub, args, dist, rank, world, with_oracle=False)
            subs.append(r)
    clocks = sampler.stop(windows) if sampler else None
    if rank != 0:
        dist.close()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    kind, arg = CONFIGS[name][1]
    roofline, kernels = None, []
    tf32_peak = None
    if world == 1:
        tf32_peak = measure_tf32_peak()
        if kind == "mlp" and not args.no_fused:
            try:
                roofline, kernels = plan_profile(name, peaks, tf32_peak)
            except Exception as e:
                kernels = [{"kernel": "plan_profile", "error": str(e)}]
        if not args.no_probes:
            kernels = kernels + probe_kernels(peaks, tf32_peak)
        if roofline is None:
            cands = [k for k in kernels if "error" not in k]
            roofline = max(cands, key=lambda k: k["launch_us"]) if cands else None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, done, dt = time_oracle(name, 10 ** 9, 1, cores, budget_s=12.0)
        v1, d1, t1 = time_oracle(name, 3, 1, 1, budget_s=8.0)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{done} oracle train steps at the config's batch size in {dt:.1f} s, {cores} BLAS threads "
                         "(NumPy/OpenBLAS restatement of the reference tape, oracle/taper_ref.py)",
               "single_thread": {"value": v1, "unit": "samples/s", "sample": f"{d1} steps in {t1:.1f} s, 1 BLAS thread"}}
    flops = mlp_flops(arg, CONFIGS[name][2]) if kind == "mlp" else None
    out = {
        "metric": "MNIST samples/sec (fwd+bwd+step)", "value": rec["value"], "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": rec["warmup"], "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": rec["config"],
        "e2e": rec["e2e"],
        "gpu_launches": rec["gpu_launches"], "launches_per_step": rec["launches_per_step"],
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "impl_details": {k: rec[k] for k in ("step_path", "gemm_math", "conv_adjoint", "last_step")},
        "dp_check": check,
        "sub_records": [{k: r[k] for k in ("config", "value", "unit", "ms_per_step", "e2e", "gpu_launches", "launches_per_step", "step_path",
                                            "gemm_math", "conv_adjoint", "last_step") if k in r} | ({"dp_check": r["dp_check"]} if "dp_check" in r else {})
                        for r in subs],
        "kernels": kernels,
        "step_gemm_flops": flops,
        "step_tensor_frac": (flops / (rec["ms_per_step"] * 1e-3) / 1e12 / tf32_peak) if (flops and tf32_peak) else None,
    }
    print(json.dumps(out), flush=True)
    dist.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=PRIMARY, choices=sorted(CONFIGS))
    ap.add_argument("--gemm-mode", type=int, default=1, choices=[0, 1, 2, 3],
                    help="GEMM math of the tape + graph path (the wide plan always runs bf16x3, the persistent kernel exact fp32)")
    ap.add_argument("--full-adjoint", action="store_true", help="CNN: compute conv dW/dX (the reference does not, SURVEY A1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-subs", action="store_true", help="skip the sub-records (configs[4], configs[1])")
    ap.add_argument("--no-probes", action="store_true", help="skip the large single-kernel evidence probes")
    ap.add_argument("--no-dp-check", action="store_true")
    ap.add_argument("--nccl-only", action="store_true", help="N>1, small models: keep the NCCL allreduce (tape + CUDA-graph path) instead "
                                                             "of the in-kernel NVLink peer-memory exchange of the persistent step")
    ap.add_argument("--no-fused", action="store_true", help="run the tape + CUDA-graph path even where a fused device step qualifies")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 300 if args.impl == "ours" else 20
    if args.impl == "reference":
        run_reference(args, args.config)
    else:
        run_ours(args, args.config)


if __name__ == "__main__":
    main()
    # The JSON line is out and the process group is closed.  Leave without interpreter finalisation: at N > 1 the process holds
    # two NCCL instances (torch's and the one libtaper_b200 binds with dlopen) plus CUDA IPC mappings of the peers' exchange
    # windows, and their static destructors racing Python's module teardown has been seen to end a rank with SIGSEGV after all
    # the work was done (1 run in ~8 of scripts/dp_check.py at 2 GPUs).
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)
