"""Where does the fused device step (tp_step_*, taper_b200/csrc/tape_step.cu) spend its time?

Drives the C ABI directly (no trainer): builds an MLP step, runs it with per-CTA SM-clock stamps on and prints, per phase,
the work time (max / mean over CTAs) and the barrier wait.  Also times the step with CUDA events (stamps off).

    python scripts/step_profile.py --dims 784,128,10 --batch 512 --opt adam [--resident]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taper_b200 import capi                                       # noqa: E402
from taper_b200.capi import check, lib                            # noqa: E402


def build(ctx, dims, batch, opt, seed=0, xchg_factory=None):
    rng = np.random.default_rng(seed)
    d = capi.StepDesc()
    L = len(dims) - 1
    d.n_layers = L
    off = 0
    chunks = []
    for l in range(L):
        d.dims[l] = dims[l]
        d.relu[l] = 1 if l < L - 1 else 0
        w = (rng.random((dims[l + 1], dims[l]), dtype=np.float32) * 2 - 1) * np.float32(np.sqrt(2.0 / dims[l]))
        d.w_off[l] = off
        chunks.append(w.reshape(-1)); off += w.size
        pad = (-off) % 4
        chunks.append(np.zeros(pad, np.float32)); off += pad
        b = (rng.standard_normal(dims[l + 1]) * 0.05).astype(np.float32)
        d.b_off[l] = off
        chunks.append(b); off += b.size
        pad = (-off) % 4
        chunks.append(np.zeros(pad, np.float32)); off += pad
    d.dims[L] = dims[L]
    d.batch = batch
    d.optimizer = {"sgd": 0, "adam": 1, "adamw": 2}[opt]
    d.arena_len = off
    xchg = xchg_factory(off) if xchg_factory else None
    d.materialize_grads = 0
    params = ctx.upload(np.concatenate(chunks))
    grads, m, v = ctx.zeros(off), ctx.zeros(off), ctx.zeros(off)
    hyper = ctx.alloc(8)
    check(lib.tp_adam_hyper_init(ctx.h, hyper.h, 1e-3, 0.9, 0.999, 1e-8, 0.0))
    result = ctx.zeros(2)
    step = C.c_void_p()
    assert lib.tp_step_supported(C.byref(d)) == 1, "step not supported"
    check(lib.tp_step_create(ctx.h, C.byref(d), params.h, grads.h, m.h, v.h, hyper.h, result.h, xchg, C.byref(step)))
    return d, step, (params, grads, m, v, hyper, result, xchg)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="784,128,10")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--opt", default="adam")
    ap.add_argument("--resident", action="store_true", help="gather rows from a 60000-row dataset (188 MB > L2)")
    ap.add_argument("--iters", type=int, default=2000)
    a = ap.parse_args()
    dims = [int(x) for x in a.dims.split(",")]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    xf = None
    if world > 1:                                   # under torchrun: in-kernel NVLink peer-memory gradient exchange
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = capi.Ctx(local)
    if world > 1:
        def xf(arena_len):
            x = C.c_void_p()
            check(lib.tp_xchg_create(ctx.h, arena_len, rank, world, C.byref(x)))
            buf = C.create_string_buffer(64)
            check(lib.tp_xchg_handle(x, buf))
            handles = [None] * world
            dist.all_gather_object(handles, buf.raw)
            check(lib.tp_xchg_connect(x, b"".join(handles), world))
            return x
    d, step, keep = build(ctx, dims, a.batch, a.opt, xchg_factory=xf)
    rng = np.random.default_rng(1)
    n = 60000 if a.resident else a.batch
    x = ctx.upload(rng.random((n, dims[0]), dtype=np.float32))
    y = ctx.upload(rng.integers(0, dims[-1], n).astype(np.float32))
    perm = ctx.upload(rng.permutation(n).astype(np.int32)) if a.resident else None
    cursor = ctx.upload(np.zeros(1, np.int32)) if a.resident else None

    def run():
        check(lib.tp_step_run(ctx.h, step, x.h, y.h, perm.h if perm else None, cursor.h if cursor else None, n if a.resident else 0,
                              -1, 0.01, 1.0 / world, None, 0))

    nph, njobs, grid = C.c_int(), C.c_int(), C.c_int()
    check(lib.tp_step_info(step, C.byref(nph), C.byref(njobs), C.byref(grid)))
    for _ in range(20):
        run()
    ctx.sync()
    if dist is not None:
        dist.barrier()
    e0, e1 = C.c_void_p(), C.c_void_p()
    check(lib.tp_event_create(ctx.h, C.byref(e0))); check(lib.tp_event_create(ctx.h, C.byref(e1)))
    check(lib.tp_event_record(ctx.h, e0))
    for _ in range(a.iters):
        run()
    check(lib.tp_event_record(ctx.h, e1)); check(lib.tp_event_sync(e1))
    ms = C.c_float()
    check(lib.tp_event_elapsed_ms(e0, e1, C.byref(ms)))
    us = ms.value * 1e3 / a.iters
    if rank != 0:
        check(lib.tp_step_set_profile(step, 1))
        for _ in range(4 if lib.tp_step_is_wide(step) == 1 else 3):      # every rank must run the same number of steps
            run()
        ctx.sync()
        dist.barrier()
        dist.destroy_process_group()
        return
    print(f"world {world} dims {dims} batch {a.batch} {a.opt} resident={a.resident}: {nph.value} phases, {njobs.value} jobs, grid {grid.value}; "
          f"{us:.2f} us/step (CUDA events, {a.iters} back-to-back launches) = {a.batch / us:.2f} M samples/s")

    if lib.tp_step_is_wide(step) == 1:
        # the wide plan: one %globaltimer stamp per kernel (taken when its dependency wait is over) for the last two steps
        check(lib.tp_step_set_profile(step, 1))
        for _ in range(4):
            run()
        L = len(dims) - 1
        names = ["input"] + [f"fwd{l}" for l in range(L - 1)] + ["head"]
        names += [f"dX{l}" for l in range(L - 2, 0, -1)]
        names += ["dW (all)", "fold", "optimizer" if world == 1 else "exchange+optimizer"]
        check(lib.tp_step_info(step, C.byref(nph), C.byref(njobs), C.byref(grid)))
        if nph.value == len(names) - 1:               # the fold rode along on extra CTAs of the dW launch
            names.remove("fold")
            names[names.index("dW (all)")] = "dW (all) + fold"
        if world > 1 and os.environ.get("TAPER_WIDE_PUSH_IN_GEMM", "0") == "1" and L >= 3:
            # data parallel: weight gradients interleaved with the dX chain, their epilogues push over NVLink
            names = ["input"] + [f"fwd{l}" for l in range(L - 1)] + ["head", f"dW{L - 1}+dW{L - 2} (push)"]
            for l in range(L - 2, 0, -1):
                names += [f"dX{l}", f"dW{l - 1} (push)" + (" + fold" if l == 1 else "")]
            names += ["exchange+optimizer"]
        buf = np.zeros(64, np.int64)
        slots = C.c_int()
        check(lib.tp_step_read_profile(step, buf.ctypes.data_as(C.POINTER(C.c_int64)), buf.size, C.byref(slots)))
        prev, last, xprev = buf[:16].astype(np.float64), buf[16:32].astype(np.float64), buf[32:48].astype(np.float64)
        n = len(names)
        print(f"per-kernel time in situ (us, %globaltimer at the end of each kernel's dependency wait; step-to-step {(last[0] - prev[0]) / 1e3:.2f} us)")
        tl = list(prev[:n]) + [last[0]]
        for i, nm in enumerate(names):
            print(f"  {nm:10s} {(tl[i + 1] - tl[i]) / 1e3:7.2f}")
        if world > 1:
            x = np.concatenate([prev[n - 1:n], xprev[1:7]])
            lab = ["A: push my parts of the other slices", "A: release + raise flags", "A: wait for the peers' flags", "B: reduce my slice + push the result",
                   "B: release + raise + wait", "C: optimizer"]
            print("  exchange kernel, CTA 0 (us): " + " | ".join(f"{l} {(x[i + 1] - x[i]) / 1e3:.2f}" for i, l in enumerate(lab)))
            y = np.concatenate([prev[n - 1:n], xprev[1:10]])
            print(f"  exchange kernel, last CTA to pass (us after the kernel's start): peers' A flags seen {(y[7] - y[0]) / 1e3:.2f} | peers' B flags seen "
                  f"{(y[8] - y[0]) / 1e3:.2f} | optimizer done {(y[9] - y[0]) / 1e3:.2f}  (CTA 0: {(y[3] - y[0]) / 1e3:.2f} | {(y[5] - y[0]) / 1e3:.2f} | {(y[6] - y[0]) / 1e3:.2f})")
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        check(lib.tp_step_destroy(step))
        return
    check(lib.tp_step_set_profile(step, 1))
    for _ in range(3):
        run()
    slots = C.c_int()
    buf = np.zeros(grid.value * 64, np.int64)
    check(lib.tp_step_read_profile(step, buf.ctypes.data_as(C.POINTER(C.c_int64)), buf.size, C.byref(slots)))
    t = buf[: grid.value * slots.value].reshape(grid.value, slots.value).astype(np.float64)
    ghz = 1.965
    print(f"per-CTA SM-clock deltas in us at {ghz} GHz (max / mean over {grid.value} CTAs)")
    print(f"  setup           {np.max(t[:, 1] - t[:, 0]) / ghz / 1e3:6.2f} / {np.mean(t[:, 1] - t[:, 0]) / ghz / 1e3:6.2f}")
    prev = t[:, 1]
    for ph in range(nph.value):
        work = t[:, 2 + 2 * ph] - prev
        wait = t[:, 3 + 2 * ph] - t[:, 2 + 2 * ph]
        print(f"  phase {ph}: work {np.max(work) / ghz / 1e3:6.2f} / {np.mean(work) / ghz / 1e3:6.2f}   barrier wait {np.max(wait) / ghz / 1e3:6.2f} / "
              f"{np.mean(wait) / ghz / 1e3:6.2f} (min {np.min(wait) / ghz / 1e3:5.2f})")
        prev = t[:, 3 + 2 * ph]
    print(f"  total           {np.max(t[:, 1 + 2 * nph.value] - t[:, 0]) / ghz / 1e3:6.2f}")
    for name, base in (("fwd GEMM item (K-major A)", 24), ("dW GEMM item (MN-major A)", 32)):
        g = t[:, base:base + 5]
        ok = (g[:, 0] > 0) & (g[:, 4] > g[:, 0])
        if ok.any():
            d = np.diff(g[ok], axis=1) / ghz / 1e3
            print(f"  {name}: issue loads {d[:, 0].mean():.2f} | first stage lands + stash + sync {d[:, 1].mean():.2f} | multiply loop "
                  f"{d[:, 2].mean():.2f} | store partials {d[:, 3].mean():.2f}  (mean us over {int(ok.sum())} CTAs)")
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    check(lib.tp_step_destroy(step))


if __name__ == "__main__":
    main()
