"""tp_conv_stack_fwd on the GPU against the oracle, per shift mode, with timings.  Prints instead of asserting, so that one
GPU call tells everything (which descriptor variant is right, where errors sit, how long each stack takes).

usage: python scripts/conv_stack_probe.py [--big] [--time]
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import taper_ref as R  # noqa: E402

F32 = np.float32


def oracle_stack(x, ws, bs, pools, relus):
    t = R.Tensor.new(x, x.shape)
    for w, b, pool, relu in zip(ws, bs, pools, relus):
        wt = R.Tensor.new(w, w.shape)
        bt = R.Tensor.new(b, b.shape) if b is not None else None
        t = t.conv2d_relu(wt, bt, (1, 1), (1, 1), (1, 1)) if relu else t.conv2d(wt, bt, (1, 1), (1, 1), (1, 1))
        if pool:
            t = t.max_pool2d((2, 2), (2, 2))
    return t.data().reshape(t.shape)


def run_stack(ctx, x, ws, bs, pools, relus, reps=0):
    from taper_b200 import capi
    lib = capi.lib
    n, c, h, w = x.shape
    L = len(ws)
    xb = ctx.upload(x)
    wb = [ctx.upload(v) for v in ws]
    bb = [ctx.upload(v) if v is not None else None for v in bs]
    hh, ww = h, w
    for p in pools:
        if p:
            hh //= 2
            ww //= 2
    cout = [v.shape[0] for v in ws]
    y = ctx.alloc(n * cout[-1] * hh * ww)
    capi.check(lib.tp_buf_fill(ctx.h, y.h, -7.0, y.n))
    W = (C.c_void_p * L)(*[b.h for b in wb])
    B = (C.c_void_p * L)(*[(b.h if b is not None else None) for b in bb])
    co = (C.c_int * L)(*cout)
    po = (C.c_int * L)(*[int(p) for p in pools])
    re = (C.c_int * L)(*[int(r) for r in relus])
    call = lambda: capi.check(lib.tp_conv_stack_fwd(ctx.h, xb.h, n, c, h, w, L, W, B, co, po, re, y.h))
    call()
    ctx.sync()
    out = y.download().reshape(n, cout[-1], hh, ww)
    us = None
    if reps:
        for _ in range(3):
            call()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        ctx.sync()
        us = (time.perf_counter() - t0) / reps * 1e6
    return out, us


def make(rng, n, c0, hw, couts):
    x = rng.random((n, c0, hw, hw)).astype(F32)
    ws, bs = [], []
    ci = c0
    for co in couts:
        ws.append((rng.standard_normal((co, ci, 3, 3)) * np.sqrt(2.0 / (ci * 9))).astype(F32))
        bs.append((rng.standard_normal(co) * 0.05).astype(F32))
        ci = co
    return x, ws, bs


def report(tag, got, ref):
    scale = float(np.abs(ref).max())
    err = np.abs(got - ref)
    rel = float(err.max()) / max(scale, 1e-6)
    bad = int((err > 1e-4 * scale).sum())
    print(f"  {tag}: rel_inf {rel:.3e}  bad {bad}/{ref.size}  nan {int(np.isnan(got).sum())}  untouched {int((got == -7.0).sum())}", flush=True)
    if bad:
        idx = np.argwhere(err > 1e-4 * scale)
        print("    first bad (n, c, y, x):", idx[:6].tolist(), " rows y:", sorted(set(idx[:, 2].tolist()))[:12], " cols x:", sorted(set(idx[:, 3].tolist()))[:12])
    return rel


CASES = [
    # name, n, c0, hw, couts, pools
    ("32->32 @28", 3, 32, 28, [32], [0]),
    ("32->32 @28 pool", 3, 32, 28, [32], [1]),
    ("32->64 @14", 5, 32, 14, [64], [0]),
    ("64->64 @14 pool", 5, 64, 14, [64], [1]),
    ("64->128 @7", 5, 64, 7, [128], [0]),
    ("64->32 @20", 2, 64, 20, [32], [0]),
    ("32->64 @5 pool", 7, 32, 5, [64], [1]),
    ("cnn5 stack", 5, 1, 28, [32, 32, 64, 64, 128], [0, 1, 0, 1, 0]),
    ("cnn2 stack", 6, 1, 28, [32, 64], [1, 1]),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    import taper_b200
    from taper_b200 import capi
    ctx = taper_b200.Ctx(0)
    rng = np.random.default_rng(0)
    for mode in (0, 1, 2):
        capi.lib.tpdbg_conv_shift_mode(mode)
        print(f"== shift mode {mode}", flush=True)
        for name, n, c0, hw, couts, pools in CASES:
            x, ws, bs = make(rng, n, c0, hw, couts)
            relus = [1] * len(couts)
            ref = oracle_stack(x, ws, bs, pools, relus)
            try:
                got, _ = run_stack(ctx, x, ws, bs, pools, relus)
                report(name, got, ref)
            except Exception as e:  # noqa: BLE001
                print(f"  {name}: FAILED {e}", flush=True)
        x, ws, bs = make(rng, 3, 32, 14, [64])
        ref = oracle_stack(x, ws, [None], [0], [0])
        got, _ = run_stack(ctx, x, ws, [None], [0], [0])
        report("32->64 @14 no bias no relu", got, ref)
    capi.lib.tpdbg_conv_shift_mode(-1)           # the library's default variant
    if a.big:
        x, ws, bs = make(rng, 256, 1, 28, [32, 32, 64, 64, 128])
        pools = [0, 1, 0, 1, 0]
        got, us = run_stack(ctx, x, ws, bs, pools, [1] * 5, reps=20)
        print(f"cnn5 stack batch 256: {us:.1f} us", flush=True)
        report("cnn5 stack batch 256", got, oracle_stack(x, ws, bs, pools, [1] * 5))
    if a.time:
        # in-stack time of a planes -> planes layer = (k-layer stack - 1-layer stack) / (k - 1); dbg flags: 1 = hi*hi products only,
        # 2 = epilogue stores nothing
        for batch in (256, 1024):
            for (ci, hw, co) in [(32, 28, 32), (32, 14, 64), (64, 14, 64), (64, 7, 128), (128, 7, 128)]:
                line = f"  {ci}->{co} @{hw} batch {batch}:"
                for flags in (0, 1, 2, 3):
                    capi.lib.tpdbg_conv_flags(flags)
                    ts = []
                    for k in (1, 4):
                        x1, w1, b1 = make(rng, batch, ci, hw, [co] + [co] * (k - 1) if ci == co else [co])
                        if ci != co:
                            # chain: ci -> co, then (k-1) x co -> co is a different layer; time ci -> co through a ci -> co -> co ... stack instead
                            pass
                        ts.append(run_stack(ctx, x1, w1, b1, [0] * len(w1), [1] * len(w1), reps=10)[1])
                    if ci == co:
                        per = (ts[1] - ts[0]) / 3
                        fl = 2.0 * batch * hw * hw * ci * 9 * co
                        line += f"  flags {flags}: {per:7.1f} us/layer ({fl / per / 1e6:6.1f} TF/s)"
                    else:
                        line += f"  flags {flags}: {ts[0]:7.1f} us single call"
                print(line, flush=True)
            capi.lib.tpdbg_conv_flags(0)
            x, ws, bs = make(rng, batch, 1, 28, [32, 32, 64, 64, 128])
            _, us = run_stack(ctx, x, ws, bs, [0, 1, 0, 1, 0], [1] * 5, reps=20)
            print(f"  cnn5 stack batch {batch}: {us:.1f} us", flush=True)
            x, ws, bs = make(rng, batch, 1, 28, [32, 64])
            _, us = run_stack(ctx, x, ws, bs, [1, 1], [1] * 2, reps=20)
            print(f"  cnn2 stack batch {batch}: {us:.1f} us", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
