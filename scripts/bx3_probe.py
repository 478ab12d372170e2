"""bf16x3 GEMM probe (development): error of tp_sgemm_rowmajor in mode 3 against an fp64 product for all four operand
majors, and graph-timed throughput at the cfg4 shapes.  TAPER_BX3_BN / TAPER_BX3_SPLITS force the tile / K-split."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import taper_b200
from taper_b200 import capi
lib = capi.lib
ctx = taper_b200.Ctx(0)
what = sys.argv[1] if len(sys.argv) > 1 else "both"
rng = np.random.default_rng(0)


def run(m, n, k, ta, tb, mode, a, b, beta=0.0, c0=None):
    da = ctx.upload(a.T.copy() if ta else a)
    db = ctx.upload(b.T.copy() if tb else b)
    dc = ctx.upload(c0) if c0 is not None else ctx.zeros(m * n)
    capi.check(lib.tp_set_gemm_mode(ctx.h, mode))
    ctx.call("sgemm_rowmajor", ta, tb, m, n, k, 1.0, da, db, beta, dc)
    return dc.download(m * n).reshape(m, n)


if what in ("err", "both"):
    worst = {1: 0.0, 3: 0.0}
    for (m, n, k) in [(128, 128, 64), (128, 64, 128), (256, 256, 256), (200, 136, 72), (1024, 1024, 784), (1024, 784, 1024),
                      (1024, 1024, 1024), (512, 128, 784), (384, 264, 2048), (128, 256, 8192), (1000, 520, 1000)]:
        for dist in ("normal", "positive"):
            a = (rng.standard_normal((m, k)) if dist == "normal" else rng.random((m, k))).astype(np.float32)
            b = (rng.standard_normal((k, n)) if dist == "normal" else rng.random((k, n))).astype(np.float32)
            ref = a.astype(np.float64) @ b.astype(np.float64)
            scale = np.abs(ref).max()
            for ta in (0, 1):
                for tb in (0, 1):
                    if (ta and m % 8) or (not ta and k % 8) or (tb and k % 8) or (not tb and n % 8):
                        continue
                    line = f"m={m:5d} n={n:5d} k={k:5d} ta={ta} tb={tb} {dist:8s}:"
                    for mode in (3, 1):
                        got = run(m, n, k, ta, tb, mode, a, b)
                        e = np.abs(got - ref).max() / scale
                        worst[mode] = max(worst[mode], e)
                        line += f"  mode{mode} err {e:.2e}"
                    print(line, flush=True)
    # beta = 1 accumulate
    m, n, k = 256, 128, 512
    a = rng.standard_normal((m, k)).astype(np.float32); b = rng.standard_normal((k, n)).astype(np.float32)
    c0 = rng.standard_normal((m, n)).astype(np.float32)
    got = run(m, n, k, 0, 0, 3, a, b, 1.0, c0)
    ref = a.astype(np.float64) @ b.astype(np.float64) + c0
    print(f"beta=1: err {np.abs(got - ref).max() / np.abs(ref).max():.2e}")
    print(f"WORST mode3 {worst[3]:.2e}  mode1 {worst[1]:.2e}")

if what in ("time", "both"):
    reps = int(os.environ.get("REPS", "50"))
    shapes = [(1024, 1024, 784, 0, 1), (1024, 1024, 1024, 0, 1), (1024, 1024, 1024, 0, 0), (1024, 1024, 1024, 1, 0), (1024, 784, 1024, 1, 0),
              (4096, 4096, 4096, 0, 1), (8192, 8192, 8192, 0, 1)]
    for (m, n, k, ta, tb) in shapes:
        a, b, c = ctx.alloc(m * k), ctx.alloc(k * n), ctx.alloc(m * n)
        capi.check(lib.tp_buf_fill(ctx.h, a.h, 0.5, m * k)); capi.check(lib.tp_buf_fill(ctx.h, b.h, 0.25, k * n))
        for mode in (3, 1):
            capi.check(lib.tp_set_gemm_mode(ctx.h, mode))
            for _ in range(3):
                ctx.call("sgemm_rowmajor", ta, tb, m, n, k, 1.0, a, b, 0.0, c)
            e0, e1 = C.c_void_p(), C.c_void_p()
            capi.check(lib.tp_event_create(ctx.h, C.byref(e0))); capi.check(lib.tp_event_create(ctx.h, C.byref(e1)))
            ctx.sync()
            r = reps if m < 4096 else 5
            capi.check(lib.tp_graph_begin(ctx.h))
            for _ in range(r):
                ctx.call("sgemm_rowmajor", ta, tb, m, n, k, 1.0, a, b, 0.0, c)
            g = C.c_void_p(); capi.check(lib.tp_graph_end(ctx.h, C.byref(g)))
            capi.check(lib.tp_graph_launch(ctx.h, g)); ctx.sync()
            capi.check(lib.tp_event_record(ctx.h, e0))
            capi.check(lib.tp_graph_launch(ctx.h, g))
            capi.check(lib.tp_event_record(ctx.h, e1)); capi.check(lib.tp_event_sync(e1))
            capi.check(lib.tp_graph_destroy(g))
            ms = C.c_float(); capi.check(lib.tp_event_elapsed_ms(e0, e1, C.byref(ms)))
            us = ms.value / r * 1e3
            if mode == 3 and os.environ.get("DBG"):
                t = (C.c_longlong * 16)()
                ctx.sync(); lib.tpdbg_bx3_times(t)
                names = ["entry", "setup", "pdl", "1st stage", "last stage", "acc full", "staged", "sync1", "stored", "end"]
                print("    cycles: " + ", ".join(f"{nm}={t[i]-t[0]}" for i, nm in enumerate(names)))
            print(f"m={m} n={n} k={k} ta={ta} tb={tb} mode={mode}: {us:9.2f} us per call (mode 3 includes the two operand-split launches)  "
                  f"{2.0*m*n*k/us/1e6:9.2f} TFLOP/s", flush=True)
