"""Times tp_sgemm_rowmajor for a few shapes / modes with CUDA events (development probe)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import taper_b200
from taper_b200 import capi
lib = capi.lib
ctx = taper_b200.Ctx(0)
reps = int(os.environ.get("REPS", "50"))
shapes = [(512, 128, 784, 0, 1), (128, 784, 512, 1, 0), (1024, 1024, 784, 0, 1), (1024, 1024, 1024, 0, 0), (4096, 4096, 4096, 0, 1),
          (8192, 8192, 8192, 0, 1), (50176, 64, 288, 0, 0)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in sys.argv[1].split(","))]
for (m, n, k, ta, tb) in shapes:
    a, b, c = ctx.alloc(m * k), ctx.alloc(k * n), ctx.alloc(m * n)
    capi.check(lib.tp_buf_fill(ctx.h, a.h, 0.5, m * k)); capi.check(lib.tp_buf_fill(ctx.h, b.h, 0.25, k * n))
    for mode in (2, 1, 0):
        if mode == 0 and m * n * k > 2e10:
            continue
        capi.check(lib.tp_set_gemm_mode(ctx.h, mode))
        for _ in range(3):
            ctx.call("sgemm_rowmajor", ta, tb, m, n, k, 1.0, a, b, 0.0, c)
        e0, e1 = C.c_void_p(), C.c_void_p()
        capi.check(lib.tp_event_create(ctx.h, C.byref(e0))); capi.check(lib.tp_event_create(ctx.h, C.byref(e1)))
        ctx.sync()
        # capture `reps` launches into one CUDA graph: measures GPU time, not the host launch path
        capi.check(lib.tp_graph_begin(ctx.h))
        for _ in range(reps):
            ctx.call("sgemm_rowmajor", ta, tb, m, n, k, 1.0, a, b, 0.0, c)
        g = C.c_void_p(); capi.check(lib.tp_graph_end(ctx.h, C.byref(g)))
        capi.check(lib.tp_graph_launch(ctx.h, g)); ctx.sync()
        capi.check(lib.tp_event_record(ctx.h, e0))
        capi.check(lib.tp_graph_launch(ctx.h, g))
        capi.check(lib.tp_event_record(ctx.h, e1)); capi.check(lib.tp_event_sync(e1))
        capi.check(lib.tp_graph_destroy(g))
        ms = C.c_float(); capi.check(lib.tp_event_elapsed_ms(e0, e1, C.byref(ms)))
        us = ms.value / reps * 1e3
        print(f"m={m} n={n} k={k} ta={ta} tb={tb} mode={mode}: {us:9.2f} us  {2.0*m*n*k/us/1e6:9.2f} TFLOP/s", flush=True)
        if mode and os.environ.get("DBG"):
            t = (C.c_longlong * 16)()
            ctx.sync(); lib.tpdbg_gemm_times(t)
            names = ["start", "setup", "1st tile ready", "last tile ready", "acc staged", "sync1 done", "stored", "end"]
            print("    cycles: " + ", ".join(f"{n}={t[i]-t[0]}" for i, n in enumerate(names)))
