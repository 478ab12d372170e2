import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, taper_b200
from taper_b200 import capi
ctx = taper_b200.Ctx(0)
rng = np.random.default_rng(0)
for dist in ("normal", "uniform01"):
  for (m, n) in ((2560, 1024), (512, 128)):
    for k in (64, 256, 784, 2048, 8192):
        A = (rng.standard_normal((m, k)) if dist == "normal" else rng.random((m, k))).astype(np.float32)
        B = (rng.standard_normal((k, n)) if dist == "normal" else rng.random((k, n))).astype(np.float32)
        ref = A.astype(np.float64) @ B.astype(np.float64)
        out = []
        for mode in (0, 1, 2):
            capi.check(capi.lib.tp_set_gemm_mode(ctx.h, mode))
            a = ctx.upload(A); b = ctx.upload(B); c = ctx.zeros(m * n)
            ctx.call("sgemm_rowmajor", 0, 0, m, n, k, 1.0, a, b, 0.0, c)
            got = c.download().reshape(m, n).astype(np.float64)
            out.append((np.max(np.abs(got - ref)) / np.max(np.abs(ref)), np.mean(got - ref) / np.max(np.abs(ref))))
        print(f"{dist:9s} m={m} n={n} k={k:5d}  fp32 {out[0][0]:.2e} (bias {out[0][1]:+.1e})  3xTF32 {out[1][0]:.2e} (bias {out[1][1]:+.1e})  1xTF32 {out[2][0]:.2e} (bias {out[2][1]:+.1e})", flush=True)
