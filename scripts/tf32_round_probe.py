"""Does tcgen05 kind::tf32 truncate or round its fp32 inputs?  A = 1 + 2^-11 + 2^-12 (RN -> 1 + 2^-10, RZ -> 1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, taper_b200
from taper_b200 import capi
ctx = taper_b200.Ctx(0)
m = n = 128; k = 32
for val, name in ((1 + 2.0**-11 + 2.0**-12, "above half"), (1 + 2.0**-11, "exact half"), (1 + 2.0**-10 + 2.0**-11, "odd+half"), (-(1 + 2.0**-11 + 2.0**-12), "negative")):
    A = np.full((m, k), val, np.float32); B = np.zeros((k, n), np.float32); B[0, :] = 1.0
    capi.check(capi.lib.tp_set_gemm_mode(ctx.h, 2))
    a, b, c = ctx.upload(A), ctx.upload(B.T.copy()), ctx.zeros(m * n)
    ctx.call("sgemm_rowmajor", 0, 1, m, n, k, 1.0, a, b, 0.0, c)
    got = c.download()[0]
    print(f"{name:12s} input {val!r:24} -> {got!r}   (trunc would give {np.float32(np.frombuffer(np.uint32(np.float32(val).view(np.uint32) & 0xFFFFE000).tobytes(), np.float32)[0])!r})")
