"""Single launches of the kernel families the BASELINE metric names, at sizes where the roofline is meaningful, for an
`ncu --set full` capture (north_star: "each kernel choice is evidenced by a committed ncu capture"):
tp_sgemm_rowmajor 8192^3 (1xTF32 / 3xTF32 / bf16x3), 4096^3 bf16x3, fused Adam and ReLU-backward over 48 Mi elements
(> L2), max-pool forward/backward and the NCHW bias-gradient at the CNN's batch-1024 shapes."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import taper_b200
from taper_b200 import capi

lib = capi.lib
ctx = taper_b200.Ctx(0)


def fill(n, v=0.5):
    b = ctx.alloc(n)
    capi.check(lib.tp_buf_fill(ctx.h, b.h, v, n))
    return b


for nn in (8192, 4096):
    a, b, c = fill(nn * nn), fill(nn * nn, 0.25), ctx.alloc(nn * nn)
    for mode in ((2, 1, 3) if nn == 8192 else (3,)):
        capi.check(lib.tp_set_gemm_mode(ctx.h, mode))
        ctx.call("sgemm_rowmajor", 0, 1, nn, nn, nn, 1.0, a, b, 0.0, c)
    ctx.sync()
    del a, b, c
capi.check(lib.tp_set_gemm_mode(ctx.h, 1))
n = 48 * 1024 * 1024
p, g, m, v, hy = fill(n), fill(n, 0.01), fill(n, 0.0), fill(n, 0.0), ctx.alloc(8)
capi.check(lib.tp_adam_hyper_init(ctx.h, hy.h, 1e-3, 0.9, 0.999, 1e-8, 0.0))
capi.check(lib.tp_adam_advance(ctx.h, hy.h))
for _ in range(1):
    capi.check(lib.tp_adam_step_dev(ctx.h, p.h, g.h, m.h, v.h, hy.h, 1.0, 0, n))
    capi.check(lib.tp_relu_bwd(ctx.h, p.h, g.h, m.h, n, 0))
    capi.check(lib.tp_sgd_step(ctx.h, p.h, g.h, 0.01, 1.0, n))
    capi.check(lib.tp_accumulate(ctx.h, m.h, g.h, 1.0, n, 1))
ctx.sync()
del p, g, m, v
# pooling / bias gradient at the CNN's batch-1024 shapes
B = 1024
x = fill(B * 32 * 28 * 28)
y, arg, gx = ctx.alloc(B * 32 * 14 * 14), ctx.alloc(B * 32 * 14 * 14), ctx.alloc(B * 32 * 28 * 28)
d = capi.PoolDesc(B, 32, 28, 28, 2, 2, 2, 2, 0, 0)
for _ in range(1):
    ctx.call("maxpool2d_fwd", x, y, arg, d)
    ctx.call("maxpool2d_bwd", y, arg, gx, d)
ctx.sync()
print("probes done")
