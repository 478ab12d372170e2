import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, taper_b200
from taper_b200 import capi
ctx = taper_b200.Ctx(0)
m, n, k, ta, tb, mode = [int(x) for x in sys.argv[1].split(",")]
rng = np.random.default_rng(0)
A = rng.integers(-2, 3, (m, k)).astype(np.float32); B = rng.integers(-2, 3, (k, n)).astype(np.float32)
capi.check(capi.lib.tp_set_gemm_mode(ctx.h, mode))
a = ctx.upload((A.T if ta else A).copy()); b = ctx.upload((B.T if tb else B).copy()); c = ctx.zeros(m * n)
ctx.call("sgemm_rowmajor", ta, tb, m, n, k, 1.0, a, b, 0.0, c)
got = c.download().reshape(m, n); ref = A @ B
bad = got != ref
print("bad fraction", bad.mean())
print("bad rows:", np.unique(np.nonzero(bad)[0])[:40], "count", len(np.unique(np.nonzero(bad)[0])))
print("bad cols:", np.unique(np.nonzero(bad)[1])[:140])
i, j = np.nonzero(bad)[0][:1], np.nonzero(bad)[1][:1]
if len(i):
    print("first bad", i, j, got[i, j], ref[i, j])
    # is got a shifted/other element of ref?
    r = i[0]
    for jj in range(0, n, max(1, n // 16)):
        print(jj, got[r, jj], ref[r, jj])
