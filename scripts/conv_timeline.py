"""Timeline of CTA 0's first tiles in conv3x3_bx3_kernel (SM-clock stamps, tpdbg_conv_times): where a tile's time goes between
the producer, the two MMA issuers and the epilogue warps.  usage: python scripts/conv_timeline.py [batch]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taper_b200
from taper_b200 import capi
from conv_stack_probe import make, run_stack

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = taper_b200.Ctx(0)
rng = np.random.default_rng(0)
for (ci, hw, co, pool) in [(32, 28, 32, 0), (32, 28, 32, 1), (64, 14, 64, 0), (64, 7, 128, 0)]:
    x, ws, bs = make(rng, batch, ci, hw, [co])
    for _ in range(2):
        run_stack(ctx, x, ws, bs, [pool], [1])
    buf = (C.c_longlong * 384)()
    capi.lib.tpdbg_conv_times(buf)
    t = np.array(buf[:], dtype=np.int64).reshape(6, 64)
    t0 = t[4, 0]
    print(f"== {ci}->{co} @{hw} pool={pool} batch {batch}: clocks relative to the first patch load; columns: tile, load issued, patch seen, "
          "acc free, committed, epi saw full, epi released")
    for it in range(14):
        print(f"  {it:2d}  {t[4, it] - t0:7d} {t[5, it] - t0:7d} {t[0, it] - t0:7d} {t[1, it] - t0:7d} {t[2, it] - t0:7d} {t[3, it] - t0:7d}"
              f"   | mma {t[1, it] - max(t[0, it], t[5, it]):5d}  commit->seen {t[2, it] - t[1, it]:5d}  epi {t[3, it] - t[2, it]:5d}")
    d = np.diff(t[3, 8:40])
    print(f"  steady state: {d.mean():.0f} clocks per tile (epilogue release to release)")
ctx.close()
