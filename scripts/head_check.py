"""Where does the per-layer tape lose accuracy on the CNN head (128-128-64-10)?  Kernel-level checks of tp_linear_bwd at the head's
shapes against fp64 NumPy, per GEMM mode, and the whole model per mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import taper_b200
from taper_b200 import capi, host
from oracle import taper_ref as R

ctx = taper_b200.Ctx(0)
rng = np.random.default_rng(0)
for mode in (1, 0, 3):
    ctx.call("set_gemm_mode", mode)
    for (B, fin, fout, want_dx) in [(256, 128, 128, 0), (256, 128, 64, 1), (96, 128, 128, 0), (96, 128, 64, 1), (1024, 128, 128, 0), (1024, 128, 64, 1)]:
        x = rng.random((B, fin)).astype(np.float32)
        w = (rng.standard_normal((fout, fin)) * 0.1).astype(np.float32)
        y = np.maximum(rng.standard_normal((B, fout)), 0).astype(np.float32)
        dy = (rng.standard_normal((B, fout)) / B).astype(np.float32)
        g = dy.astype(np.float64) * (y > 0)
        dw_ref, db_ref, dx_ref = g.T @ x.astype(np.float64), g.sum(0), g @ w.astype(np.float64)
        X, W, Y, DY = ctx.upload(x), ctx.upload(w), ctx.upload(y), ctx.upload(dy)
        dx, dw, db = ctx.alloc(x.size), ctx.alloc(w.size), ctx.alloc(fout)
        ctx.call("linear_bwd", X, W, DY, Y, dx if want_dx else None, dw, db, B, fin, fout, 0, 0, 0)
        e = lambda a, r: float(np.abs(a.reshape(r.shape) - r).max() / np.abs(r).max())
        print(f"mode {mode} B={B} {fin}->{fout}: dW {e(dw.download(), dw_ref):.1e} db {e(db.download(), db_ref):.1e}"
              + (f" dX {e(dx.download(), dx_ref):.1e}" if want_dx else ""), flush=True)
ctx.close()
spec = "linear:128:128,relu,linear:128:64,relu,linear:64:10"
for mode in (1, 0, 3):
    host.config(conv_full_adjoint=0, fuse_linear_relu=1, reference_op_sequence=0, gemm_mode=mode)
    host.config_small_mlp(0)
    for batch in (96, 256):
        rng = np.random.default_rng(5)
        ref = R.build_mlp([128, 128, 64, 10], np.random.default_rng(9))
        x = rng.random((batch, 128)).astype(np.float32); y = rng.integers(0, 10, batch).astype(np.float32)
        R.Tape.reset()
        l_ref = R.cross_entropy_loss(ref.forward(R.Tensor.new(x, x.shape)), R.Tensor.new(y, y.shape)); l_ref.backward()
        for fl in (1, 0):
            host.config(fuse_linear_relu=fl)
            m = host.Model(spec, 0); m.load_from_oracle(ref); m.zero_grad()
            loss, correct, _ = m.loss_backward(x, y)
            errs = []
            for j, p in enumerate(ref.parameters()):
                g = m.get_grad(j); r = np.asarray(p.grad()).reshape(-1)
                errs.append(float(np.abs(np.asarray(g).reshape(-1) - r).max() / max(np.abs(r).max(), 1e-9)))
            print('model mode', mode, 'batch', batch, 'fuse_linear_relu', fl, ['%.1e' % e for e in errs], flush=True)
        R.Tape.reset()
host.config(fuse_linear_relu=1, gemm_mode=1)
host.config_small_mlp(1)
