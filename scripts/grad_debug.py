import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import taper_ref as R
from taper_b200 import host
F32 = np.float32
sizes = [784, 128, 10]; spec = host.MLP_784_128_10; B = 512
if len(sys.argv) > 1 and sys.argv[1] == "cfg4":
    sizes = [784, 1024, 1024, 10]; spec = host.MLP_784_1024_1024_10; B = 1024
rng = np.random.default_rng(0)
ref = R.build_mlp(sizes, rng)
for p in ref.parameters():
    if len(p.shape) == 1: p._data[:] = rng.standard_normal(p._data.size).astype(F32) * F32(0.05)
m = host.Model(spec, 0); m.load_from_oracle(ref)
tr = host.Trainer(m, "adam", lr=1e-3)
opt = R.Adam(ref.parameters(), 1e-3)
rng = np.random.default_rng(1)
for step in range(8):
    x = rng.random((B, sizes[0])).astype(F32); y = rng.integers(0, 10, B).astype(F32)
    # gradients at the ORACLE's current parameters on both sides (teacher forcing for the comparison only)
    m2 = host.Model(spec, 0); m2.load_from_oracle(ref)
    m2.loss_backward(x, y)
    R.Tape.reset()
    l = R.cross_entropy_loss(ref.forward(R.Tensor.new(x, x.shape)), R.Tensor.new(y, y.shape)); l.backward()
    for i, p in enumerate(ref.parameters()):
        g_ref = p.grad(); g = m2.get_grad(i).reshape(-1)
        err = np.abs(g - g_ref); k = int(np.argmax(err))
        print(f"step {step} param {i}: |g|max {np.abs(g_ref).max():.2e} max err {err.max():.2e} at {k} (ref {g_ref[k]:.3e} got {g[k]:.3e}); zeros ref {np.sum(g_ref==0)} got {np.sum(g==0)}; |g|<1e-7: {np.sum(np.abs(g_ref)<1e-7)}")
    for p in ref.parameters(): p.zero_grad()
    R.train_step(ref, opt, R.Tensor.new(x, x.shape), R.Tensor.new(y, y.shape))
    tr.step(x, y)
    for i, p in enumerate(ref.parameters()):
        d = np.abs(m.get_param(i).reshape(-1) - p.data()); k = int(np.argmax(d))
        print(f"   after step {step} param {i}: max param diff {d.max():.3e} at {k}  frac>1e-5: {np.mean(d>1e-5):.2e}")
