"""Wide step plan probe (development): losses of the plan vs the tape + CUDA-graph path on the same batches, then timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from taper_b200 import host
host.set_device(0)
F32 = np.float32
spec = os.environ.get("SPEC", host.MLP_784_1024_1024_10)
B = int(os.environ.get("BATCH", "1024"))
opt = os.environ.get("OPT", "adam")
rng = np.random.default_rng(0)
data = [(rng.random((B, 784)).astype(F32), rng.integers(0, 10, B).astype(F32)) for _ in range(6)]
outs = []
for fused in (False, True):
    m = host.Model(spec, 0)
    tr = host.Trainer(m, opt, lr=1e-3 if opt != "sgd" else 0.05)
    tr.set_use_fused(fused)
    res = [tr.step(x, y) for x, y in data]
    print("fused" if fused else "tape ", [f"{l:.6f}/{int(c)}" for l, c in res], "kind", tr.fused_kind(), "fused_steps", tr.fused_steps(), flush=True)
    outs.append((res, [m.get_param(i) for i in range(m.num_params())]))
for (la, ca), (lb, cb) in zip(outs[0][0], outs[1][0]):
    print(f"loss rel diff {abs(la - lb) / abs(la):.2e}  correct {ca} {cb}")
for i, (a, b) in enumerate(zip(outs[0][1], outs[1][1])):
    print(f"param {i}: max abs diff {np.abs(a - b).max():.3e} of {np.abs(a).max():.3e}")
# u8 path equals f32 path bitwise when the f32 pixels are u8 / 255
xu = rng.integers(0, 256, (B, 784)).astype(np.uint8)
y = rng.integers(0, 10, B).astype(F32)
ps = []
for u8 in (False, True):
    m = host.Model(spec, 0)
    tr = host.Trainer(m, opt, lr=1e-3 if opt != "sgd" else 0.05)
    for _ in range(3):
        if u8:
            tr.step_async_u8(xu, y, pinned=False)
        else:
            tr.step_async((xu.astype(F32) / F32(255.0)).astype(F32), y, pinned=False)
        r = tr.fetch()
    ps.append((r, [m.get_param(i) for i in range(m.num_params())]))
print("u8 vs f32:", ps[0][0], ps[1][0], "params equal:", all(np.array_equal(a, b) for a, b in zip(ps[0][1], ps[1][1])))
# timing: resident dataset
m = host.Model(spec, 0)
tr = host.Trainer(m, opt, lr=1e-3 if opt != "sgd" else 0.05)
X = rng.random((60000, 784)).astype(F32); Y = rng.integers(0, 10, 60000).astype(F32)
tr.load_dataset(X, Y, np.random.default_rng(1).permutation(60000).astype(np.uint32))
for _ in range(20):
    tr.step_resident(B); tr.fetch()
e0, e1 = host.Event(), host.Event()
host.sync(); e0.record()
n = 500
for i in range(n):
    if tr.pending() >= 6:
        tr.fetch()
    tr.step_resident(B)
e1.record()
while tr.pending():
    last = tr.fetch()
host.sync()
print(f"resident: {e0.elapsed_ms(e1) / n * 1e3:.2f} us/step, last {last}")
