"""Per-layer intermediates of the CNN head (three single-layer models chained through the Tensor API) against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from taper_b200 import host
from oracle import taper_ref as R

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
batch = 96
host.config(conv_full_adjoint=0, fuse_linear_relu=1, reference_op_sequence=0, gemm_mode=mode)
host.config_small_mlp(0)
rng = np.random.default_rng(5)
ref = R.build_mlp([128, 128, 64, 10], np.random.default_rng(9))
x = rng.random((batch, 128)).astype(np.float32); y = rng.integers(0, 10, batch).astype(np.float32)
R.Tape.reset()
X = R.Tensor.new(x, x.shape)
h, acts = X, []
for l in ref.layers if hasattr(ref, 'layers') else ref.modules:
    h = l.forward(h); acts.append(h)
l_ref = R.cross_entropy_loss(h, R.Tensor.new(y, y.shape)); l_ref.backward()
print('oracle layers', [type(l).__name__ for l in (ref.layers if hasattr(ref, 'layers') else ref.modules)])
ms = [host.Model("linear:128:128,relu", 0), host.Model("linear:128:64,relu", 0), host.Model("linear:64:10", 0)]
ps = ref.parameters()
for i, m in enumerate(ms):
    m.set_param(0, np.asarray(ps[2 * i].data())); m.set_param(1, np.asarray(ps[2 * i + 1].data())); m.zero_grad()
host.tape_reset()
t = host.Tensor(x, x.shape)
outs = []
for m in ms:
    t = m.forward_tensor(t); outs.append(t)
L = host.loss("cross_entropy", t, host.Tensor(y, y.shape))
L.backward()
e = lambda a, r: float(np.abs(np.asarray(a, np.float64).reshape(-1) - np.asarray(r, np.float64).reshape(-1)).max() / max(np.abs(np.asarray(r)).max(), 1e-12))
print('loss', float(L.data()[0]), float(l_ref.data()[0]))
oa = [a for a in acts if True]
print('n oracle acts', len(oa), 'shapes', [a.shape for a in oa])
# oracle activations after each (linear, relu) pair
pairs = []
names = [type(l).__name__ for l in (ref.layers if hasattr(ref, 'layers') else ref.modules)]
idx = [i for i, n in enumerate(names) if n == 'Linear']
post = []
for k, i in enumerate(idx):
    j = i + 1 if i + 1 < len(names) and names[i + 1] == 'ReLU' else i
    post.append(oa[j])
for k in range(3):
    g = outs[k].grad(); rg = post[k].grad()
    print('layer', k, 'act err', e(outs[k].data(), post[k].data()), 'out-grad err', None if g is None or rg is None else e(g, rg))
    if g is not None and rg is not None:
        d = np.abs(np.asarray(g).reshape(batch, -1) - np.asarray(rg).reshape(batch, -1))
        bad = np.argwhere(d > 1e-4 * np.abs(rg).max())
        print('   bad elements', len(bad), bad[:10].tolist())
        for (r, c) in bad[:5]:
            print('    ', r, c, 'got', np.asarray(g).reshape(batch, -1)[r, c], 'ref', np.asarray(rg).reshape(batch, -1)[r, c], 'act', np.asarray(outs[k].data()).reshape(batch, -1)[r, c], 'ref act', np.asarray(post[k].data()).reshape(batch, -1)[r, c])
    print('   dW err', e(ms[k].get_grad(0), ps[2 * k].grad()), 'db err', e(ms[k].get_grad(1), ps[2 * k + 1].grad()))
