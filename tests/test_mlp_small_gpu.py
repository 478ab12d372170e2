"""tp_mlp_small_fwd / tp_mlp_small_bwd (mlp_small.cu): a chain of small Linear(+ReLU) layers as one forward and one backward
launch, against the oracle's per-layer tape (Linear::forward src/nn.rs:54-60, matmul backward src/ops.rs:254-291, add_broadcast
backward src/tensor.rs:680-691, ReLU src/ops.rs:312-374) — activations, input gradient, every weight / bias gradient, first-touch
and accumulating writes, ragged batches — and the Sequential peephole that feeds it (the example CNN's 128-128-64-10 head).
Exact fp32 FFMA in ascending k: the tolerance is summation-order noise, 1e-5 of the reference's max."""
import ctypes as C

import numpy as np
import pytest

from oracle import taper_ref as R

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.fixture()
def ctx():
    import taper_b200
    c = taper_b200.Ctx(0)
    yield c
    c.close()


def close(got, ref, tol, what):
    scale = max(float(np.abs(ref).max()), 1e-6)
    err = float(np.abs(np.asarray(got, np.float64).reshape(-1) - np.asarray(ref, np.float64).reshape(-1)).max())
    assert err <= tol * scale, f"{what}: max |diff| {err:.3e} > {tol:g} * {scale:.3e}"


def oracle_chain(x, ws, bs, relus, gout):
    R.Tape.reset()
    X = R.Tensor.new(x, x.shape).requires_grad_()
    _ = X.reshape(x.shape)                                       # node 0 is the reference's sentinel
    Ws = [R.Tensor.new(w, w.shape).requires_grad_() for w in ws]
    Bs = [R.Tensor.new(b, b.shape).requires_grad_() if b is not None else None for b in bs]
    h, acts = X, []
    for W, B, relu in zip(Ws, Bs, relus):
        h = h.matmul(W.transpose())
        if B is not None:
            h = h.add_broadcast(B)
        if relu:
            h = h.relu()
        acts.append(h.data().copy().reshape(h.shape))
    h._grad[0] = gout.reshape(-1).copy()
    R.tape_backward(len(R.Tape.nodes) - 1)
    out = (acts, X.grad().copy(), [W.grad().copy() for W in Ws], [B.grad().copy() if B is not None else None for B in Bs])
    R.Tape.reset()
    return out


CASES = [
    # batch, dims, relus, biases
    (1024, [128, 128, 64, 10], [1, 1, 0], True),        # the example CNN's head at configs[4]'s batch
    (37, [128, 128, 64, 10], [1, 1, 0], True),          # ragged last CTA
    (256, [64, 32, 10], [1, 0], True),
    (33, [100, 128], [1], False),                       # one layer, ReLU on the output, no bias
    (8, [16, 120, 7, 128, 3], [1, 1, 1, 0], True),      # four layers, odd widths
]


@pytest.mark.parametrize("batch,dims,relus,bias", CASES)
def test_mlp_small_chain_vs_oracle(ctx, batch, dims, relus, bias):
    from taper_b200 import capi
    lib = capi.lib
    rng = np.random.default_rng(batch + sum(dims))
    L = len(dims) - 1
    x = (rng.random((batch, dims[0])) - 0.3).astype(F32)
    ws = [(rng.standard_normal((dims[l + 1], dims[l])) * np.sqrt(2.0 / dims[l])).astype(F32) for l in range(L)]
    bs = [(rng.standard_normal(dims[l + 1]) * 0.1).astype(F32) if bias else None for l in range(L)]
    gout = (rng.standard_normal((batch, dims[-1])) / batch).astype(F32)
    acts_ref, dx_ref, dw_ref, db_ref = oracle_chain(x, ws, bs, relus, gout)
    assert lib.tp_mlp_small_supported(L, (C.c_int * (L + 1))(*dims), batch) == 1
    xb, gb = ctx.upload(x), ctx.upload(gout)
    wb = [ctx.upload(w) for w in ws]
    bb = [ctx.upload(b) if b is not None else None for b in bs]
    ab = [ctx.alloc(batch * dims[l + 1]) for l in range(L)]
    D = (C.c_int * (L + 1))(*dims)
    Wp = (C.c_void_p * L)(*[b.h for b in wb])
    Bp = (C.c_void_p * L)(*[(b.h if b is not None else None) for b in bb])
    Ap = (C.c_void_p * L)(*[b.h for b in ab])
    Re = (C.c_int * L)(*relus)
    capi.check(lib.tp_mlp_small_fwd(ctx.h, xb.h, L, D, Wp, Bp, Re, Ap, batch))
    for l in range(L):
        close(ab[l].download(), acts_ref[l], 1e-5, f"activation {l}")
    dx = ctx.alloc(x.size)
    dw = [ctx.alloc(w.size) for w in ws]
    db = [ctx.alloc(dims[l + 1]) if bias else None for l in range(L)]
    DW = (C.c_void_p * L)(*[b.h for b in dw])
    DB = (C.c_void_p * L)(*[(b.h if b is not None else None) for b in db])
    zero, one = (C.c_int * L)(*([0] * L)), (C.c_int * L)(*([1] * L))
    capi.check(lib.tp_mlp_small_bwd(ctx.h, xb.h, L, D, Wp, Re, Ap, gb.h, dx.h, DW, DB, 0, zero, zero, batch))
    close(dx.download(), dx_ref, 1e-5, "dx")
    for l in range(L):
        close(dw[l].download(), dw_ref[l], 1e-5, f"dW {l}")
        if bias:
            close(db[l].download(), db_ref[l], 1e-5, f"db {l}")
    # accumulating writes (a live gradient, src/ops.rs:250-253) and no input gradient requested
    capi.check(lib.tp_mlp_small_bwd(ctx.h, xb.h, L, D, Wp, Re, Ap, gb.h, None, DW, DB, 1, one, one, batch))
    for l in range(L):
        close(dw[l].download(), 2 * dw_ref[l], 1e-5, f"dW {l} accumulated")
        if bias:
            close(db[l].download(), 2 * db_ref[l], 1e-5, f"db {l} accumulated")
    close(dx.download(), dx_ref, 1e-5, "dx untouched")


def test_widths_above_128_are_not_supported():
    from taper_b200 import capi
    assert capi.lib.tp_mlp_small_supported(2, (C.c_int * 3)(784, 128, 10), 64) == 0
    assert capi.lib.tp_mlp_small_supported(5, (C.c_int * 6)(8, 8, 8, 8, 8, 8), 64) == 0


def test_sequential_peephole_on_the_cnn_head_matches_per_layer_and_oracle():
    """The example CNN's head as a model of its own (features in, logits out): loss, every gradient and the launch count with
    the small-MLP peephole on and off, and against the oracle."""
    from taper_b200 import host
    host.config(conv_full_adjoint=0, fuse_linear_relu=1, reference_op_sequence=0, gemm_mode=1)
    spec = "linear:128:128,relu,linear:128:64,relu,linear:64:10"
    rng = np.random.default_rng(10)
    ref = R.build_mlp([128, 128, 64, 10], np.random.default_rng(9))
    x = rng.random((96, 128)).astype(F32)
    y = rng.integers(0, 10, 96).astype(F32)
    # no pre-activation within rounding distance of the ReLU threshold: a unit at |z| ~ 1e-7 is on for one summation order and
    # off for another (seed 5 has z = +9.1e-8 in fp64 that the oracle's fp32 sum rounds to -1.0e-7), and one flipped unit moves
    # its bias gradient by 1/sqrt(batch).  The seed is chosen so the comparison below is about the kernels.
    h = x.astype(np.float64)
    for lin in [l for l in ref.layers if hasattr(l, "weight")][:2]:
        z = h @ np.asarray(lin.weight.data(), np.float64).reshape(lin.weight.shape).T + np.asarray(lin.bias.data(), np.float64)
        assert np.abs(z).min() > 5e-5
        h = np.maximum(z, 0.0)
    R.Tape.reset()
    l_ref = R.cross_entropy_loss(ref.forward(R.Tensor.new(x, x.shape)), R.Tensor.new(y, y.shape))
    l_ref.backward()
    res, launches = [], []
    for fuse in (1, 0):
        host.config_small_mlp(fuse)
        m = host.Model(spec, 0)
        m.load_from_oracle(ref)
        m.zero_grad()
        m.loss_backward(x, y)                                               # sizes the allocator caches
        m.zero_grad()
        l0 = host.launches()
        loss, correct, _ = m.loss_backward(x, y)
        launches.append(host.launches() - l0)
        res.append((loss, correct, [m.get_grad(j) for j in range(m.num_params())]))
    host.config_small_mlp(1)
    assert launches[0] + 8 <= launches[1], launches
    for k in (0, 1):
        assert res[k][0] == pytest.approx(float(l_ref.data()[0]), rel=1e-5)
        for j, p in enumerate(ref.parameters()):
            close(res[k][2][j], p.grad(), 1e-4, f"path {k} grad {j}")
    R.Tape.reset()
