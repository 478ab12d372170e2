"""GPU parity tests of every C-ABI kernel entry point against the CPU oracle (oracle/taper_ref.py)
on identical seeded inputs.  Bit-exact for integer/index work and for fp32 ops whose operation
order is fixed (add/mul/relu/optimizer formulas); tolerance 1e-4 relative (north_star) elsewhere:
    max|x - ref| <= 1e-4 * max(||ref||_inf, 1e-6)
"""
import ctypes as C

import numpy as np
import pytest

from oracle import taper_ref as R

pytestmark = pytest.mark.gpu

F32 = np.float32
SIZES = [1, 3, 4, 7, 1023, 4096, (1 << 20) + 3]


def close(got, ref, tol=1e-4):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    assert got.shape == ref.shape
    scale = max(np.max(np.abs(ref)) if ref.size else 0.0, 1e-6)
    err = np.max(np.abs(got - ref)) if ref.size else 0.0
    assert err <= tol * scale, f"max abs err {err:.3e} > {tol} * {scale:.3e}"


def exact(got, ref):
    np.testing.assert_array_equal(np.asarray(got).reshape(-1), np.asarray(ref, dtype=np.asarray(got).dtype).reshape(-1))


def rnd(rng, *shape):
    return rng.standard_normal(shape).astype(F32)


# ------------------------------------------------------------------------------------------------
# runtime
# ------------------------------------------------------------------------------------------------
def test_buffers_roundtrip_slice_copy_fill(ctx):
    from taper_b200 import capi
    rng = np.random.default_rng(0)
    a = rnd(rng, 100003)
    b = ctx.upload(a)
    exact(b.download(), a)
    c = ctx.alloc(100003)
    ctx.call("buf_copy", c, b, 100003)
    exact(c.download(), a)
    ctx.call("buf_fill", c, 2.5, 1000)
    exact(c.download()[:1000], np.full(1000, 2.5, F32))
    exact(c.download()[1000:], a[1000:])
    s = C.c_void_p()
    capi.check(capi.lib.tp_buf_slice(b.h, 10, 50, C.byref(s)))
    sl = capi.Buf(ctx, s, 50)
    exact(sl.download(), a[10:60])
    assert capi.lib.tp_buf_slice(b.h, 100000, 50, C.byref(s)) == 1       # out of range -> TP_ERR_INVALID
    with pytest.raises(capi.TaperError):
        ctx.call("add", b, b, ctx.alloc(4), 100)                           # short output buffer


def test_large_pageable_upload_staging(ctx):
    rng = np.random.default_rng(1)
    a = rnd(rng, (9 << 20) // 4 + 5)          # larger than the 8 MiB pinned ring
    exact(ctx.upload(a).download(), a)


def test_graph_capture_and_replay(ctx):
    from taper_b200 import capi
    a = ctx.upload(np.ones(1024, F32))
    acc = ctx.zeros(1024)
    ctx.call("accumulate", acc, a, 1.0, 1024, 1)     # warm (no allocation inside capture)
    ctx.call("buf_fill", acc, 0.0, 1024)
    ctx.sync()
    capi.check(capi.lib.tp_graph_begin(ctx.h))
    ctx.call("accumulate", acc, a, 1.0, 1024, 1)
    ctx.call("accumulate", acc, a, 2.0, 1024, 1)
    g = C.c_void_p()
    capi.check(capi.lib.tp_graph_end(ctx.h, C.byref(g)))
    before = ctx.launches()
    for _ in range(5):
        capi.check(capi.lib.tp_graph_launch(ctx.h, g))
    assert ctx.launches() - before == 10
    exact(acc.download(), np.full(1024, 15.0, F32))
    capi.check(capi.lib.tp_graph_destroy(g))


# ------------------------------------------------------------------------------------------------
# elementwise family (src/tensor.rs:14-234, src/ops.rs:8-151, 312-496)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", SIZES)
def test_binary_ops_bit_exact(ctx, n):
    rng = np.random.default_rng(n)
    a, b = rnd(rng, n), rnd(rng, n)
    b[np.abs(b) < 1e-3] = 1.0
    A, B, O = ctx.upload(a), ctx.upload(b), ctx.alloc(n)
    for name, ref in (("add", a + b), ("sub", a - b), ("mul", a * b), ("div", a / b)):
        ctx.call(name, A, B, O, n)
        exact(O.download(), ref)


def test_binary_ops_unaligned_slices(ctx):
    """Operands at non-16-byte-aligned addresses take the scalar path and stay exact."""
    from taper_b200 import capi
    rng = np.random.default_rng(5)
    a, b = rnd(rng, 1001), rnd(rng, 1001)
    A, B, O = ctx.upload(a), ctx.upload(b), ctx.alloc(1001)
    sa, sb = C.c_void_p(), C.c_void_p()
    capi.check(capi.lib.tp_buf_slice(A.h, 1, 1000, C.byref(sa)))
    capi.check(capi.lib.tp_buf_slice(B.h, 1, 1000, C.byref(sb)))
    SA, SB = capi.Buf(ctx, sa, 1000), capi.Buf(ctx, sb, 1000)
    ctx.call("add", SA, SB, O, 1000)
    exact(O.download()[:1000], a[1:] + b[1:])


@pytest.mark.parametrize("n", SIZES)
def test_accumulate_and_backward_family(ctx, n):
    rng = np.random.default_rng(100 + n)
    g, a, b, d0 = rnd(rng, n), rnd(rng, n), rnd(rng, n), rnd(rng, n)
    b[np.abs(b) < 1e-2] = 1.0
    G, A, B = ctx.upload(g), ctx.upload(a), ctx.upload(b)
    D = ctx.upload(d0)
    ctx.call("accumulate", D, G, 1.0, n, 1)                 # accumulate_grad  src/ops.rs:124-137
    exact(D.download(), d0 + g)
    ctx.call("accumulate", D, G, -1.0, n, 0)                # first touch, scaled  src/ops.rs:140-151
    exact(D.download(), F32(-1.0) * g)
    D.upload(d0); ctx.call("accumulate", D, G, 0.5, n, 1)
    exact(D.download(), d0 + F32(0.5) * g)
    D.upload(d0); ctx.call("mul_bwd", G, B, D, n, 1)        # src/ops.rs:79-115
    exact(D.download(), d0 + g * b)
    D.upload(d0); ctx.call("div_bwd_a", G, B, D, n, 1)      # src/ops.rs:466-476
    exact(D.download(), d0 + g / b)
    D.upload(d0); ctx.call("div_bwd_b", G, A, B, D, n, 1)   # src/ops.rs:478-491
    exact(D.download(), d0 + (-(g * a / (b * b))))
    ctx.call("relu_fwd", A, D, n)                           # src/ops.rs:312-349
    exact(D.download(), np.maximum(a, 0))
    D.upload(d0); ctx.call("relu_bwd", A, G, D, n, 1)       # src/ops.rs:358-370
    exact(D.download(), d0 + np.where(a > 0, g, F32(0)))
    ctx.call("relu_bwd", A, G, D, n, 0)
    exact(D.download(), np.where(a > 0, g, F32(0)))
    ctx.call("scale", D, 0.25, n)
    exact(D.download(), np.where(a > 0, g, F32(0)) * F32(0.25))


def test_relu_edge_values(ctx):
    x = np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1e-45, -1e-45, 3.0], F32)
    g = np.arange(1, 9).astype(F32)
    X, G, O = ctx.upload(x), ctx.upload(g), ctx.alloc(8)
    ctx.call("relu_bwd", X, G, O, 8, 0)
    exact(O.download(), np.where(x > 0, g, F32(0)))          # NaN > 0 is false: grad 0, like the reference


@pytest.mark.parametrize("n", [1, 5, 4096, 100003])
def test_exp_log(ctx, n):
    rng = np.random.default_rng(n)
    x = (rng.random(n) * 8 - 4).astype(F32)
    g = rnd(rng, n)
    X, G, Y, D = ctx.upload(x), ctx.upload(g), ctx.alloc(n), ctx.alloc(n)
    ctx.call("exp_fwd", X, Y, n)
    y = Y.download()
    np.testing.assert_allclose(y, np.exp(x), rtol=3e-7 * 4)       # expf: <= 2 ulp
    ctx.call("exp_bwd", Y, G, D, n, 0)
    exact(D.download(), g * y)
    p = np.abs(x) + F32(0.1)
    P = ctx.upload(p)
    ctx.call("log_fwd", P, Y, n)
    np.testing.assert_allclose(Y.download(), np.log(p), rtol=2e-6, atol=2e-7)
    ctx.call("log_bwd", P, G, D, n, 0)
    exact(D.download(), g / p)


# ------------------------------------------------------------------------------------------------
# broadcast / reduce / layout (src/tensor.rs:544-1088)
# ------------------------------------------------------------------------------------------------
SHAPES = [(1, 1), (3, 10), (64, 10), (96, 128), (512, 128), (1000, 7), (33, 1024), (2048, 16)]


@pytest.mark.parametrize("rows,cols", SHAPES)
def test_broadcast_reduce_ops(ctx, rows, cols):
    rng = np.random.default_rng(rows * 131 + cols)
    a, bias, r = rnd(rng, rows, cols), rnd(rng, cols), rnd(rng, rows)
    A, BI, RR = ctx.upload(a), ctx.upload(bias), ctx.upload(r)
    O = ctx.alloc(rows * cols)
    ctx.call("add_broadcast_fwd", A, BI, O, rows, cols, 0)
    exact(O.download(), a + bias[None, :])
    ctx.call("add_broadcast_fwd", A, BI, O, rows, cols, 1)
    exact(O.download(), np.maximum(a + bias[None, :], 0))
    ctx.call("sub_broadcast_rows_fwd", A, RR, O, rows, cols)
    exact(O.download(), a - r[:, None])
    cs = ctx.upload(bias)
    ctx.call("colsum", A, cs, rows, cols, 1.0, 1)
    close(cs.download(), bias + a.sum(axis=0, dtype=np.float64), 1e-5)
    ctx.call("colsum", A, cs, rows, cols, 1.0, 0)
    close(cs.download(), a.sum(axis=0, dtype=np.float64), 1e-5)
    rs = ctx.upload(r)
    ctx.call("rowsum", A, rs, rows, cols, -1.0, 1)
    close(rs.download(), r - a.sum(axis=1, dtype=np.float64), 1e-5)
    one = ctx.alloc(1)
    ctx.call("sum_all", A, one, rows * cols)
    close(one.download(), [a.sum(dtype=np.float64)], 1e-5 * max(1.0, np.sqrt(a.size) / 10))
    for mode, src, ref in ((0, r, np.broadcast_to(r[:, None], a.shape)), (1, bias, np.broadcast_to(bias[None, :], a.shape)),
                           (2, r[:1], np.full(a.shape, r[0], F32))):
        S = ctx.upload(src)
        O.upload(a.reshape(-1))
        ctx.call("broadcast_bwd", S, O, rows, cols, mode, 1)
        exact(O.download(), a + ref)
    ctx.call("transpose2d", A, O, rows, cols, 0)
    exact(O.download(), a.T.copy())
    ctx.call("transpose2d", A, O, rows, cols, 1)            # transpose backward accumulates  src/tensor.rs:575-586
    exact(O.download(), a.T.copy() + a.T.copy())


@pytest.mark.parametrize("rows,cols", SHAPES)
def test_max_argmax_bit_exact(ctx, rows, cols):
    rng = np.random.default_rng(rows * 7 + cols)
    a = rng.integers(-3, 4, (rows, cols)).astype(F32)          # many ties
    if a.size > 4:
        a.reshape(-1)[rng.integers(0, a.size, max(1, a.size // 50))] = np.nan
    A = ctx.upload(a)
    V, I = ctx.alloc(max(rows, cols)), ctx.alloc(max(rows, cols))
    t = R.Tensor.new(a, a.shape)
    ctx.call("max_rows", A, V, I, rows, cols)
    v, i = t.max(1)
    exact(V.download()[:rows], v.data()); exact(I.download()[:rows], i.data())
    ctx.call("max_cols", A, V, I, rows, cols)
    v, i = t.max(0)
    exact(V.download()[:cols], v.data()); exact(I.download()[:cols], i.data())


def test_max_all_last_of_equal_maxima(ctx):                    # Iterator::max_by  src/tensor.rs:1072-1080
    a = np.array([1, 5, 2, 5, 0, 5, 1], F32)
    A, V, I = ctx.upload(a), ctx.alloc(1), ctx.alloc(1)
    ctx.call("max_all", A, V, I, a.size)
    assert V.download()[0] == 5 and I.download()[0] == 5
    rng = np.random.default_rng(0)
    b = rng.integers(0, 50, 100000).astype(F32)
    B = ctx.upload(b)
    ctx.call("max_all", B, V, I, b.size)
    v, i = R.Tensor.new(b, b.shape).max(None)
    exact(V.download(), v.data()); exact(I.download(), i.data())


# ------------------------------------------------------------------------------------------------
# the reference's operator boundary: sgemm_rowmajor (src/gemm.rs:8-49, 72-119)
# ------------------------------------------------------------------------------------------------
GEMM_SHAPES = [(2, 2, 3), (64, 128, 784), (64, 10, 128), (512, 128, 784), (96, 10, 128), (784, 128, 512),
               (128, 10, 512), (512, 128, 10), (1024, 1024, 784), (1, 1, 1), (17, 33, 65), (130, 250, 9), (300, 32, 288)]


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_sgemm_rowmajor_all_transposes(ctx, m, n, k, mode):
    from taper_b200 import capi
    rng = np.random.default_rng(m * 1000003 + n * 1009 + k)
    A, B, C0 = rnd(rng, m, k), rnd(rng, k, n), rnd(rng, m, n)
    ref64 = A.astype(np.float64) @ B.astype(np.float64)
    tol = {0: 2e-6 * max(1, k) ** 0.5, 1: 1e-5 * max(1, k) ** 0.5 / 4 + 2e-6, 2: 2e-3}[mode]
    capi.check(capi.lib.tp_set_gemm_mode(ctx.h, mode))
    try:
        for ta in (0, 1):
            for tb in (0, 1):
                a = ctx.upload((A.T if ta else A).copy())
                b = ctx.upload((B.T if tb else B).copy())
                c = ctx.upload(C0)
                ctx.call("sgemm_rowmajor", ta, tb, m, n, k, 1.0, a, b, 0.0, c)
                close(c.download(), ref64, tol)
                c.upload(C0.reshape(-1))
                ctx.call("sgemm_rowmajor", ta, tb, m, n, k, 1.0, a, b, 1.0, c)       # backward form, beta = 1
                close(c.download(), ref64 + C0, tol)
                c.upload(C0.reshape(-1))
                ctx.call("sgemm_rowmajor", ta, tb, m, n, k, 0.5, a, b, -2.0, c)
                close(c.download(), 0.5 * ref64 - 2.0 * C0, tol)
    finally:
        capi.check(capi.lib.tp_set_gemm_mode(ctx.h, 1))


def test_sgemm_kat(ctx):                                       # tests/smoke.rs:46-70
    a = ctx.upload(np.array([1, 2, 3, 4, 5, 6], F32))
    b = ctx.upload(np.array([7, 8, 9, 10, 11, 12], F32))
    c = ctx.alloc(4)
    ctx.call("sgemm_rowmajor", 0, 0, 2, 2, 3, 1.0, a, b, 0.0, c)
    exact(c.download(), [58, 64, 139, 154])
    g = ctx.upload(np.ones(4, F32))
    da, db = ctx.zeros(6), ctx.zeros(6)
    ctx.call("sgemm_rowmajor", 0, 1, 2, 3, 2, 1.0, g, b, 1.0, da)      # dA += dC * B^T  src/ops.rs:254-265
    ctx.call("sgemm_rowmajor", 1, 0, 3, 2, 2, 1.0, a, g, 1.0, db)      # dB += A^T * dC  src/ops.rs:280-291
    exact(da.download(), [15, 19, 23, 15, 19, 23])
    exact(db.download(), [5, 5, 7, 7, 9, 9])


def test_sgemm_empty_dims(ctx):
    a, b, c = ctx.alloc(8), ctx.alloc(8), ctx.upload(np.full(8, 3.0, F32))
    ctx.call("sgemm_rowmajor", 0, 0, 0, 4, 2, 1.0, a, b, 0.0, c)
    exact(c.download(), np.full(8, 3.0, F32))


# ------------------------------------------------------------------------------------------------
# fused Linear (src/nn.rs:54-60) vs the oracle's transpose + matmul + add_broadcast (+ relu) tape
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("batch,fin,fout", [(64, 784, 128), (512, 784, 128), (96, 128, 10), (512, 128, 10),
                                           (32, 784, 10), (1024, 1024, 1024), (16, 3136, 10), (4, 128, 64)])
@pytest.mark.parametrize("relu", [0, 1])
def test_linear_fwd_bwd_vs_oracle(ctx, batch, fin, fout, relu):
    rng = np.random.default_rng(batch + fin + fout)
    R.Tape.reset()
    lin = R.Linear(fin, fout, True, rng)
    lin.bias._data[:] = rnd(rng, fout) * F32(0.1)
    x = rnd(rng, batch, fin)
    X = R.Tensor.new(x, x.shape).requires_grad_()
    _ = X.reshape(x.shape)                                 # occupy node 0
    Xr = X
    out = lin.forward(Xr)
    if relu:
        out = out.relu()
    gy = rnd(rng, batch, fout)
    out._grad[0] = gy.reshape(-1).copy()
    R.tape_backward(len(R.Tape.nodes) - 1)

    dX, dW, dB = ctx.upload(x), ctx.upload(lin.weight.data()), ctx.upload(lin.bias.data())
    Y = ctx.alloc(batch * fout)
    ctx.call("linear_fwd", dX, dW, dB, Y, batch, fin, fout, relu)
    y = Y.download()
    close(y, out.data())
    G = ctx.upload(gy)
    gx, gw, gb = ctx.alloc(batch * fin), ctx.alloc(fout * fin), ctx.alloc(fout)
    ctx.call("linear_bwd", dX, dW, G, Y if relu else None, gx, gw, gb, batch, fin, fout, 0, 0, 0)
    close(gx.download(), X.grad())
    close(gw.download(), lin.weight.grad())
    close(gb.download(), lin.bias.grad())
    # accumulate = 1 adds onto existing grads (src/ops.rs:250-253)
    ctx.call("linear_bwd", dX, dW, G, Y if relu else None, gx, gw, gb, batch, fin, fout, 1, 1, 1)
    close(gw.download(), 2 * lin.weight.grad())
    close(gb.download(), 2 * lin.bias.grad())
    close(gx.download(), 2 * X.grad())


# ------------------------------------------------------------------------------------------------
# softmax / cross-entropy / accuracy (src/loss.rs)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,cols", [(1, 2), (2, 3), (64, 10), (96, 10), (512, 10), (1024, 10), (100, 1000), (7, 33)])
def test_softmax_xent_vs_oracle(ctx, rows, cols):
    rng = np.random.default_rng(rows * 17 + cols)
    z = (rnd(rng, rows, cols) * F32(3)).astype(F32)
    t = rng.integers(0, cols, rows).astype(F32)
    R.Tape.reset()
    Z = R.Tensor.new(z, z.shape).requires_grad_()
    Tt = R.Tensor.new(t, t.shape)
    logp_ref = R.log_softmax(R.Tensor.new(z, z.shape)).data()
    loss = R.cross_entropy_loss(Z, Tt)
    loss.backward()

    dZ, dT = ctx.upload(z), ctx.upload(t)
    LP, L = ctx.alloc(rows * cols), ctx.alloc(1)
    ctx.call("softmax_xent_fwd", dZ, dT, LP, L, rows, cols)
    np.testing.assert_allclose(LP.download(), logp_ref, rtol=1e-5, atol=2e-6)
    close(L.download(), loss.data(), 1e-5)
    # fused head of a training step: the same loss / logp plus accuracy's correct count, one launch
    LP2, L2, Cn = ctx.alloc(rows * cols), ctx.alloc(1), ctx.alloc(1)
    for _ in range(2):                          # twice: the "last block" ticket must reset itself
        ctx.call("softmax_xent_acc_fwd", dZ, dT, LP2, L2, Cn, rows, cols)
        exact(LP2.download(), LP.download())
        close(L2.download(), loss.data(), 1e-5)
        assert Cn.download()[0] == round(float(R.accuracy(R.Tensor.new(z, z.shape), Tt)) * rows)
    ctx.call("softmax_xent_acc_fwd", dZ, dT, LP2, L2, None, rows, cols)
    close(L2.download(), loss.data(), 1e-5)
    one = ctx.upload(np.ones(1, F32))
    G = ctx.alloc(rows * cols)
    ctx.call("softmax_xent_bwd", LP, dT, one, G, rows, cols, 0)
    close(G.download(), Z.grad(), 1e-5)
    ctx.call("softmax_xent_bwd", LP, dT, one, G, rows, cols, 1)
    close(G.download(), 2 * Z.grad(), 1e-5)
    P = ctx.alloc(rows * cols)
    ctx.call("softmax_fwd", dZ, P, rows, cols)
    p = P.download().reshape(rows, cols)
    np.testing.assert_allclose(p.sum(axis=1), 1.0, atol=1e-5)
    np.testing.assert_allclose(p, R.softmax(R.Tensor.new(z, z.shape)).numpy(), rtol=1e-5, atol=1e-7)
    ctx.call("log_softmax_fwd", dZ, P, rows, cols)
    np.testing.assert_allclose(P.download(), logp_ref, rtol=1e-5, atol=2e-6)


def test_xent_kats(ctx):                                       # src/loss.rs:315-340; tests/smoke.rs:450-458
    for z, t, loss_ref, g_ref in (
            ([2.0, 1.0, -1.0, 3.0], [0, 1], 0.1657058, [-0.1344707, 0.1344707, 0.0089931, -0.0089931]),
            ([2, 1, 0, 0, 1, 2], [0, 2], 0.407606, [-0.1673795, 0.1223642, 0.0450153, 0.0450153, 0.1223642, -0.1673795])):
        rows = len(t); cols = len(z) // rows
        Z, Tt = ctx.upload(np.array(z, F32)), ctx.upload(np.array(t, F32))
        LP, L, G = ctx.alloc(len(z)), ctx.alloc(1), ctx.alloc(len(z))
        ctx.call("softmax_xent_fwd", Z, Tt, LP, L, rows, cols)
        ctx.call("softmax_xent_bwd", LP, Tt, ctx.upload(np.ones(1, F32)), G, rows, cols, 0)
        assert L.download()[0] == pytest.approx(loss_ref, abs=1e-6)
        np.testing.assert_allclose(G.download(), g_ref, atol=1e-6)


def test_softmax_stability_at_1000(ctx):                       # tests/smoke.rs:505-523
    Z = ctx.upload(np.array([1000, 1001, 1002], F32))
    P = ctx.alloc(3)
    ctx.call("softmax_fwd", Z, P, 1, 3)
    p = P.download()
    assert np.isfinite(p).all() and ((p >= 0) & (p <= 1)).all() and abs(p.sum() - 1) < 1e-6
    ctx.call("log_softmax_fwd", Z, P, 1, 3)
    assert np.isfinite(P.download()).all()


def test_xent_target_out_of_bounds_sets_device_error(ctx):     # reference asserts, src/loss.rs:161
    from taper_b200 import capi
    Z, Tt = ctx.upload(np.zeros(6, F32)), ctx.upload(np.array([0, 7], F32))
    LP, L = ctx.alloc(6), ctx.alloc(1)
    ctx.call("softmax_xent_fwd", Z, Tt, LP, L, 2, 3)
    flag = C.c_int()
    capi.check(capi.lib.tp_ctx_device_error(ctx.h, C.byref(flag)))
    assert flag.value == 1
    capi.check(capi.lib.tp_ctx_device_error(ctx.h, C.byref(flag)))
    assert flag.value == 0


@pytest.mark.parametrize("rows,cols", [(3, 2), (64, 10), (96, 10), (1000, 10), (16, 100)])
def test_accuracy_count_bit_exact(ctx, rows, cols):            # src/loss.rs:271-290
    rng = np.random.default_rng(rows + cols)
    p = rng.integers(0, 4, (rows, cols)).astype(F32)           # ties -> first max wins
    t = rng.integers(0, cols, rows).astype(F32)
    ref = float(R.accuracy(R.Tensor.new(p, p.shape), R.Tensor.new(t, t.shape))) * rows
    Cn = ctx.alloc(1)
    ctx.call("accuracy_count", ctx.upload(p), ctx.upload(t), Cn, rows, cols)
    assert Cn.download()[0] == round(ref)


def test_accuracy_kat(ctx):                                    # src/loss.rs:359-373 -> 2/3
    Cn = ctx.alloc(1)
    ctx.call("accuracy_count", ctx.upload(np.array([0.1, 0.9, 0.8, 0.2, 0.3, 0.7], F32)),
             ctx.upload(np.array([1, 0, 0], F32)), Cn, 3, 2)
    assert Cn.download()[0] == 2.0


# ------------------------------------------------------------------------------------------------
# conv2d / pooling (src/tensor.rs:1221-1285, 1391-1660, 1663-1780, 1972-2076)
# ------------------------------------------------------------------------------------------------
CONVS = [  # n, cin, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw
    (2, 1, 28, 28, 32, 3, 3, 1, 1, 1, 1, 1, 1),
    (3, 32, 14, 14, 64, 3, 3, 1, 1, 1, 1, 1, 1),
    (2, 64, 7, 7, 128, 3, 3, 1, 1, 1, 1, 1, 1),
    (2, 3, 9, 11, 5, 3, 3, 2, 2, 0, 1, 1, 1),
    (1, 2, 8, 8, 4, 1, 1, 1, 1, 0, 0, 1, 1),
    (2, 2, 10, 10, 3, 3, 2, 1, 2, 2, 0, 2, 1),
]


@pytest.mark.parametrize("cfg", CONVS)
def test_im2col_and_conv_fwd_vs_oracle(ctx, cfg):
    from taper_b200 import ConvDesc
    n, cin, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw = cfg
    rng = np.random.default_rng(sum(cfg))
    x, wt, b = rnd(rng, n, cin, h, w), rnd(rng, cout, cin, kh, kw), rnd(rng, cout)
    d = ConvDesc(*cfg)
    X = R.Tensor.new(x, x.shape)
    ho = (h + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    wo = (w + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    col_ref = X._im2col(kh, kw, (sh, sw), (ph, pw), (dh, dw), ho, wo).data()
    dx, dwt, db = ctx.upload(x), ctx.upload(wt), ctx.upload(b)
    col = ctx.alloc(col_ref.size)
    ctx.call("im2col", dx, col, d)
    exact(col.download(), col_ref)
    for relu in (0, 1):
        ref = (X.conv2d_relu if relu else X.conv2d)(R.Tensor.new(wt, wt.shape), R.Tensor.new(b, b.shape), (sh, sw), (ph, pw), (dh, dw))
        y = ctx.alloc(ref.data().size)
        ctx.call("conv2d_fwd", dx, dwt, db, y, d, relu)
        close(y.download(), ref.data())
    ref = X.conv2d(R.Tensor.new(wt, wt.shape), None, (sh, sw), (ph, pw), (dh, dw))
    y = ctx.alloc(ref.data().size)
    ctx.call("conv2d_fwd", dx, dwt, None, y, d, 0)
    close(y.download(), ref.data())


@pytest.mark.parametrize("cfg", CONVS)
@pytest.mark.parametrize("relu", [0, 1])
def test_conv_bwd_strict_and_full_adjoint_vs_oracle(ctx, cfg, relu):
    from taper_b200 import ConvDesc
    n, cin, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw = cfg
    rng = np.random.default_rng(sum(cfg) + 99)
    x, wt, b = rnd(rng, n, cin, h, w), rnd(rng, cout, cin, kh, kw), rnd(rng, cout)
    d = ConvDesc(*cfg)
    res = {}
    for strict in (True, False):
        R.Config.strict_reference_conv = strict
        R.Tape.reset()
        X = R.Tensor.new(x, x.shape).requires_grad_()
        _ = X.reshape(x.shape)
        W = R.Tensor.new(wt, wt.shape).requires_grad_()
        B = R.Tensor.new(b, b.shape).requires_grad_()
        out = (X.conv2d_relu if relu else X.conv2d)(W, B, (sh, sw), (ph, pw), (dh, dw))
        gy = np.random.default_rng(7).standard_normal(out.data().size).astype(F32)
        out._grad[0] = gy.copy()
        R.tape_backward(len(R.Tape.nodes) - 1)
        res[strict] = (out.data().copy(), X.grad(), W.grad(), B.grad(), gy)
    R.Config.strict_reference_conv = True
    y_ref, _, _, db_ref, gy = res[True]
    assert res[True][1] is None and res[True][2] is None             # A1: the reference drops both
    _, dx_ref, dw_ref, db_full, _ = res[False]
    dxb, dwb, dbb = ctx.upload(x), ctx.upload(wt), ctx.upload(b)
    Y = ctx.upload(y_ref)
    G = ctx.upload(gy)
    gb = ctx.alloc(cout)
    ctx.call("conv2d_bwd", dxb, dwb, G, Y if relu else None, None, None, gb, d, 0, 0, 0)      # strict_reference
    close(gb.download(), db_ref, 2e-5)
    gx, gw = ctx.alloc(x.size), ctx.alloc(wt.size)
    ctx.call("conv2d_bwd", dxb, dwb, G, Y if relu else None, gx, gw, gb, d, 0, 0, 0)          # full adjoint
    close(gb.download(), db_full, 2e-5)
    close(gw.download(), dw_ref)
    close(gx.download(), dx_ref)
    ctx.call("conv2d_bwd", dxb, dwb, G, Y if relu else None, gx, gw, gb, d, 1, 1, 1)
    close(gw.download(), 2 * dw_ref)
    close(gx.download(), 2 * dx_ref)
    close(gb.download(), 2 * db_full, 2e-5)


def test_col2im_is_adjoint_of_im2col(ctx):
    """<im2col(x), g> == <x, col2im(g)> — size-independent property."""
    from taper_b200 import ConvDesc
    cfg = (4, 8, 14, 14, 16, 3, 3, 1, 1, 1, 1, 1, 1)
    d = ConvDesc(*cfg)
    rng = np.random.default_rng(0)
    x = rnd(rng, 4, 8, 14, 14)
    g = rnd(rng, 4 * 14 * 14, 8 * 9)
    X, G = ctx.upload(x), ctx.upload(g)
    col, gx = ctx.alloc(g.size), ctx.alloc(x.size)
    ctx.call("im2col", X, col, d)
    ctx.call("col2im", G, gx, d, 0)
    lhs = np.dot(col.download().astype(np.float64), g.reshape(-1).astype(np.float64))
    rhs = np.dot(x.reshape(-1).astype(np.float64), gx.download().astype(np.float64))
    assert lhs == pytest.approx(rhs, rel=1e-6)


def test_bias4d_and_layout_permutes(ctx):
    rng = np.random.default_rng(3)
    n, c, h, w = 3, 5, 6, 7
    x, b, g = rnd(rng, n, c, h, w), rnd(rng, c), rnd(rng, n, c, h, w)
    X, B, G, Y = ctx.upload(x), ctx.upload(b), ctx.upload(g), ctx.alloc(x.size)
    ctx.call("add_bias_4d", X, B, Y, n, c, h * w, 0)
    exact(Y.download(), x + b[None, :, None, None])
    ctx.call("add_bias_4d", X, B, Y, n, c, h * w, 1)
    exact(Y.download(), np.maximum(x + b[None, :, None, None], 0))
    gb = ctx.alloc(c)
    ctx.call("bias_grad_4d", G, gb, n, c, h * w, 0)
    close(gb.download(), g.sum(axis=(0, 2, 3), dtype=np.float64), 1e-5)
    ctx.call("nchw_to_nhwc", X, Y, n, c, h, w)
    exact(Y.download(), x.transpose(0, 2, 3, 1).copy())
    Z = ctx.alloc(x.size)
    ctx.call("nhwc_to_nchw", Y, Z, n, h, w, c)
    exact(Z.download(), x)


POOLS = [  # n, c, h, w, kh, kw, sh, sw, ph, pw
    (2, 32, 28, 28, 2, 2, 2, 2, 0, 0),
    (3, 64, 14, 14, 2, 2, 2, 2, 0, 0),
    (2, 3, 9, 7, 3, 3, 2, 2, 1, 1),
    (2, 4, 8, 8, 3, 3, 1, 1, 1, 1),
    (2, 128, 7, 7, 7, 7, 7, 7, 0, 0),
]


@pytest.mark.parametrize("cfg", POOLS)
def test_maxpool_fwd_bwd_bit_exact(ctx, cfg):
    from taper_b200 import PoolDesc
    n, c, h, w, kh, kw, sh, sw, ph, pw = cfg
    rng = np.random.default_rng(sum(cfg))
    x = rng.integers(-4, 5, (n, c, h, w)).astype(F32)             # ties everywhere: first max must win
    R.Tape.reset()
    X = R.Tensor.new(x, x.shape).requires_grad_()
    _ = X.reshape(x.shape)
    out = X.max_pool2d((kh, kw), (sh, sw), (ph, pw))
    g = rnd(rng, out.data().size)
    X.set_grad(np.full(x.size, 7.0, F32))                         # must be overwritten (A6)
    out._grad[0] = g.copy()
    R.tape_backward(len(R.Tape.nodes) - 1)
    d = PoolDesc(*cfg)
    dX, Y, A = ctx.upload(x), ctx.alloc(out.data().size), ctx.alloc(out.data().size)
    ctx.call("maxpool2d_fwd", dX, Y, A, d)
    exact(Y.download(), out.data())
    gin = ctx.upload(np.full(x.size, 7.0, F32))
    ctx.call("maxpool2d_bwd", ctx.upload(g), A, gin, d)
    if kh <= sh and kw <= sw:
        exact(gin.download(), X.grad())                           # non-overlapping windows: one tap each, exact
    else:
        close(gin.download(), X.grad(), 1e-6)


@pytest.mark.parametrize("cfg", POOLS)
def test_avgpool_fwd_bwd(ctx, cfg):
    from taper_b200 import PoolDesc
    n, c, h, w, kh, kw, sh, sw, ph, pw = cfg
    rng = np.random.default_rng(sum(cfg) + 1)
    x = rnd(rng, n, c, h, w)
    R.Tape.reset()
    X = R.Tensor.new(x, x.shape).requires_grad_()
    _ = X.reshape(x.shape)
    out = X.avg_pool2d((kh, kw), (sh, sw), (ph, pw))
    g = rnd(rng, out.data().size)
    g0 = rnd(rng, x.size)
    X.set_grad(g0)
    out._grad[0] = g.copy()
    R.tape_backward(len(R.Tape.nodes) - 1)
    d = PoolDesc(*cfg)
    dX, Y = ctx.upload(x), ctx.alloc(out.data().size)
    ctx.call("avgpool2d_fwd", dX, Y, d)
    close(Y.download(), out.data(), 1e-5)
    gin = ctx.upload(g0)
    ctx.call("avgpool2d_bwd", ctx.upload(g), gin, d, 1)           # accumulates (A6)
    close(gin.download(), X.grad(), 1e-5)


# ------------------------------------------------------------------------------------------------
# optimizer steps (src/optim.rs:21-33, 83-113, 148-168)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 5, 1280, 101770, 1863690])
def test_sgd_adam_adamw_vs_oracle(ctx, n):
    from taper_b200 import capi
    rng = np.random.default_rng(n)
    p0 = rnd(rng, n)
    grads = [rnd(rng, n) * F32(0.1) for _ in range(3)]
    # SGD
    P = R.Tensor.new(p0.copy(), (n,)).requires_grad_()
    sgd = R.SGD([P], 0.01)
    dP = ctx.upload(p0)
    for g in grads:
        P.set_grad(g); sgd.step()
        ctx.call("sgd_step", dP, ctx.upload(g), 0.01, 1.0, n)
    exact(dP.download(), P.data())
    # Adam (wd = 0 and wd = 1e-4) and AdamW
    for kind, wd in (("adam", 0.0), ("adam", 1e-4), ("adamw", 1e-2)):
        P = R.Tensor.new(p0.copy(), (n,)).requires_grad_()
        opt = (R.Adam if kind == "adam" else R.AdamW)([P], 1e-3, None, None, wd)
        dP, M, V = ctx.upload(p0), ctx.zeros(n), ctx.zeros(n)
        for t, g in enumerate(grads, 1):
            P.set_grad(g); opt.step()
            ss = capi.lib.tp_adam_step_size(1e-3, 0.9, 0.999, t)
            if kind == "adam":
                ctx.call("adam_step", dP, ctx.upload(g), M, V, ss, 0.9, 0.999, 1e-8, wd, 1.0, n)
            else:
                decay = float(F32(1.0) - F32(1e-3) * F32(wd))
                ctx.call("adamw_step", dP, ctx.upload(g), M, V, ss, 0.9, 0.999, 1e-8, decay, 1.0, n)
        got, ref = dP.download(), P.data()
        np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-7)
        adam = opt if kind == "adam" else opt.adam
        np.testing.assert_allclose(M.download(), adam.m[0], rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(V.download(), adam.v[0], rtol=2e-6, atol=1e-12)


def test_adam_kat(ctx):                                        # src/optim.rs:360-389: grad 0.1, lr 1e-3, step 1
    from taper_b200 import capi
    P, G, M, V = ctx.upload(np.ones(4, F32)), ctx.upload(np.full(4, 0.1, F32)), ctx.zeros(4), ctx.zeros(4)
    ss = capi.lib.tp_adam_step_size(1e-3, 0.9, 0.999, 1)
    ctx.call("adam_step", P, G, M, V, ss, 0.9, 0.999, 1e-8, 0.0, 1.0, 4)
    d = 1.0 - P.download()
    assert (np.abs(d) > 1e-6).all()
    np.testing.assert_allclose(d, 0.00099999684, rtol=1e-4)


def test_grad_scale_folds_data_parallel_mean(ctx):
    rng = np.random.default_rng(0)
    p0, g = rnd(rng, 1000), rnd(rng, 1000)
    a, b = ctx.upload(p0), ctx.upload(p0)
    ctx.call("sgd_step", a, ctx.upload(g * F32(4)), 0.1, 0.25, 1000)     # summed grad of 4 ranks, scale 1/4
    ctx.call("sgd_step", b, ctx.upload(g), 0.1, 1.0, 1000)
    exact(a.download(), b.download())
