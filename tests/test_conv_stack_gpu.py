"""tp_conv_stack_fwd (conv_bx3.cu): chains of 3x3 / s1 / p1 Conv + bias (+ ReLU) (+ 2x2 max-pool) on the TMA-fed tcgen05 bf16x3
kernel over NHWC bf16 hi/lo planes, against the oracle's layer-by-layer conv2d_relu / max_pool2d (src/tensor.rs:1221-1285,
1379-1389, 1391-1464).  Covers both tap-addressing variants of the kernel (three atom-aligned patches; one patch read through
row-shifted descriptors), every tile geometry the MNIST models produce (28x28: 4 rows x 32 columns per tile, 14x14: 8 x 16,
7x7: two images per tile), ragged batches against the images-per-tile packing, resident and streamed weights, the planes
-> planes and planes -> NCHW epilogues with and without pooling, and the Sequential peephole that feeds it.
Tolerance: 1e-4 of ||ref||_inf (north_star); measured ~5e-6 per layer, ~1.5e-5 after five layers."""
import ctypes as C

import numpy as np
import pytest

from oracle import taper_ref as R

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.fixture()
def ctx():
    import taper_b200
    from taper_b200 import capi
    c = taper_b200.Ctx(0)
    yield c
    capi.lib.tpdbg_conv_shift_mode(-1)
    c.close()


def oracle_stack(x, ws, bs, pools, relus):
    t = R.Tensor.new(x, x.shape)
    for w, b, pool, relu in zip(ws, bs, pools, relus):
        wt = R.Tensor.new(w, w.shape)
        bt = R.Tensor.new(b, b.shape) if b is not None else None
        t = t.conv2d_relu(wt, bt, (1, 1), (1, 1), (1, 1)) if relu else t.conv2d(wt, bt, (1, 1), (1, 1), (1, 1))
        if pool:
            t = t.max_pool2d((2, 2), (2, 2))
    return t.data().reshape(t.shape)


def run_stack(ctx, x, ws, bs, pools, relus, expect_rc=0):
    from taper_b200 import capi
    lib = capi.lib
    n, c, h, w = x.shape
    L = len(ws)
    xb = ctx.upload(x)
    wb = [ctx.upload(v) for v in ws]
    bb = [ctx.upload(v) if v is not None else None for v in bs]
    hh, ww = h, w
    for p in pools:
        if p:
            hh //= 2
            ww //= 2
    cout = [v.shape[0] for v in ws]
    y = ctx.alloc(n * cout[-1] * hh * ww)
    capi.check(lib.tp_buf_fill(ctx.h, y.h, -7.0, y.n))
    W = (C.c_void_p * L)(*[b.h for b in wb])
    B = (C.c_void_p * L)(*[(b.h if b is not None else None) for b in bb])
    co = (C.c_int * L)(*cout)
    po = (C.c_int * L)(*[int(p) for p in pools])
    re = (C.c_int * L)(*[int(r) for r in relus])
    rc = lib.tp_conv_stack_fwd(ctx.h, xb.h, n, c, h, w, L, W, B, co, po, re, y.h)
    if expect_rc:
        assert rc == expect_rc, rc
        return None
    capi.check(rc)
    return y.download().reshape(n, cout[-1], hh, ww)


def make(rng, n, c0, hw, couts, bias=True):
    x = rng.random((n, c0, hw, hw)).astype(F32)
    ws, bs = [], []
    ci = c0
    for co in couts:
        ws.append((rng.standard_normal((co, ci, 3, 3)) * np.sqrt(2.0 / (ci * 9))).astype(F32))
        bs.append((rng.standard_normal(co) * 0.05).astype(F32) if bias else None)
        ci = co
    return x, ws, bs


def close(got, ref, tol, what):
    scale = max(float(np.abs(ref).max()), 1e-6)
    err = float(np.abs(got - ref).max())
    assert not np.isnan(got).any(), what
    assert err <= tol * scale, f"{what}: max |diff| {err:.3e} > {tol:g} * {scale:.3e}"


CASES = [
    # name, n, c0, hw, couts, pools
    ("32->32 @28", 3, 32, 28, [32], [0]),
    ("32->32 @28 pool", 3, 32, 28, [32], [1]),
    ("32->64 @14", 5, 32, 14, [64], [0]),
    ("64->64 @14 pool", 5, 64, 14, [64], [1]),
    ("64->128 @7, odd batch on two images per tile", 5, 64, 7, [128], [0]),
    ("64->128 @7, one image", 1, 64, 7, [128], [0]),
    ("64->32 @20 (ragged last row block)", 2, 64, 20, [32], [0]),
    ("32->64 @5 pool (odd size, floor pooling)", 7, 32, 5, [64], [1]),
    ("128->128 @7 (streamed weights)", 3, 128, 7, [128], [0]),
    ("128->64 @14 pool (streamed weights)", 2, 128, 14, [64], [1]),
    ("96->32 @9 (three channel blocks)", 4, 96, 9, [32], [0]),
    ("example CNN stack", 5, 1, 28, [32, 32, 64, 64, 128], [0, 1, 0, 1, 0]),
    ("contract CNN stack", 6, 1, 28, [32, 64], [1, 1]),
    ("3-channel image stack", 3, 3, 16, [32, 32], [0, 1]),
    ("32-channel input stack", 3, 32, 12, [64, 64, 32], [1, 0, 0]),
    # several tiles per persistent CTA: both MMA issuers at work, odd tile counts, the weight ring shared by the tile pair
    ("128->128 @7 batch 700 (350 tiles, streamed weights)", 700, 128, 7, [128], [0]),
    ("64->64 @14 batch 301 pool (602 tiles, streamed weights)", 301, 64, 14, [64], [1]),
    ("32->32 @28 batch 75 (525 tiles, resident weights)", 75, 32, 28, [32], [0]),
    ("96->64 @14 batch 223 (446 tiles, three channel blocks)", 223, 96, 14, [64], [0]),
]


@pytest.mark.parametrize("mode", [2, 0])
@pytest.mark.parametrize("name,n,c0,hw,couts,pools", CASES, ids=[c[0] for c in CASES])
def test_conv_stack_vs_oracle(ctx, mode, name, n, c0, hw, couts, pools):
    from taper_b200 import capi
    capi.lib.tpdbg_conv_shift_mode(mode)
    rng = np.random.default_rng(abs(hash((n, c0, hw, tuple(couts)))) % (2 ** 31))
    x, ws, bs = make(rng, n, c0, hw, couts)
    relus = [1] * len(couts)
    ref = oracle_stack(x, ws, bs, pools, relus)
    got = run_stack(ctx, x, ws, bs, pools, relus)
    assert not (got == -7.0).any(), "output elements left unwritten"
    close(got, ref, 1e-4, name)
    flips = int(np.sum((got > 0) != (ref > 0)))
    assert flips <= 1e-4 * ref.size + 2, f"{flips} of {ref.size} ReLU decisions differ"


def test_conv_stack_without_bias_and_relu(ctx):
    rng = np.random.default_rng(5)
    x, ws, _ = make(rng, 3, 32, 14, [64, 32])
    x = (x - 0.5).astype(F32)                                              # signed inputs: cancellation in the sums
    ref = oracle_stack(x, ws, [None, None], [0, 0], [0, 0])
    got = run_stack(ctx, x, ws, [None, None], [0, 0], [0, 0])
    close(got, ref, 1e-4, "no bias, no relu")
    assert (got < 0).any()


TP_ERR_UNSUPPORTED = 5            # include/taper_b200.h


def test_conv_stack_rejects_shapes_outside_the_kernel(ctx):
    """W >= 32, C_in not a multiple of 32, C_out outside {32, 64, 128}, a lone small-K layer: status 5 and nothing launched (the
    host layer then runs the layers one by one)."""
    rng = np.random.default_rng(6)
    for (c0, hw, couts) in [(32, 32, [32]), (16, 14, [32]), (32, 14, [48]), (1, 28, [32])]:
        x, ws, bs = make(rng, 2, c0, hw, couts)
        l0 = ctx.launches()
        run_stack(ctx, x, ws, bs, [0] * len(couts), [1] * len(couts), expect_rc=TP_ERR_UNSUPPORTED)
        assert ctx.launches() - l0 <= 1            # the fill of y in run_stack


def test_example_cnn_stack_at_baseline_batch_vs_oracle(ctx):
    """The five conv layers of the shipped example model (examples/train_mnist_cnn.rs:35-100) at configs[2]'s batch 256 through
    one tp_conv_stack_fwd: 1792 + 512 + 512 + 128 tensor-core tiles."""
    rng = np.random.default_rng(11)
    x, ws, bs = make(rng, 256, 1, 28, [32, 32, 64, 64, 128])
    pools = [0, 1, 0, 1, 0]
    ref = oracle_stack(x, ws, bs, pools, [1] * 5)
    got = run_stack(ctx, x, ws, bs, pools, [1] * 5)
    close(got, ref, 1e-4, "example CNN stack, batch 256")


def test_conv2_layer_at_config4_batch_vs_oracle(ctx):
    """The 32 -> 32 layer at 28x28 with its pool at configs[4]'s batch 1024: 7168 tiles over the persistent CTAs."""
    rng = np.random.default_rng(12)
    x, ws, bs = make(rng, 1024, 32, 28, [32])
    ref = oracle_stack(x, ws, bs, [1], [1])
    got = run_stack(ctx, x, ws, bs, [1], [1])
    close(got, ref, 1e-4, "32 -> 32 @28 + pool, batch 1024")


def test_sequential_peephole_runs_the_stack_and_matches_layer_by_layer():
    """Sequential::forward with the conv-stack peephole on / off: same output within the bf16x3 bound, far fewer launches; the
    backward through the stack's one tape node delivers the last conv's bias gradient exactly like the per-layer tape."""
    from taper_b200 import host
    host.config(conv_full_adjoint=0, fuse_linear_relu=1, reference_op_sequence=0, gemm_mode=3)     # per-layer convs on the same kernel
    rng = np.random.default_rng(3)
    x = rng.random((6, 1, 28, 28)).astype(F32)
    outs, launches = [], []
    for fuse in (1, 0):
        host.config_conv_stack(fuse)
        m = host.Model(host.CNN5, 7)
        m.forward(x)                                                        # sizes the allocator caches
        l0 = host.launches()
        outs.append(m.forward(x))
        launches.append(host.launches() - l0)
    host.config_conv_stack(1)
    host.config(gemm_mode=1)
    close(outs[0], outs[1], 1e-4, "fused vs per-layer logits")
    assert launches[0] + 6 <= launches[1], launches


def test_gap_count_and_bias_gradient_vs_oracle(ctx):
    """tp_gap_count_fwd / tp_gap_relu_bias_grad against the oracle's chain relu -> avg_pool2d (global) and its backward into the
    conv bias (src/tensor.rs:1524-1660, src/ops.rs:358-370, src/tensor.rs:2003-2027): mean exact to summation order, count exact,
    bias gradient 1e-5."""
    rng = np.random.default_rng(21)
    for (n, c, hw) in [(1024, 128, 49), (37, 64, 196), (5, 32, 9)]:
        side = int(round(hw ** 0.5))
        z = (rng.standard_normal((n, c, side, side))).astype(F32)
        y = np.maximum(z, 0).astype(F32)
        g = (rng.standard_normal((n, c)) * 1e-3).astype(F32)
        mean, cnt, gb = ctx.alloc(n * c), ctx.alloc(n * c), ctx.alloc(c)
        ctx.call("gap_count_fwd", ctx.upload(y), mean, cnt, n, c, hw)
        np.testing.assert_allclose(mean.download().reshape(n, c), y.reshape(n, c, hw).mean(axis=2), rtol=2e-6, atol=1e-7)
        np.testing.assert_array_equal(cnt.download().reshape(n, c), (y.reshape(n, c, hw) > 0).sum(axis=2).astype(F32))
        # oracle: loss = sum(avg_pool2d(relu(z)) * G)  =>  z.grad summed over n, h, w is what add_bias_4d hands to the bias
        R.Tape.reset()
        zt = R.Tensor.new(z, z.shape).requires_grad_()
        pooled = zt.relu().avg_pool2d((side, side), (side, side))
        (pooled * R.Tensor.new(g.reshape(n, c, 1, 1), (n, c, 1, 1))).sum().backward()
        ref = np.asarray(zt.grad(), np.float64).reshape(n, c, hw).sum(axis=(0, 2))
        R.Tape.reset()
        ctx.call("gap_relu_bias_grad", ctx.upload(g), cnt, gb, n, c, hw, 0)
        got = gb.download()
        assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-6)
        ctx.call("gap_relu_bias_grad", ctx.upload(g), cnt, gb, n, c, hw, 1)          # accumulate
        assert np.abs(gb.download() - 2 * ref).max() <= 2e-5 * max(np.abs(ref).max(), 1e-6)
        ctx.call("gap_relu_bias_grad", ctx.upload(g), None, gb, n, c, hw, 0)         # no ReLU: every unit passes
        assert np.abs(gb.download() - g.astype(np.float64).sum(axis=0)).max() <= 1e-5 * np.abs(g.sum(axis=0)).max() + 1e-9


def test_example_cnn_step_fused_vs_per_layer_tape():
    """loss_backward of the shipped example model with the conv-stack (+ global average pool) node against the per-layer tape of
    the same library: same loss, same None pattern, the same bias / Linear gradients to 1e-4 (+ the ReLU-threshold slack is not
    needed here: in bf16x3 mode both paths share the conv kernel, so the masks agree)."""
    from taper_b200 import host
    host.config(conv_full_adjoint=0, fuse_linear_relu=1, reference_op_sequence=0, gemm_mode=3)
    rng = np.random.default_rng(4)
    x = rng.random((32, 1, 28, 28)).astype(F32)
    y = rng.integers(0, 10, 32).astype(F32)
    res = []
    for fuse in (1, 0):
        host.config_conv_stack(fuse)
        m = host.Model(host.CNN5, 9)
        m.zero_grad()
        loss, correct, _ = m.loss_backward(x, y)
        res.append((loss, correct, [m.get_grad(j) for j in range(m.num_params())]))
    host.config_conv_stack(1)
    host.config(gemm_mode=1)
    assert res[0][0] == pytest.approx(res[1][0], rel=1e-5)
    assert res[0][1] == res[1][1]
    for j, (ga, gb) in enumerate(zip(res[0][2], res[1][2])):
        assert (ga is None) == (gb is None), j
        if ga is not None:
            close(ga, gb, 1e-4, f"grad {j}")


@pytest.mark.parametrize("n,cin,hw,cout,relu", [
    (5, 32, 28, 32, 1),          # conv2 of the example model
    (6, 32, 14, 64, 1),          # conv3: gradient planes of 64 channels, 32 output channels
    (7, 64, 14, 64, 0),          # conv4, no ReLU
    (9, 64, 7, 128, 1),          # conv5: two images per tile, 128-channel gradient (four channel blocks)
    (300, 64, 14, 64, 1),        # several tiles per persistent CTA; dW accumulates over image chunks
])
@pytest.mark.parametrize("mode", [1, 3])
def test_full_adjoint_input_gradient_on_the_conv_kernel_vs_oracle(ctx, mode, n, cin, hw, cout, relu):
    """tp_conv2d_bwd with dx / dw requested (full adjoint, the two tape links the reference drops): dX comes from the
    implicit-GEMM kernel run on the masked gradient with mirrored, transposed weights (no [M, 9*C_in] matrix, no col2im); dW
    from the implicit GEMM over pixels (conv_dw_kernel: MN-major planes, two taps per MMA) in both tensor-core modes.  Against the
    oracle's matmul backward through im2col (src/ops.rs:254-291)."""
    from taper_b200 import ConvDesc
    rng = np.random.default_rng(n * 131 + cin + cout + hw)
    x = (rng.random((n, cin, hw, hw)) - 0.3).astype(F32)
    wt = (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (cin * 9))).astype(F32)
    b = (rng.standard_normal(cout) * 0.05).astype(F32)
    R.Config.strict_reference_conv = False
    try:
        R.Tape.reset()
        X = R.Tensor.new(x, x.shape).requires_grad_()
        _ = X.reshape(x.shape)
        W = R.Tensor.new(wt, wt.shape).requires_grad_()
        B = R.Tensor.new(b, b.shape).requires_grad_()
        out = (X.conv2d_relu if relu else X.conv2d)(W, B, (1, 1), (1, 1), (1, 1))
        gy = np.random.default_rng(7).standard_normal(out.data().size).astype(F32)
        out._grad[0] = gy.copy()
        R.tape_backward(len(R.Tape.nodes) - 1)
        y_ref, dx_ref, dw_ref, db_ref = out.data().copy(), X.grad().copy(), W.grad().copy(), B.grad().copy()
    finally:
        R.Config.strict_reference_conv = True
        R.Tape.reset()
    d = ConvDesc(n, cin, hw, hw, cout, 3, 3, 1, 1, 1, 1, 1, 1)
    xb, wb, Y, G = ctx.upload(x), ctx.upload(wt), ctx.upload(y_ref), ctx.upload(gy)
    gx, gw, gb = ctx.alloc(x.size), ctx.alloc(wt.size), ctx.alloc(cout)
    mem0 = ctx.launches()
    ctx.call("set_gemm_mode", mode)
    ctx.call("conv2d_bwd", xb, wb, G, Y if relu else None, gx, gw, gb, d, 0, 0, 0)
    close(gx.download(), dx_ref.reshape(-1), 1e-4, "dX")
    close(gw.download(), dw_ref.reshape(-1), 1e-4, "dW")
    close(gb.download(), db_ref.reshape(-1), 1e-4, "db")
    ctx.call("conv2d_bwd", xb, wb, G, Y if relu else None, gx, gw, gb, d, 1, 1, 1)          # accumulate into live gradients
    close(gx.download(), 2 * dx_ref.reshape(-1), 1e-4, "dX accumulated")
    close(gw.download(), 2 * dw_ref.reshape(-1), 1e-4, "dW accumulated")
    assert ctx.launches() > mem0


@pytest.mark.parametrize("n,c0,hw,couts,pools", [
    (5, 1, 28, [32, 32, 64, 64, 128], [0, 1, 0, 1, 0]),       # the example model: 7x7 last layer, pooled in the conv epilogue
    (301, 64, 7, [128], [0]),                                   # several tiles per CTA, odd batch on two images per tile
    (3, 32, 5, [64, 32], [0, 0]),                               # three 5x5 images per tile
    (4, 32, 14, [64], [0]),                                     # 14x14: images span two tiles -> scratch NCHW + gap kernel
    (6, 1, 28, [32, 64], [1, 1]),                               # pooled last layer -> scratch NCHW + gap kernel
])
def test_conv_stack_with_global_average_pool_vs_oracle(ctx, n, c0, hw, couts, pools):
    """tp_conv_stack_gap_fwd: mean and positive-unit count of the last layer's planes against the oracle's conv chain followed by
    avg_pool2d over the whole plane (src/nn.rs:670-686)."""
    from taper_b200 import capi
    lib = capi.lib
    rng = np.random.default_rng(n + c0 + hw + len(couts))
    x, ws, bs = make(rng, n, c0, hw, couts)
    relus = [1] * len(couts)
    ref = oracle_stack(x, ws, bs, pools, relus)                               # [n, C, h, w]
    L = len(ws)
    xb = ctx.upload(x)
    wb = [ctx.upload(v) for v in ws]
    bb = [ctx.upload(v) for v in bs]
    mean, cnt = ctx.alloc(n * couts[-1]), ctx.alloc(n * couts[-1])
    W = (C.c_void_p * L)(*[b.h for b in wb])
    B = (C.c_void_p * L)(*[b.h for b in bb])
    co = (C.c_int * L)(*couts)
    po = (C.c_int * L)(*pools)
    re = (C.c_int * L)(*relus)
    capi.check(lib.tp_conv_stack_gap_fwd(ctx.h, xb.h, n, c0, hw, hw, L, W, B, co, po, re, mean.h, cnt.h))
    flat = ref.reshape(n, couts[-1], -1)
    close(mean.download().reshape(n, -1), flat.mean(axis=2), 1e-4, "pooled features")
    got_cnt = cnt.download().reshape(n, -1)
    ref_cnt = (flat > 0).sum(axis=2)
    # a unit within the conv error of zero may land on either side of the ReLU
    near = (np.abs(flat) <= 4e-5 * np.abs(flat).max()).sum(axis=2)
    assert (np.abs(got_cnt - ref_cnt) <= near).all()
