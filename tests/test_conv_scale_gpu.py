"""Convolution parity AT THE BASELINE SIZES (configs[2]: batch 256; configs[4]: batch 1024): every 3x3 layer of the shipped
5-conv example model (examples/train_mnist_cnn.rs:35-100) through tp_conv2d_fwd against the oracle's im2col + sgemm
(src/tensor.rs:1221-1285, 1728-1780), with the kernel that ran asserted — the implicit-GEMM tcgen05 instantiations
(128 x 32 / 128 x 64 tiles, two CTAs per SM; 128 x 128) are exercised at the tile counts bench.py times (1568 - 6272 tiles),
not only at toy batches — and one whole training step of the example model at batch 256.
Tolerance (north_star): 1e-4 of ||ref||_inf per tensor."""
import numpy as np
import pytest

from oracle import taper_ref as R
from test_step_gpu import close, make_pair, defaults  # noqa: F401

pytestmark = pytest.mark.gpu
F32 = np.float32

# (batch, C_in, H = W, C_out, expected path: 1 direct small-K kernel, 2 implicit GEMM on tcgen05 with LSU-gathered A tiles
# (3xTF32, gemm_tc.cu: tp_set_gemm_mode 1), 4 TMA-fed implicit GEMM on bf16 hi/lo planes (conv_bx3.cu: tp_set_gemm_mode 3 — and,
# whatever the mode, the fused conv stack of Sequential, tests/test_conv_stack_gpu.py))
LAYERS = [
    (256, 1, 28, 32, 1),          # conv1: K = 9, direct kernel on the CUDA cores
    (256, 32, 28, 32, 4),         # conv2: M = 200704, K = 288, N = 32  (1792 tiles of 4 x 32 pixels)
    (256, 32, 14, 64, 4),         # conv3: M = 50176,  K = 288, N = 64
    (256, 64, 14, 64, 4),         # conv4: M = 50176,  K = 576, N = 64
    (256, 64, 7, 128, 4),         # conv5: M = 12544,  K = 576, N = 128 (two images per tile, streamed weights)
    (1024, 32, 28, 32, 4),        # conv2 at configs[4]'s batch: 7168 tiles
    (256, 32, 28, 32, 2),         # the same layers on the 3xTF32 kernel: 128 x 32 tiles, two CTAs per SM
    (256, 64, 14, 64, 2),
    (256, 64, 7, 128, 2),         # 128 x 128 tiles
]


@pytest.fixture()
def ctx():
    import taper_b200
    c = taper_b200.Ctx(0)
    yield c
    c.close()


@pytest.mark.parametrize("n,cin,hw,cout,path", LAYERS)
def test_conv_relu_layer_at_baseline_batch_vs_oracle(ctx, n, cin, hw, cout, path):
    from taper_b200 import ConvDesc, capi
    rng = np.random.default_rng(n + cin + hw + cout)
    x = rng.random((n, cin, hw, hw)).astype(F32)                         # post-ReLU / pixel-like inputs: same-signed sums
    w = (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (cin * 9))).astype(F32)
    b = (rng.standard_normal(cout) * 0.05).astype(F32)
    d = ConvDesc(n, cin, hw, hw, cout, 3, 3, 1, 1, 1, 1, 1, 1)
    ref = R.Tensor.new(x, x.shape).conv2d_relu(R.Tensor.new(w, w.shape), R.Tensor.new(b, b.shape), (1, 1), (1, 1), (1, 1)).data()
    y = ctx.alloc(ref.size)
    ctx.call("set_gemm_mode", 3 if path == 4 else 1)          # eager conv: bf16x3 mode -> conv_bx3.cu, 3xTF32 mode -> gemm_tc.cu
    ctx.call("conv2d_fwd", ctx.upload(x), ctx.upload(w), ctx.upload(b), y, d, 1)
    assert capi.lib.tpdbg_last_conv_path() == path
    got = y.download()
    close(got, ref, 1e-4, f"conv_relu {n}x{cin}x{hw}x{hw} -> {cout}")
    # the ReLU zero pattern is part of the result: only units within summation noise of 0 may differ
    flips = np.sum((got > 0) != (ref.reshape(-1) > 0))
    assert flips <= 1e-5 * ref.size + 2, f"{flips} of {ref.size} ReLU decisions differ"


def test_example_cnn_training_step_at_batch_256():
    """One train_epoch iteration of the shipped example model at configs[2]'s batch (strict-reference conv autograd, SURVEY A1:
    conv5's bias and the three Linear layers train): loss, correct count and every parameter against the oracle."""
    from taper_b200 import host
    ref, m = make_pair(R.build_cnn5, host.CNN5, 3)
    tr = host.Trainer(m, "adam", lr=0.01, weight_decay=1e-4, eps=0.1)
    opt = R.Adam(ref.parameters(), 0.01, None, 0.1, 1e-4)
    rng = np.random.default_rng(7)
    for i in range(2):
        x = rng.random((256, 1, 28, 28)).astype(F32)
        y = rng.integers(0, 10, 256).astype(F32)
        lg = ref.forward(R.Tensor.new(x, x.shape)).numpy().astype(np.float64)
        top2 = np.sort(lg, axis=1)[:, -2:]
        near = int(np.sum(top2[:, 1] - top2[:, 0] <= 1e-4 * np.max(np.abs(lg))))
        loss_ref, acc_ref = R.train_step(ref, opt, R.Tensor.new(x, x.shape), R.Tensor.new(y, y.shape))
        loss, correct = tr.step(x, y)
        assert abs(loss - loss_ref) <= 1e-4 * abs(loss_ref), (i, loss, loss_ref)
        assert abs(correct - round(acc_ref * 256)) <= near
    for j, p in enumerate(ref.parameters()):
        close(m.get_param(j), p.data(), 1e-4, f"param {j}")
