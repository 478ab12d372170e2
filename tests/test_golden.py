"""Golden fixtures (tests/golden/): the reference's own known-answer vectors (reference_kats.json, transcribed from its tests)
and seeded oracle outputs (*.npz, written by tests/golden/make_golden.py).

CPU: the oracle reproduces both (pins the restatement and guards it against drift).
GPU: the CUDA path reproduces them through the C ABI — tape + CUDA-graph path and fused device step alike.
"""
import json
import os

import numpy as np
import pytest

from oracle import taper_ref as R

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(HERE, "reference_kats.json")))
F32 = np.float32
T = R.Tensor


def close(got, ref, tol=1e-4, what=""):
    got = np.asarray(got, np.float64).reshape(-1)
    ref = np.asarray(ref, np.float64).reshape(-1)
    assert got.shape == ref.shape, what
    assert np.max(np.abs(got - ref)) <= tol * max(np.max(np.abs(ref)), 1e-6), f"{what}: {np.max(np.abs(got - ref)):.3e}"


@pytest.fixture(autouse=True)
def fresh():
    R.Tape.reset()
    R.Config.strict_reference_conv = True
    R.Config.node0_sentinel = False
    yield
    R.Tape.reset()


# ---- CPU: oracle vs the reference's vectors -----------------------------------------------------------------------------
def test_oracle_matches_reference_kats():
    k = KATS["matmul"]
    a = T.new(k["a"], tuple(k["a_shape"])).requires_grad_()
    b = T.new(k["b"], tuple(k["b_shape"])).requires_grad_()
    c = a.matmul(b)
    np.testing.assert_allclose(c.data(), k["c"], atol=k["tol"])
    c.backward()
    np.testing.assert_allclose(a.grad(), k["derived_grad_a"], atol=k["tol"])
    np.testing.assert_allclose(b.grad(), k["derived_grad_b"], atol=k["tol"])
    k = KATS["sum"]
    x = T.new(k["x"], tuple(k["shape"]))
    assert x.sum().data()[0] == k["all"]
    np.testing.assert_array_equal(x.sum(0).data(), k["dim0"])
    np.testing.assert_array_equal(x.sum(1).data(), k["dim1"])
    assert x.sum(1, True).shape == tuple(k["keepdim_shape"])
    k = KATS["max_argmax"]
    x = T.new(k["x"], tuple(k["shape"]))
    v, i = x.max(0)
    np.testing.assert_array_equal(v.data(), k["max_dim0"])
    np.testing.assert_array_equal(i.data(), k["idx_dim0"])
    np.testing.assert_array_equal(x.argmax(1).data(), k["argmax_dim1"])
    k = KATS["exp"]
    np.testing.assert_allclose(T.new(k["x"], (3,)).exp().data(), k["y"], atol=k["tol"])
    for name in ("ce_2x2", "ce_2x3"):
        k = KATS[name]
        R.Tape.reset()
        z = T.new(k["logits"], tuple(k["shape"])).requires_grad_()
        loss = R.cross_entropy_loss(z, T.new(k["targets"], (len(k["targets"]),)))
        loss.backward()
        assert abs(loss.data()[0] - k["derived_loss"]) <= k["tol"]
        np.testing.assert_allclose(z.grad(), k["derived_grad"], atol=k["tol"])
    k = KATS["accuracy"]
    acc = R.accuracy(T.new(k["pred"], tuple(k["shape"])), T.new(k["targets"], (3,)))
    assert abs(acc - k["value"]) < 1e-6
    k = KATS["adam_step1"]
    p = T.new([1.0, -2.0, 0.5], (3,)).requires_grad_()
    before = p.data().copy()
    p.set_grad(np.full(3, k["grad"], F32))
    R.Adam([p], k["lr"]).step()
    np.testing.assert_allclose(before - p.data(), k["derived_delta"], rtol=2e-4)
    k = KATS["schedulers"]
    s = R.StepLR(k["step_lr"]["base"], k["step_lr"]["step"], k["step_lr"]["gamma"])
    for _ in range(3):
        s.step()
    assert abs(s.get_lr() - k["step_lr"]["after_3"]) < 1e-7
    e = R.ExponentialLR(k["exponential"]["base"], k["exponential"]["gamma"])
    e.step()
    assert abs(e.get_lr() - k["exponential"]["after_1"]) < 1e-7
    pl = R.ReduceLROnPlateau(k["plateau"]["base"], k["plateau"]["factor"], k["plateau"]["patience"])
    pl.step(1.0)
    for _ in range(2):
        pl.step(1.0)
    assert abs(pl.get_lr() - k["plateau"]["after_3_flat"]) < 1e-7


def _replay_oracle(fx):
    dims = [int(d) for d in fx["dims"]]
    kind = str(fx["kind"])
    model = R.build_mlp(dims, np.random.default_rng(0))
    for i, p in enumerate(model.parameters()):
        p._data[:] = fx[f"init_{i}"]
    params = model.parameters()
    opt = {"sgd": lambda: R.SGD(params, 0.05), "adam": lambda: R.Adam(params, 0.05, None, 0.1, 1e-3),
           "adamw": lambda: R.AdamW(params, 0.05, None, 0.1, 1e-2)}[kind]()
    losses = []
    for s in range(len(fx["loss"])):
        x, y = fx[f"x_{s}"], fx[f"y_{s}"]
        loss, _ = R.train_step(model, opt, T.new(x, x.shape), T.new(y, y.shape))
        losses.append(loss)
    return losses, [p.data() for p in model.parameters()]


@pytest.mark.parametrize("name", ["mlp_sgd", "mlp_adam", "mlp_adamw"])
def test_oracle_reproduces_golden_steps(name):
    fx = np.load(os.path.join(HERE, name + ".npz"))
    losses, params = _replay_oracle(fx)
    np.testing.assert_allclose(losses, fx["loss"], rtol=1e-6)
    for i, p in enumerate(params):
        close(p, fx[f"final_{i}"], 1e-6, f"{name} param {i}")


def test_oracle_reproduces_golden_conv_pool():
    fx = np.load(os.path.join(HERE, "conv_pool.npz"))
    X = T.new(fx["x"], fx["x"].shape)
    W = T.new(fx["w"], fx["w"].shape).requires_grad_()
    B = T.new(fx["b"], fx["b"].shape).requires_grad_()
    y = X.conv2d_relu(W, B, (1, 1), (1, 1), (1, 1))
    close(y.data(), fx["conv_relu"], 1e-6)
    mp = y.max_pool2d((2, 2), (2, 2))
    close(mp.data(), fx["maxpool"], 1e-6)
    gap = mp.avg_pool2d((3, 3), (3, 3))
    close(gap.data(), fx["gap"], 1e-6)
    gap.sum().backward()
    close(B.grad(), fx["grad_b"], 1e-6)
    assert W.grad() is None and bool(fx["grad_w_is_none"][0])


def test_oracle_reproduces_golden_conv_igemm():
    fx = np.load(os.path.join(HERE, "conv_igemm.npz"))
    X = T.new(fx["x"], fx["x"].shape)
    y = X.conv2d_relu(T.new(fx["w"], fx["w"].shape), T.new(fx["b"], fx["b"].shape), (1, 1), (1, 1), (1, 1))
    close(y.data(), fx["conv_relu"], 1e-6)
    close(y.max_pool2d((2, 2), (2, 2)).data(), fx["maxpool"], 1e-6)


# ---- GPU: the CUDA path vs the same fixtures, through the C ABI ----------------------------------------------------------
def _spec(dims):
    parts = []
    for i in range(len(dims) - 1):
        parts.append(f"linear:{dims[i]}:{dims[i + 1]}")
        if i < len(dims) - 2:
            parts.append("relu")
    return ",".join(parts)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name", ["mlp_sgd", "mlp_adam", "mlp_adamw"])
def test_cuda_reproduces_golden_steps(name, fused):
    from taper_b200 import host
    host.config(conv_full_adjoint=0, fuse_linear_relu=1, reference_op_sequence=0, gemm_mode=1)
    fx = np.load(os.path.join(HERE, name + ".npz"))
    dims = [int(d) for d in fx["dims"]]
    kind = str(fx["kind"])
    m = host.Model(_spec(dims), 0)
    for i in range(m.num_params()):
        m.set_param(i, fx[f"init_{i}"])
    hp = {"sgd": dict(lr=0.05), "adam": dict(lr=0.05, eps=0.1, weight_decay=1e-3), "adamw": dict(lr=0.05, eps=0.1, weight_decay=1e-2)}[kind]
    tr = host.Trainer(m, kind, **hp)
    tr.set_use_fused(fused)
    n = len(fx["loss"])
    for s in range(n):
        loss, correct = tr.step(fx[f"x_{s}"], fx[f"y_{s}"])
        assert abs(loss - fx["loss"][s]) <= 1e-4 * abs(fx["loss"][s]), (s, loss, fx["loss"][s])
        assert abs(correct - fx["correct"][s]) <= 1, (s, correct, fx["correct"][s])          # one near-tied row at most
    assert tr.fused_steps() == (n if fused else 0)
    for i in range(m.num_params()):
        close(m.get_param(i), fx[f"final_{i}"], 1e-4, f"{name} param {i}")


@pytest.mark.gpu
def test_cuda_reproduces_golden_conv_pool():
    from taper_b200 import host
    host.config(conv_full_adjoint=0, fuse_linear_relu=1, reference_op_sequence=0, gemm_mode=1)
    fx = np.load(os.path.join(HERE, "conv_pool.npz"))
    m = host.Model("conv_relu:3:4:3:1:1,maxpool:2:2,avgpool:3:3", 0)
    m.set_param(0, fx["w"])
    m.set_param(1, fx["b"])
    out = m.forward(fx["x"])
    close(out, fx["gap"], 1e-4, "conv_relu -> maxpool -> avgpool")
    m2 = host.Model("conv_relu:3:4:3:1:1", 0)
    m2.set_param(0, fx["w"])
    m2.set_param(1, fx["b"])
    close(m2.forward(fx["x"]), fx["conv_relu"], 1e-4, "conv_relu")


@pytest.mark.gpu
@pytest.mark.parametrize("path", [4, 2])
def test_cuda_reproduces_golden_conv_on_the_tcgen05_path(path):
    """C_out = 32, K = 288: this fixture goes through the implicit-GEMM tensor-core kernels (asserted: 4 = the TMA-fed bf16x3
    kernel of conv_bx3.cu, 2 = round 1's 3xTF32 kernel with gathered A tiles), unlike conv_pool.npz whose C_out = 4 layer takes
    the direct kernel."""
    from taper_b200 import host, capi
    host.config(conv_full_adjoint=0, fuse_linear_relu=1, reference_op_sequence=0, gemm_mode=3 if path == 4 else 1)
    try:
        fx = np.load(os.path.join(HERE, "conv_igemm.npz"))
        m = host.Model("conv_relu:32:32:3:1:1", 0)
        m.set_param(0, fx["w"])
        m.set_param(1, fx["b"])
        close(m.forward(fx["x"]), fx["conv_relu"], 1e-4, "conv_relu 32 -> 32")
        assert capi.lib.tpdbg_last_conv_path() == path
        m2 = host.Model("conv_relu:32:32:3:1:1,maxpool:2:2", 0)
        m2.set_param(0, fx["w"])
        m2.set_param(1, fx["b"])
        close(m2.forward(fx["x"]), fx["maxpool"], 1e-4, "conv_relu -> maxpool")
    finally:
        host.config(gemm_mode=1)
