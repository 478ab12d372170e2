"""Pins the CPU oracle (oracle/taper_ref.py) against every known-answer test the reference's own
tests hold for the hot path.  Citations: path:line in vaibhawvipul/taper."""
import math

import numpy as np
import pytest

from oracle import taper_ref as R

T = R.Tensor


@pytest.fixture(autouse=True)
def fresh_tape():
    R.Tape.reset()
    R.Config.strict_reference_conv = True
    R.Config.node0_sentinel = False
    yield
    R.Tape.reset()


def test_mul_grads():                       # tests/smoke.rs:20-30 (intent; A3 fixed by id+1 stamping)
    x = T.scalar(2.0).requires_grad_()
    y = T.scalar(3.0).requires_grad_()
    z = x * y
    z.backward()
    assert z.data()[0] == 6.0
    assert x.grad()[0] == 3.0 and y.grad()[0] == 2.0


def test_node0_sentinel_reproduces_reference_noop():   # src/tensor.rs:524-528 + src/tape.rs:65 (A3)
    R.Config.node0_sentinel = True
    x = T.scalar(2.0).requires_grad_()
    y = T.scalar(3.0).requires_grad_()
    z = x * y
    z.backward()
    assert x.grad() is None and y.grad() is None


def test_mul_add_chain():                   # tests/smoke.rs:33-43
    a = T.scalar(2.0).requires_grad_()
    b = T.scalar(3.0).requires_grad_()
    c = a * b + a
    c.backward()
    assert c.data()[0] == 8.0
    assert a.grad()[0] == 4.0 and b.grad()[0] == 2.0


def test_matmul_shapes_and_grads():         # tests/smoke.rs:46-70
    a = T.new([1, 2, 3, 4, 5, 6], (2, 3)).requires_grad_()
    b = T.new([7, 8, 9, 10, 11, 12], (3, 2)).requires_grad_()
    c = a.matmul(b)
    assert c.shape == (2, 2)
    np.testing.assert_allclose(c.data(), [58, 64, 139, 154], atol=1e-4)
    c.backward()
    assert a.grad().shape == (6,) and b.grad().shape == (6,)
    np.testing.assert_allclose(a.grad(), [15, 19, 23, 15, 19, 23], atol=1e-4)
    np.testing.assert_allclose(b.grad(), [5, 5, 7, 7, 9, 9], atol=1e-4)


def test_sgemm_rowmajor_ld_rules():         # src/gemm.rs:21-29, 88-98
    rng = np.random.default_rng(0)
    m, n, k = 5, 7, 3
    A = rng.standard_normal((m, k)).astype(np.float32)
    B = rng.standard_normal((k, n)).astype(np.float32)
    C0 = rng.standard_normal((m, n)).astype(np.float32)
    for ta in (False, True):
        for tb in (False, True):
            a = (A.T if ta else A).copy().reshape(-1)
            b = (B.T if tb else B).copy().reshape(-1)
            c = C0.copy().reshape(-1)
            R.sgemm_rowmajor(ta, tb, m, n, k, 1.0, a, b, 1.0, c)
            np.testing.assert_allclose(c.reshape(m, n), A @ B + C0, rtol=1e-5, atol=1e-5)


def test_reshape_shapes_and_grad():         # tests/smoke.rs:262-307
    x = T.new(np.arange(6), (2, 3)).requires_grad_()
    r = x.reshape((3, 2))
    assert r.shape == (3, 2)
    assert x.flatten(0).shape == (6,)
    s = r.sum()
    s.backward()
    np.testing.assert_array_equal(x.grad(), np.ones(6, np.float32))


def test_sum_dims():                        # tests/smoke.rs:310-336
    x = T.new([1, 2, 3, 4, 5, 6], (2, 3))
    assert x.sum().data()[0] == 21.0
    np.testing.assert_array_equal(x.sum(0).data(), [5, 7, 9])
    np.testing.assert_array_equal(x.sum(1).data(), [6, 15])
    assert x.sum(1, True).shape == (2, 1)


def test_max_argmax():                      # tests/smoke.rs:357-377
    x = T.new([1, 3, 2, 4, 6, 5], (2, 3))
    v, i = x.max(0)
    assert v.shape == (1, 3)
    np.testing.assert_array_equal(v.data(), [4, 6, 5])
    np.testing.assert_array_equal(i.data(), [1, 1, 1])
    am = x.argmax(1)
    assert am.shape == (2, 1)
    np.testing.assert_array_equal(am.data(), [1, 1])


def test_argmax_first_max_wins_and_nan():   # src/tensor.rs:1062 strict '>' (A7)
    x = T.new([2, 2, 1, np.nan, 0, 0], (2, 3))
    np.testing.assert_array_equal(x.argmax(1).data(), [0, 1])


def test_exp_log_values():                  # tests/smoke.rs:380-406
    x = T.new([0.0, 1.0, 2.0], (3,))
    np.testing.assert_allclose(x.exp().data(), [1.0, 2.71828, 7.38906], atol=1e-4)
    np.testing.assert_allclose(x.exp().log().data(), x.data(), atol=1e-5)


def test_exp_log_grads():                   # tests/smoke.rs:409-435
    x = T.new([0.5, 1.0, 2.0], (3,)).requires_grad_()
    _pad = x.reshape((3,))                  # occupy node 0 the way a real graph would
    y = x.exp().sum()
    y.backward()
    np.testing.assert_allclose(x.grad(), np.exp(x.data()), atol=1e-5)
    R.Tape.reset()
    x = T.new([0.5, 1.0, 2.0], (3,)).requires_grad_()
    y = x.log().sum()
    y.backward()
    np.testing.assert_allclose(x.grad(), 1.0 / x.data(), atol=1e-5)


def test_softmax_rows_sum_to_one_and_stable():   # tests/smoke.rs:438-447, 505-523; src/loss.rs:298-312 (intent, A13)
    x = T.new([1, 2, 3, 1, 2, 3], (2, 3))
    p = R.softmax(x).numpy()
    np.testing.assert_allclose(p.sum(axis=1), 1.0, atol=1e-6)
    assert (p > 0).all()
    big = T.new([1000, 1001, 1002], (1, 3))
    pb = R.softmax(big).data()
    assert np.isfinite(pb).all() and ((pb >= 0) & (pb <= 1)).all()
    assert np.isfinite(R.log_softmax(big).data()).all()


def test_cross_entropy_gradient_signs_and_values():   # src/loss.rs:315-340
    logits = T.new([2.0, 1.0, -1.0, 3.0], (2, 2)).requires_grad_()
    targets = T.new([0.0, 1.0], (2,))
    loss = R.cross_entropy_loss(logits, targets)
    loss.backward()
    g = logits.grad()
    assert g[0] < 0 and g[3] < 0
    assert loss.data()[0] == pytest.approx(0.1657058, abs=1e-6)
    np.testing.assert_allclose(g, [-0.1344707, 0.1344707, 0.0089931, -0.0089931], atol=1e-6)


def test_cross_entropy_smoke():             # tests/smoke.rs:450-458
    logits = T.new([2, 1, 0, 0, 1, 2], (2, 3)).requires_grad_()
    loss = R.cross_entropy_loss(logits, T.new([0, 2], (2,)))
    assert loss.data()[0] > 0
    assert loss.data()[0] == pytest.approx(0.407606, abs=1e-6)
    loss.backward()
    np.testing.assert_allclose(logits.grad(), [-0.1673795, 0.1223642, 0.0450153, 0.0450153, 0.1223642, -0.1673795], atol=1e-6)


def test_cross_entropy_log_softmax_nodes_are_dead():   # src/loss.rs:174-191 (A5): 5 log_softmax nodes + CE node
    logits = T.new([2, 1, 0, 0, 1, 2], (2, 3)).requires_grad_()
    R.cross_entropy_loss(logits, T.new([0, 2], (2,)))
    assert len(R.Tape.nodes) == 6


def test_accuracy():                        # src/loss.rs:359-373
    pred = T.new([0.1, 0.9, 0.8, 0.2, 0.3, 0.7], (3, 2))
    assert float(R.accuracy(pred, T.new([1, 0, 0], (3,)))) == pytest.approx(2 / 3, abs=1e-6)


def test_adam_step_changes_all_params():    # src/optim.rs:360-389
    p = T.new(np.ones(4), (4,)).requires_grad_()
    opt = R.Adam([p], 1e-3)
    p.set_grad(np.full(4, 0.1, np.float32))
    before = p.data().copy()
    opt.step()
    d = np.abs(p.data() - before)
    assert (d > 1e-6).all()
    np.testing.assert_allclose(d, 0.00099999684, rtol=1e-4)


def test_adam_skips_gradless_and_t_increments():   # src/optim.rs:86, 93
    p = T.new(np.ones(2), (2,)).requires_grad_()
    opt = R.Adam([p], 1e-3)
    opt.step()
    assert opt.t == 1
    np.testing.assert_array_equal(p.data(), [1, 1])


def test_adamw_decays_gradless_params():    # src/optim.rs:154-161 (A4)
    p = T.new(np.ones(2), (2,)).requires_grad_()
    opt = R.AdamW([p], 0.1, None, None, 0.5)
    opt.step()
    np.testing.assert_allclose(p.data(), [0.95, 0.95], rtol=1e-6)


def test_schedulers():                      # src/optim.rs:392-422
    s = R.StepLR(0.1, 3, 0.5)
    for _ in range(3):
        s.step()
    assert float(s.get_lr()) == pytest.approx(0.05, abs=1e-6)
    e = R.ExponentialLR(0.1, 0.9)
    e.step()
    assert float(e.get_lr()) == pytest.approx(0.09, abs=1e-6)
    c = R.CosineAnnealingLR(0.1, 10, None)
    prev = 0.1
    for _ in range(5):
        c.step()
        assert float(c.get_lr()) < prev
        prev = float(c.get_lr())
    p = R.ReduceLROnPlateau(0.1, 0.5, 2, None, None)
    p.step(1.0)
    p.step(1.0)
    p.step(1.0)
    assert float(p.get_lr()) == pytest.approx(0.05, abs=1e-6)


def test_train_epoch_smoke():               # src/train.rs:388-417: 100x784 randn, MLP 784-128-10, B=32, Adam 1e-3
    rng = np.random.default_rng(0)
    model = R.build_mlp([784, 128, 10], rng)
    opt = R.Adam(model.parameters(), 1e-3)
    X = rng.standard_normal((100, 784)).astype(np.float32)
    y = rng.integers(0, 10, 100).astype(np.float32)
    for s in range(0, 100, 32):
        xb, yb = X[s:s + 32], y[s:s + 32]          # last batch is ragged (4 rows), src/data/mnist.rs:377
        loss, acc = R.train_step(model, opt, T.new(xb, xb.shape), T.new(yb, yb.shape))
        assert loss > 0 and 0.0 <= acc <= 1.0


def test_mini_mnist_step():                 # tests/smoke.rs:461-502
    rng = np.random.default_rng(1)
    lin = R.Linear(784, 10, True, rng)
    x = T.new(rng.random((4, 784)), (4, 784))
    y = T.new([1, 0, 4, 9], (4,))
    logits = lin.forward(x)
    loss = R.cross_entropy_loss(logits, y)
    loss.backward()
    assert loss.data()[0] > 0
    assert lin.weight.grad() is not None and lin.bias.grad() is not None
    assert lin.weight.grad().shape == (7840,)


def test_mlp_tape_node_sequence():          # 13 nodes for Linear-ReLU-Linear + CE
    rng = np.random.default_rng(2)
    model = R.build_mlp([784, 128, 10], rng)
    x = T.new(rng.random((8, 784)), (8, 784))
    R.Tape.reset()
    logits = model.forward(x)
    assert len(R.Tape.nodes) == 7
    R.cross_entropy_loss(logits, T.new(np.zeros(8), (8,)))
    assert len(R.Tape.nodes) == 13


def test_conv_strict_reference_only_bias_gets_grad():   # src/tensor.rs:1725, 2075 (A1)
    rng = np.random.default_rng(3)
    conv = R.Conv2dReLU(1, 4, (3, 3), (1, 1), (1, 1), None, None, True, rng)
    x = T.new(rng.random((2, 1, 6, 6)), (2, 1, 6, 6))
    out = conv.forward(x)
    out.sum().backward()
    assert conv.weight.grad() is None
    assert conv.bias.grad() is not None


def test_conv_matches_torch_with_A2_weight_layout():    # src/tensor.rs:1262 (A2)
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(4)
    cin, cout = 3, 5
    x = rng.standard_normal((2, cin, 7, 7)).astype(np.float32)
    w = rng.standard_normal((cout, cin, 3, 3)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    out = T.new(x, x.shape).conv2d(T.new(w, w.shape), T.new(b, b.shape), (1, 1), (1, 1), (1, 1)).numpy()
    wt = torch.from_numpy(w.reshape(-1).reshape(cin, 3, 3, cout)).permute(3, 0, 1, 2).contiguous()
    ref = torch.nn.functional.conv2d(torch.from_numpy(x), wt, torch.from_numpy(b), padding=1).numpy()
    np.testing.assert_allclose(out, ref, atol=2e-5)


def test_conv_full_adjoint_matches_torch_autograd():
    torch = pytest.importorskip("torch")
    R.Config.strict_reference_conv = False
    rng = np.random.default_rng(5)
    cin, cout = 2, 3
    x = rng.standard_normal((2, cin, 6, 6)).astype(np.float32)
    w = rng.standard_normal((cout, cin, 3, 3)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    X = T.new(x, x.shape).requires_grad_()
    W = T.new(w, w.shape).requires_grad_()
    Bb = T.new(b, b.shape).requires_grad_()
    out = X.conv2d_relu(W, Bb, (1, 1), (1, 1), (1, 1))
    out.sum().backward()
    tx = torch.from_numpy(x).requires_grad_()
    tw2 = torch.from_numpy(w.reshape(-1).copy()).requires_grad_()
    tb = torch.from_numpy(b).requires_grad_()
    wt = tw2.reshape(cin, 3, 3, cout).permute(3, 0, 1, 2)
    torch.relu(torch.nn.functional.conv2d(tx, wt, tb, padding=1)).sum().backward()
    np.testing.assert_allclose(X.grad(), tx.grad.numpy().reshape(-1), atol=1e-4)
    np.testing.assert_allclose(W.grad(), tw2.grad.numpy(), atol=1e-4)
    np.testing.assert_allclose(Bb.grad(), tb.grad.numpy(), atol=1e-4)


def test_maxpool_first_max_and_overwrite_backward():    # src/tensor.rs:1451, 1498-1514 (A6)
    x = T.new([1, 1, 1, 1,
               0, 5, 5, 0,
               0, 0, 2, 3,
               0, 0, 3, 3], (1, 1, 4, 4)).requires_grad_()
    _pad = x.reshape((1, 1, 4, 4))
    y = x.max_pool2d((2, 2), None, (0, 0))
    np.testing.assert_array_equal(y.data(), [5, 5, 0, 3])
    x.set_grad(np.full(16, 7.0, np.float32))            # pre-existing grad is overwritten by the plane zeroing
    y.sum().backward()
    g = x.grad().reshape(4, 4)
    expect = np.zeros((4, 4), np.float32)
    expect[1, 1] = 1; expect[1, 2] = 1; expect[2, 0] = 1; expect[2, 3] = 1
    np.testing.assert_array_equal(g, expect)


def test_pools_match_torch():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(6)
    x = rng.standard_normal((2, 3, 8, 8)).astype(np.float32)
    X = T.new(x, x.shape)
    np.testing.assert_array_equal(X.max_pool2d((2, 2), (2, 2), (0, 0)).numpy(),
                                  torch.nn.functional.max_pool2d(torch.from_numpy(x), 2).numpy())
    np.testing.assert_allclose(R.AdaptiveAvgPool2d.global_().forward(X).numpy(),
                               torch.from_numpy(x).mean(dim=(2, 3), keepdim=True).numpy(), atol=1e-6)


def test_mlp_step_matches_torch_autograd():
    """Everything in the MLP step is standard math: cross-check loss and grads with PyTorch."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(7)
    model = R.build_mlp([784, 128, 10], rng)
    x = rng.random((16, 784)).astype(np.float32)
    y = rng.integers(0, 10, 16)
    R.Tape.reset()
    logits = model.forward(T.new(x, x.shape))
    loss = R.cross_entropy_loss(logits, T.new(y.astype(np.float32), (16,)))
    loss.backward()
    ps = [torch.from_numpy(p.numpy().copy()).requires_grad_() for p in model.parameters()]
    h = torch.relu(torch.from_numpy(x) @ ps[0].T + ps[1])
    tl = torch.nn.functional.cross_entropy(h @ ps[2].T + ps[3], torch.from_numpy(y))
    tl.backward()
    assert loss.data()[0] == pytest.approx(tl.item(), rel=1e-5)
    for p, tp in zip(model.parameters(), ps):
        np.testing.assert_allclose(p.grad(), tp.grad.numpy().reshape(-1), atol=2e-6, rtol=1e-4)


def test_powi_matches_float32_square_and_multiply():    # src/optim.rs:88-89 `powi`
    assert float(R.powi_f32(0.9, 1)) == float(np.float32(0.9))
    assert float(R.powi_f32(0.999, 3)) == float(np.float32(np.float32(0.999) * np.float32(np.float32(0.999) * np.float32(0.999)))) or \
        math.isclose(float(R.powi_f32(0.999, 3)), 0.999 ** 3, rel_tol=1e-6)


# ---- sigmoid / mean / pow / sqrt / BCE / MSE / one-hot cross-entropy -------------------------------
def test_sqrt_pow_kats():                   # tests/smoke.rs:380-406
    x = T.new([1.0, 4.0, 9.0], (3,)).requires_grad_()
    y = x.sqrt()
    np.testing.assert_allclose(y.data(), [1.0, 2.0, 3.0], atol=1e-6)
    y.sum().backward()
    np.testing.assert_allclose(x.grad(), [0.5, 0.25, 1.0 / 6.0], rtol=1e-6)        # d/dx sqrt(x) = 0.5 / sqrt(x)
    np.testing.assert_allclose(T.new([2.0, 3.0], (2,)).pow(2.0).data(), [4.0, 9.0], rtol=1e-6)


def test_one_hot_kat():                     # src/loss.rs:343-356
    oh = R.one_hot(T.new([0, 2, 1], (3,)), 3)
    assert oh.shape == (3, 3)
    np.testing.assert_array_equal(oh.data(), [1, 0, 0, 0, 0, 1, 0, 1, 0])


def test_bce_mse_sigmoid_against_torch():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(5)
    p = rng.uniform(0.02, 0.98, 12).astype(np.float32)
    t = rng.integers(0, 2, 12).astype(np.float32)
    P = T.new(p, (12,)).requires_grad_()
    l = R.bce_loss(P, T.new(t, (12,)))
    l.backward()
    pt = torch.tensor(p, requires_grad=True)
    lt = torch.nn.functional.binary_cross_entropy(pt, torch.tensor(t))
    lt.backward()
    assert abs(float(l.data()[0]) - lt.item()) < 1e-6
    np.testing.assert_allclose(P.grad(), pt.grad.numpy(), rtol=2e-5)
    R.Tape.reset()
    a = rng.standard_normal(10).astype(np.float32)
    b = rng.standard_normal(10).astype(np.float32)
    A = T.new(a, (10,)).requires_grad_()
    m = R.mse_loss(A, T.new(b, (10,)))
    m.backward()
    at = torch.tensor(a, requires_grad=True)
    mt = torch.nn.functional.mse_loss(at, torch.tensor(b))
    mt.backward()
    assert abs(float(m.data()[0]) - mt.item()) < 1e-6
    np.testing.assert_allclose(A.grad(), at.grad.numpy(), rtol=2e-5, atol=1e-7)
    R.Tape.reset()
    z = T.new(a, (10,)).requires_grad_()
    s = z.sigmoid()
    s.sum().backward()
    zt = torch.tensor(a, requires_grad=True)
    st = torch.sigmoid(zt)
    st.sum().backward()
    np.testing.assert_allclose(s.data(), st.detach().numpy(), rtol=1e-6)
    np.testing.assert_allclose(z.grad(), zt.grad.numpy(), rtol=2e-5)


def test_ce_onehot_equals_index_ce():       # src/loss.rs:202-245 vs :136-195 on the KAT of tests/smoke.rs:450-458
    lg = T.new([2, 1, 0, 0, 1, 2], (2, 3)).requires_grad_()
    l = R.cross_entropy_loss_onehot(lg, R.one_hot(T.new([0, 2], (2,)), 3))
    l.backward()
    assert abs(float(l.data()[0]) - 0.407606) < 2e-6
    np.testing.assert_allclose(lg.grad(), [-0.1673795, 0.1223642, 0.0450153, 0.0450153, 0.1223642, -0.1673795], atol=2e-6)
