"""The Tensor / Tape / Module / loss / optimizer handles of the host C ABI (tp_tensor_*, tp_module_forward, tp_loss,
tp_optimizer_* in include/taper_b200_host.h) — the surface a Rust shim binds so that examples/train_mnist.rs-style code runs
unmodified — against the reference's own known-answer tests and the oracle's loop."""
import json
import os

import numpy as np
import pytest

from oracle import taper_ref as R
from test_step_gpu import close, make_pair, defaults  # noqa: F401

pytestmark = pytest.mark.gpu
F32 = np.float32
KATS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")))


def test_matmul_kat_and_gradient_shapes():                  # tests/smoke.rs:46-70
    from taper_b200 import host
    host.tape_reset()
    k = KATS["matmul"]
    a = host.Tensor(k["a"], k["a_shape"], requires_grad=True)
    b = host.Tensor(k["b"], k["b_shape"], requires_grad=True)
    c = a.matmul(b)
    np.testing.assert_allclose(c.data().reshape(-1), k["c"], atol=k["tol"])
    assert c.data().reshape(-1)[0] == pytest.approx(58.0, abs=1e-4) and c.data().reshape(-1)[3] == pytest.approx(154.0, abs=1e-4)
    c.backward()
    assert a.grad().shape == (2, 3) and b.grad().shape == (3, 2)
    np.testing.assert_allclose(a.grad().reshape(-1), k["derived_grad_a"], atol=k["tol"])
    np.testing.assert_allclose(b.grad().reshape(-1), k["derived_grad_b"], atol=k["tol"])


def test_elementwise_graph_matches_the_oracle():
    """z = sum((a * b + a) / b) with exp / log / relu in between: values and both gradients against the oracle's tape."""
    from taper_b200 import host
    rng = np.random.default_rng(0)
    av, bv = rng.standard_normal((4, 6)).astype(F32), (rng.random((4, 6)) + 0.5).astype(F32)
    host.tape_reset()
    a, b = host.Tensor(av, requires_grad=True), host.Tensor(bv, requires_grad=True)
    z = (((a * b + a) / b).relu().exp() + b.log()).sum()
    z.backward()
    R.Tape.reset()
    ra, rb = R.Tensor.new(av, av.shape).requires_grad_(), R.Tensor.new(bv, bv.shape).requires_grad_()
    rz = (((ra * rb + ra) / rb).relu().exp() + rb.log()).sum()
    rz.backward()
    assert z.data()[0] == pytest.approx(float(rz.data()[0]), rel=1e-5)
    close(a.grad(), ra.grad(), 1e-5)
    close(b.grad(), rb.grad(), 1e-5)
    assert host.tape_len() >= 7
    host.tape_reset()
    assert host.tape_len() == 0


def test_example_training_loop_through_the_handles():
    """examples/train_mnist.rs:28-135 written against the handle API: model.forward(&images), cross_entropy_loss, accuracy,
    loss.backward(), optimizer.step(), optimizer.zero_grad(), Tape::reset() — loss and accuracy per step and the first
    step's gradients against the oracle (Adam(1e-3, wd 1e-4); eps 0.1 so the parameters compare at 1e-4 afterwards)."""
    from taper_b200 import host
    ref, m = make_pair(lambda r: R.build_mlp([784, 128, 64, 10], r), host.MLP_EXAMPLE, 1)
    params = m.parameters()
    opt = host.Optimizer("adam", params, 1e-3, eps=0.1, weight_decay=1e-4)
    ropt = R.Adam(ref.parameters(), 1e-3, None, 0.1, 1e-4)
    rng = np.random.default_rng(2)
    for step in range(4):
        x = rng.random((256, 784)).astype(F32)
        y = rng.integers(0, 10, 256).astype(F32)
        host.tape_reset()
        images, labels = host.Tensor(x), host.Tensor(y)
        logits = m.forward_tensor(images)
        loss = host.loss("cross_entropy", logits, labels)
        acc = host.accuracy(logits, labels)
        loss.backward()
        R.Tape.reset()
        rl = R.cross_entropy_loss(ref.forward(R.Tensor.new(x, x.shape)), R.Tensor.new(y, y.shape))
        racc = R.accuracy(ref.forward(R.Tensor.new(x, x.shape)), R.Tensor.new(y, y.shape))
        R.Tape.reset()
        rl = R.cross_entropy_loss(ref.forward(R.Tensor.new(x, x.shape)), R.Tensor.new(y, y.shape))
        rl.backward()
        assert loss.data()[0] == pytest.approx(float(rl.data()[0]), rel=1e-5), step
        assert abs(acc - float(racc)) <= 1.5 / 256
        if step == 0:
            for p, rp in zip(params, ref.parameters()):
                close(p.grad(), rp.grad(), 1e-4)
        opt.step(); opt.zero_grad()
        ropt.step(); ropt.zero_grad()
        assert all(p.grad() is None for p in params)                # zero_grad: grad = None (src/tensor.rs:531-533)
    for p, rp in zip(params, ref.parameters()):
        close(p.data(), rp.data(), 1e-4)
    assert opt.get_lr() == pytest.approx(1e-3)


def test_conv_module_forward_and_argmax_through_the_handles():
    from taper_b200 import host
    ref, m = make_pair(R.build_cnn2, host.CNN2, 2)
    x = np.random.default_rng(3).random((5, 1, 28, 28)).astype(F32)
    host.tape_reset()
    out = m.forward_tensor(host.Tensor(x))
    rout = ref.forward(R.Tensor.new(x, x.shape))
    close(out.data(), rout.data(), 1e-4)
    np.testing.assert_array_equal(out.argmax(1).data().reshape(-1), rout.argmax(1).data().reshape(-1))
    assert out.reshape((50,)).shape == (50,)
