"""SURVEY 8(f)-4 on the GPU: sigmoid / pow / sqrt / mean / BCE kernels through the C ABI against the oracle, the host-level
bce / mse / one-hot cross-entropy losses through a Sigmoid MLP, and the reference's XOR demo (src/main.rs) ported 1:1."""
import os
import subprocess

import numpy as np
import pytest

from oracle import taper_ref as R

pytestmark = pytest.mark.gpu
F32 = np.float32
T = R.Tensor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def close(got, ref, tol=1e-5, what=""):
    got = np.asarray(got, np.float64).reshape(-1)
    ref = np.asarray(ref, np.float64).reshape(-1)
    assert np.max(np.abs(got - ref)) <= tol * max(np.max(np.abs(ref)), 1e-6), f"{what}: {np.max(np.abs(got - ref)):.3e}"


@pytest.fixture(scope="module")
def ctx():
    import taper_b200
    c = taper_b200.Ctx(0)
    yield c
    c.close()


@pytest.fixture(autouse=True)
def fresh():
    R.Tape.reset()
    yield
    R.Tape.reset()


@pytest.mark.parametrize("n", [1, 7, 1000, 4099])
def test_sigmoid_pow_mean_kernels(ctx, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 3).astype(F32)
    g = rng.standard_normal(n).astype(F32)
    X, G, Y, D = ctx.upload(x), ctx.upload(g), ctx.alloc(n), ctx.alloc(n)
    xt = T.new(x, (n,)).requires_grad_()
    s = xt.sigmoid()
    ctx.call("sigmoid_fwd", X, Y, n)
    close(Y.download(), s.data(), 2e-6, "sigmoid")
    ctx.call("sigmoid_bwd", Y, G, D, n, 0)
    close(D.download(), g * s.data() * (1 - s.data()), 2e-6, "sigmoid bwd")      # src/tensor.rs:627
    xp = np.abs(x) + F32(0.1)
    XP = ctx.upload(xp)
    for e in (0.5, 2.0, 3.0):
        ctx.call("pow_fwd", XP, Y, e, n)
        close(Y.download(), T.new(xp, (n,)).pow(e).data(), 2e-6, f"pow {e}")
        ctx.call("pow_bwd", XP, G, D, e, n, 0)
        close(D.download(), g * F32(e) * np.power(xp, F32(e - 1.0)), 2e-6, f"pow bwd {e}")
    one = ctx.alloc(1)
    ctx.call("mean_fwd", X, one, n)
    close(one.download(), T.new(x, (n,)).mean().data(), 2e-5, "mean")
    d0 = rng.standard_normal(n).astype(F32)
    D.upload(d0)
    ctx.call("mean_bwd", ctx.upload(np.array([0.75], F32)), D, n, 1)
    close(D.download(), d0 + F32(0.75) / F32(n), 1e-6, "mean bwd")


@pytest.mark.parametrize("n", [4, 37, 2048])
def test_bce_kernels_vs_oracle(ctx, n):
    rng = np.random.default_rng(n)
    p = rng.uniform(0.0, 1.0, n).astype(F32)
    p[0], p[-1] = 0.0, 1.0                                     # the clamp to [1e-7, 1 - 1e-7] (src/loss.rs:7, 19)
    t = rng.integers(0, 2, n).astype(F32)
    P, Tg, L = ctx.upload(p), ctx.upload(t), ctx.alloc(1)
    pt = T.new(p, (n,)).requires_grad_()
    tt = T.new(t, (n,)).requires_grad_()
    l = R.bce_loss(pt, tt)
    l.backward()
    ctx.call("bce_fwd", P, Tg, L, n)
    close(L.download(), l.data(), 1e-5, "bce")
    GP, GT = ctx.alloc(n), ctx.alloc(n)
    ctx.call("bce_bwd", P, Tg, ctx.upload(np.ones(1, F32)), GP, GT, n, 0, 0)
    close(GP.download(), pt.grad(), 1e-5, "bce dp")
    close(GT.download(), tt.grad(), 1e-5, "bce dt")


@pytest.mark.parametrize("kind", ["bce", "mse", "ce_onehot"])
def test_sigmoid_mlp_losses_vs_oracle(kind):
    from taper_b200 import host
    host.config(conv_full_adjoint=0, fuse_linear_relu=1, reference_op_sequence=0, gemm_mode=0)
    rng = np.random.default_rng(3)
    dims = [6, 8, 4]
    layers = [R.Linear(6, 8, True, rng), R.Sigmoid(), R.Linear(8, 4, True, rng)] + ([R.Sigmoid()] if kind == "bce" else [])
    ref = R.Sequential(layers)
    for p_ in ref.parameters():
        if len(p_.shape) == 1:
            p_._data[:] = (rng.standard_normal(p_._data.size) * 0.1).astype(F32)
    m = host.Model("linear:6:8,sigmoid,linear:8:4" + (",sigmoid" if kind == "bce" else ""), 0)
    m.load_from_oracle(ref)
    x = rng.standard_normal((5, dims[0])).astype(F32)
    if kind == "ce_onehot":
        tgt = np.eye(4, dtype=F32)[rng.integers(0, 4, 5)]
    else:
        tgt = rng.integers(0, 2, (5, 4)).astype(F32)
    out = ref.forward(T.new(x, x.shape))
    tt = T.new(tgt, tgt.shape)
    l = {"bce": R.bce_loss, "mse": R.mse_loss, "ce_onehot": R.cross_entropy_loss_onehot}[kind](out, tt)
    l.backward()
    loss = m.regression_backward(x, tgt, kind)
    assert abs(loss - float(l.data()[0])) <= 1e-5 * max(abs(float(l.data()[0])), 1e-6)
    for i, p_ in enumerate(ref.parameters()):
        close(m.get_grad(i), p_.grad(), 1e-4, f"{kind} grad {i}")
    host.config(gemm_mode=1)


def test_xor_demo_learns():                                    # src/main.rs: 2-4-1 Sigmoid MLP, BCE, SGD(0.10)
    exe = os.path.join(ROOT, "build", "xor")
    if not os.path.exists(exe):
        pytest.skip("build/xor not built (make examples)")
    r = subprocess.run([exe, "50000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "learned XOR" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
