"""Whole-step parity of the host layer (Tensor / Tape / nn / loss / optim / train, C++ over the C ABI)
against the CPU oracle: forward record + reverse replay + optimizer step on identical seeded inputs,
for BASELINE.json's configs (sizes the oracle finishes in seconds) and their ragged tails.
Tolerance (north_star): max|x - ref| <= 1e-4 * max(||ref||_inf, 1e-6) per tensor; counts are exact.
"""
import numpy as np
import pytest

from oracle import taper_ref as R

pytestmark = pytest.mark.gpu
F32 = np.float32


def close(got, ref, tol=1e-4, what=""):
    got = np.asarray(got, np.float64).reshape(-1)
    ref = np.asarray(ref, np.float64).reshape(-1)
    assert got.shape == ref.shape, what
    scale = max(np.max(np.abs(ref)), 1e-6)
    err = np.max(np.abs(got - ref))
    assert err <= tol * scale, f"{what}: max abs err {err:.3e} > {tol} * {scale:.3e}"


def close_after_adam(got, ref, lr, steps, what=""):
    """Parameters after `steps` default-eps Adam steps on two INDEPENDENT trajectories (a wiring check, not the parity gate).

    The parity gate is in three tight pieces: gradients agree to 1e-4 of ||g||_inf wherever the parameters agree
    (test_teacher_forced_gradients), the optimizer kernels agree with the oracle to rounding on identical gradients
    (test_kernels_gpu.py), and the whole trainer loop — bias corrections, step counter, weight decay, graph replay — agrees
    to 1e-4 over a free run when Adam's normaliser is made benign (eps = 0.1, test_*_adam_large_eps below).
    With eps = 1e-8 the update lr*m/(sqrt(v)+eps) divides by |g|: wherever |g| is within ~100x of its own fp32 summation
    noise (~1e-8 here; a third of cfg4's first-layer weights) the two trajectories move apart by a few % of one update per
    step, which is already more than 1e-4*||W||_inf = 0.5 % of an update.  The reference differs from itself in the same
    way under another sgemm summation order.  A wiring error (a skipped or doubled step, a wrong bias
    correction or counter) moves nearly every element by a large fraction of lr, so: at most 35 % of the elements beyond the
    plain 1e-4 bound, the median error within it, and nobody further away than one update per step."""
    got = np.asarray(got, np.float64).reshape(-1)
    ref = np.asarray(ref, np.float64).reshape(-1)
    scale = max(np.max(np.abs(ref)), 1e-6)
    err = np.abs(got - ref)
    frac = np.mean(err > 1e-4 * scale)
    assert frac <= 0.35, f"{what}: {frac:.2%} of the elements exceed 1e-4 relative (max {err.max():.3e}, lr {lr})"
    assert np.median(err) <= 1e-4 * scale, f"{what}: median abs err {np.median(err):.3e}"
    assert err.max() <= 1e-4 * scale + 1.05 * lr * steps, f"{what}: max abs err {err.max():.3e} (lr {lr}, {steps} steps)"


@pytest.fixture(autouse=True)
def defaults():
    from taper_b200 import host
    host.config(conv_full_adjoint=0, fuse_linear_relu=1, reference_op_sequence=0, gemm_mode=1)
    R.Config.strict_reference_conv = True
    R.Tape.reset()
    yield
    host.config(conv_full_adjoint=0, fuse_linear_relu=1, reference_op_sequence=0, gemm_mode=1)
    R.Config.strict_reference_conv = True


def make_pair(builder, spec, seed=0):
    from taper_b200 import host
    rng = np.random.default_rng(seed)
    ref = builder(rng)
    for p in ref.parameters():                     # non-zero biases make the bias path observable
        if len(p.shape) == 1:
            p._data[:] = rng.standard_normal(p._data.size).astype(F32) * F32(0.05)
    m = host.Model(spec, seed)
    m.load_from_oracle(ref)
    return ref, m


def batches(rng, n_steps, batch, sample_shape, ragged=None, ragged_at=None):
    for i in range(n_steps):
        b = ragged if (ragged and i == (n_steps - 1 if ragged_at is None else ragged_at)) else batch
        x = rng.random((b,) + tuple(sample_shape)).astype(F32)
        y = rng.integers(0, 10, b).astype(F32)
        yield x, y


def oracle_opt(kind, params, lr, wd, eps=None):
    if kind == "sgd":
        return R.SGD(params, lr)
    if kind == "adam":
        return R.Adam(params, lr, None, eps, wd)
    return R.AdamW(params, lr, None, eps, wd)


def run_parity(builder, spec, kind, lr, wd, batch, sample_shape, steps, ragged=None, tol=1e-4, use_graph=True, seed=0, eps=None,
               fused=False, ragged_at=None):
    """fused=False: tape + CUDA-graph path (one kernel per op); fused=True: the device tape (tp_step_*, one persistent
    kernel per step) where the model qualifies."""
    from taper_b200 import host
    ref, m = make_pair(builder, spec, seed)
    tr = host.Trainer(m, kind, lr=lr, weight_decay=wd, eps=1e-8 if eps is None else eps)
    tr.set_use_graph(use_graph)
    tr.set_use_fused(fused)
    opt = oracle_opt(kind, ref.parameters(), lr, wd, eps)
    rng = np.random.default_rng(seed + 1)
    for i, (x, y) in enumerate(batches(rng, steps, batch, sample_shape, ragged, ragged_at)):
        # rows whose two largest logits agree to within the parity tolerance may legitimately break the tie either way
        # (after the first Adam step the parameters themselves are only comparable to a few % of one update, see
        # close_after_adam, so the margin is taken 20x wider there)
        lg = ref.forward(R.Tensor.new(x, x.shape)).numpy().astype(np.float64)
        top2 = np.sort(lg, axis=1)[:, -2:]
        margin = (1e-4 if (kind == "sgd" or i == 0) else 2e-3) * np.max(np.abs(lg))
        near_ties = int(np.sum(top2[:, 1] - top2[:, 0] <= margin))
        loss_ref, acc_ref = R.train_step(ref, opt, R.Tensor.new(x, x.shape), R.Tensor.new(y, y.shape))
        loss, correct = tr.step(x, y)
        assert abs(loss - loss_ref) <= tol * max(abs(loss_ref), 1e-6), f"step {i}: loss {loss} vs {loss_ref}"
        assert abs(correct - round(acc_ref * x.shape[0])) <= near_ties, \
            f"step {i}: correct {correct} vs {acc_ref * x.shape[0]} ({near_ties} near-tied rows)"
    for j, p in enumerate(ref.parameters()):
        if kind == "sgd" or (eps is not None and eps >= 1e-2):
            close(m.get_param(j), p.data(), tol, f"param {j} after {steps} steps")
        else:
            close_after_adam(m.get_param(j), p.data(), lr, steps, f"param {j} after {steps} steps")
    return tr, m


@pytest.mark.parametrize("fused", [False, True])
def test_ragged_batch_in_the_middle_of_an_epoch_sequence(fused):
    """The last batch of an epoch is ragged (60000 = 234 x 256 + 96) and the next epoch returns to the full size: the steps
    compiled for the two batch sizes (device tapes / captured graphs) must coexist — a later, smaller one may not take
    anything (shared-memory opt-in, scratch, result slots) away from an earlier, larger one."""
    from taper_b200 import host
    tr, _ = run_parity(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, "sgd", 0.01, 0.0, 512, (784,), 8, ragged=96,
                       ragged_at=3, fused=fused)
    assert tr.fused_steps() == (8 if fused else 0)


# ---- cfg1: MLP 784-128-10, batch 64, SGD (src/train.rs:390-394 model) --------------------------------------
@pytest.mark.parametrize("gemm_mode", [0, 1])
def test_cfg1_mlp_sgd_b64(gemm_mode):
    from taper_b200 import host
    host.config(gemm_mode=gemm_mode)
    run_parity(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, "sgd", 0.01, 0.0, 64, (784,), 12, ragged=32)


def test_cfg1_mlp_sgd_b64_fused():
    from taper_b200 import host
    tr, _ = run_parity(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, "sgd", 0.01, 0.0, 64, (784,), 12, ragged=32,
                       fused=True)
    assert tr.fused_steps() == 12 and tr.graph_replays() == 0


# ---- cfg2: same MLP, batch 512, Adam ---------------------------------------------------------------------------
@pytest.mark.parametrize("wd", [0.0, 1e-4])
def test_cfg2_mlp_adam_b512(wd):
    from taper_b200 import host
    tr, _ = run_parity(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, "adam", 1e-3, wd, 512, (784,), 8, ragged=96)
    assert tr.graph_replays() >= 5             # steps 3.. of the B=512 shape ran as CUDA-graph replays


@pytest.mark.parametrize("wd", [0.0, 1e-4])
def test_cfg2_mlp_adam_b512_fused(wd):
    from taper_b200 import host
    tr, _ = run_parity(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, "adam", 1e-3, wd, 512, (784,), 8, ragged=96,
                       fused=True)
    assert tr.fused_steps() == 8 and tr.graph_replays() == 0


@pytest.mark.parametrize("kind,wd", [("adam", 0.0), ("adam", 1e-4), ("adamw", 1e-2)])
def test_cfg2_mlp_adam_large_eps_tight(kind, wd):
    """Free-running trainer loop at the plain 1e-4 bound: eps = 0.1 >> sqrt(v) keeps Adam's update linear in the gradient, so
    nothing amplifies summation noise while t, both bias corrections, lr, weight decay and the graph replay all still act."""
    from taper_b200 import host
    tr, _ = run_parity(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, kind, 0.1, wd, 512, (784,), 10,
                       ragged=96, eps=0.1)
    assert tr.graph_replays() >= 7


@pytest.mark.parametrize("kind,wd", [("adam", 0.0), ("adam", 1e-4), ("adamw", 1e-2), ("sgd", 0.0)])
def test_cfg2_mlp_large_eps_tight_fused(kind, wd):
    """The same free-running loop through the device tape: exact fp32 FMA products, so the plain 1e-4 bound with room."""
    from taper_b200 import host
    tr, _ = run_parity(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, kind, 0.1 if kind != "sgd" else 0.05, wd, 512,
                       (784,), 10, ragged=96, eps=0.1, fused=True)
    assert tr.fused_steps() == 10


def test_cfg4_mlp_wide_adam_large_eps_tight():
    from taper_b200 import host
    # 2M hidden activations per step: a few land within summation noise of the ReLU threshold on every step (see
    # relu_tie_slack) and each such flip moves one sample's contribution, ~1e-3 of a gradient row; hence 3e-4, which a
    # wiring error (>= 10 % of an update ~ 1e-2 of ||W||_inf here) still cannot meet.
    run_parity(lambda r: R.build_mlp([784, 1024, 1024, 10], r), host.MLP_784_1024_1024_10, "adam", 0.1, 0.0, 1024, (784,), 5,
               eps=0.1, tol=3e-4)


def test_cfg2_eager_equals_graph_bitwise():
    from taper_b200 import host
    rng = np.random.default_rng(3)
    data = list(batches(rng, 6, 512, (784,)))
    outs = []
    for use_graph in (False, True):
        _, m = make_pair(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, 5)
        tr = host.Trainer(m, "adam", lr=1e-3)
        tr.set_use_graph(use_graph)
        tr.set_use_fused(False)
        losses = [tr.step(x, y) for x, y in data]
        outs.append((losses, [m.get_param(i) for i in range(m.num_params())]))
    assert outs[0][0] == outs[1][0]
    for a, b in zip(outs[0][1], outs[1][1]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("fused", [False, True])
def test_example_mlp_adam_wd_b256(fused):       # examples/train_mnist.rs:28-61: 784-128-64-10, Adam(1e-3, wd 1e-4), B=256
    from taper_b200 import host
    tr, _ = run_parity(lambda r: R.build_mlp([784, 128, 64, 10], r), host.MLP_EXAMPLE, "adam", 1e-3, 1e-4, 256, (784,), 6, ragged=96,
                       fused=fused)
    assert tr.fused_steps() == (6 if fused else 0)


@pytest.mark.parametrize("dims,spec,batch", [
    ([784, 10], "linear:784:10", 200),                                                       # softmax regression: head only
    ([784, 64, 32, 16, 10], "linear:784:64,relu,linear:64:32,relu,linear:32:16,relu,linear:16:10", 130),
    ([100, 36, 12], "linear:100:36,relu,linear:36:12", 77),                                  # ragged tiles everywhere
])
def test_fused_step_other_chains_large_eps_tight(dims, spec, batch):
    from taper_b200 import host
    rng0 = np.random.default_rng(21)
    in_f = dims[0]
    ref, m = make_pair(lambda r: R.build_mlp(dims, r), spec, 4)
    tr = host.Trainer(m, "adam", lr=0.05, weight_decay=1e-3, eps=0.1)
    opt = R.Adam(ref.parameters(), 0.05, None, 0.1, 1e-3)
    for i in range(6):
        x = rng0.random((batch, in_f)).astype(F32)
        y = rng0.integers(0, dims[-1], batch).astype(F32)
        loss_ref, _ = R.train_step(ref, opt, R.Tensor.new(x, x.shape), R.Tensor.new(y, y.shape))
        loss, _ = tr.step(x, y)
        assert abs(loss - loss_ref) <= 1e-4 * abs(loss_ref), (i, loss, loss_ref)
    assert tr.fused_steps() == 6
    for j, p in enumerate(ref.parameters()):
        close(m.get_param(j), p.data(), 1e-4, f"param {j}")


def test_fused_step_falls_back_when_model_does_not_qualify():
    """No ReLU-free classifier tail / too much work per step: the trainer silently keeps the tape + graph path."""
    from taper_b200 import host
    rng = np.random.default_rng(2)
    m = host.Model("linear:784:128,relu,linear:128:10,relu", 0)             # ends in a ReLU: not a classifier chain
    tr = host.Trainer(m, "adam", lr=1e-3)
    x = rng.random((64, 784)).astype(F32); y = rng.integers(0, 10, 64).astype(F32)
    for _ in range(3):
        tr.step(x, y)
    assert tr.fused_steps() == 0 and tr.graph_replays() >= 1
    m2 = host.Model(host.MLP_784_1024_1024_10, 0)                            # 9.8 GFLOP per step: not the persistent kernel
    tr2 = host.Trainer(m2, "adam", lr=1e-3)                                  # but the wide plan of tcgen05 kernels (step_wide.cu)
    x = rng.random((1024, 784)).astype(F32); y = rng.integers(0, 10, 1024).astype(F32)
    for _ in range(3):
        tr2.step(x, y)
    assert tr2.fused_steps() == 3 and tr2.fused_kind() == 2
    m3 = host.Model("linear:784:1024,relu,linear:1024:2048,relu,linear:2048:10", 0)      # classifier input wider than 1024: neither
    tr3 = host.Trainer(m3, "adam", lr=1e-3)
    for _ in range(3):
        tr3.step(x, y)
    assert tr3.fused_steps() == 0


def test_fused_step_matches_graph_path_closely():
    """Device tape vs one-kernel-per-op path (exact-fp32 GEMM mode) on the same data: same algorithm, different summation
    order only."""
    from taper_b200 import host
    host.config(gemm_mode=0)
    rng = np.random.default_rng(3)
    data = list(batches(rng, 6, 512, (784,)))
    outs = []
    for fused in (False, True):
        _, m = make_pair(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, 5)
        tr = host.Trainer(m, "adam", lr=0.05, eps=0.1)
        tr.set_use_fused(fused)
        losses = [tr.step(x, y) for x, y in data]
        outs.append((losses, [m.get_param(i) for i in range(m.num_params())]))
    for (la, ca), (lb, cb) in zip(outs[0][0], outs[1][0]):
        assert abs(la - lb) <= 2e-6 * abs(la) and ca == cb
    for a, b in zip(outs[0][1], outs[1][1]):
        close(b, a, 1e-5)


def test_cfg4_mlp_wide_adam_b1024():
    from taper_b200 import host
    run_parity(lambda r: R.build_mlp([784, 1024, 1024, 10], r), host.MLP_784_1024_1024_10, "adam", 1e-3, 0.0, 1024, (784,), 3)


def test_reference_op_sequence_records_reference_nodes_and_matches():
    """Linear as transpose -> matmul -> add_broadcast (src/nn.rs:54-60) gives the same numbers as the fused node."""
    from taper_b200 import host
    rng = np.random.default_rng(11)
    x = rng.random((64, 784)).astype(F32)
    y = rng.integers(0, 10, 64).astype(F32)
    res = {}
    for ref_seq in (0, 1):
        host.config(reference_op_sequence=ref_seq)
        ref, m = make_pair(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, 2)
        loss, correct, tape_len = m.loss_backward(x, y)
        res[ref_seq] = (loss, correct, tape_len, [m.get_grad(i) for i in range(4)])
    assert res[0][2] == 3                        # fused: Linear+ReLU, Linear, CE
    assert res[1][2] == 8                        # 2 x (transpose, matmul, add_broadcast) + relu + CE (the reference adds 5 dead log_softmax nodes)
    assert res[0][1] == res[1][1]
    assert res[0][0] == pytest.approx(res[1][0], rel=1e-6)
    R.Tape.reset()
    logits = ref.forward(R.Tensor.new(x, x.shape))
    l = R.cross_entropy_loss(logits, R.Tensor.new(y, y.shape))
    l.backward()
    for k in (0, 1):
        assert res[k][0] == pytest.approx(float(l.data()[0]), rel=1e-5)
        for g, p in zip(res[k][3], ref.parameters()):
            close(g, p.grad(), 1e-4)


# ---- gradients along the ORACLE's trajectory (teacher forcing): the tight parity statement -------------------------------
def relu_tie_slack(ref, x, y, batch, thr=8e-6):
    """ReLU is discontinuous in its mask: a hidden pre-activation within fp32 summation noise of 0 (|z| <= thr max|z|; the
    default 8e-6 is three times the measured 3xTF32 / fp32 GEMM error, CNNs pass 4e-5 = three times the 1.3e-5 measured after
    five bf16x3 conv layers) may legitimately land on either side — the reference itself would, under another sgemm summation
    order (the order is implementation-defined).  One such unit switches one sample's path through that unit on or
    off.  Returns, per parameter, 4 x the summed ||per-sample gradient||_inf of the samples that own a tied unit (0 when
    there is none, the usual case below cfg4's 2M hidden activations per step), and the tie count.  (Conv2dReLU layers: see
    conv_bias_tie_slack.)"""
    h = R.Tensor.new(x, x.shape)
    tied = np.zeros(batch, bool)
    for l in ref.layers:
        if isinstance(l, R.ReLU):
            z = h.numpy().reshape(batch, -1)
            tied |= (np.abs(z) <= thr * np.max(np.abs(z))).any(axis=1)
        h = l.forward(h)
    R.Tape.reset()
    slack = [0.0] * len(ref.parameters())
    for b in np.nonzero(tied)[0]:
        for p in ref.parameters():
            p.zero_grad()
        xb, yb = x[b:b + 1], y[b:b + 1]
        R.cross_entropy_loss(ref.forward(R.Tensor.new(xb, xb.shape)), R.Tensor.new(yb, yb.shape)).backward()
        for j, p in enumerate(ref.parameters()):
            if p.grad() is not None:
                slack[j] += 4.0 * float(np.max(np.abs(p.grad()))) / batch
        R.Tape.reset()
    for p in ref.parameters():
        p.zero_grad()
    return slack, int(tied.sum())


def conv_bias_tie_slack(ref, x, y, thr=4e-5):
    """Strict-reference CNNs: the only gradient that crosses a conv layer's ReLU mask is the LAST conv's bias
    gradient, bias.grad[c] = sum_{n,p} g[n,c,p] * [z[n,c,p] > 0] (src/ops.rs:358-370, src/tensor.rs:2003-2027).  A unit whose
    pre-activation z lies within the conv path's error of 0 (thr = 4e-5 of max|z|: three times the 1.3e-5 measured after five
    bf16x3 layers) may land on either side and moves bias.grad[c] by |g[n,c,p]|.  Returns {parameter index: that sum} and the
    number of such units, computed from the oracle's own activations and gradients."""
    R.Tape.reset()
    for p in ref.parameters():
        p.zero_grad()
    h = R.Tensor.new(x, x.shape)
    last = None
    for l in ref.layers:
        if isinstance(l, R.Conv2dReLU):
            z = h.conv2d(l.weight, l.bias, l.stride, l.padding, l.dilation).numpy()
            h = l.forward(h)
            last = (l, z, h)
        else:
            h = l.forward(h)
    R.cross_entropy_loss(h, R.Tensor.new(y, y.shape)).backward()
    out = {}
    n_units = 0
    if last is not None and last[2].grad() is not None:
        l, z, act = last
        g = np.asarray(act.grad()).reshape(z.shape)
        tied = np.abs(z) <= thr * np.max(np.abs(z))
        n_units = int(tied.sum())
        j = [id(p) for p in ref.parameters()].index(id(l.bias))
        out[j] = float(np.max(np.sum(np.abs(g) * tied, axis=(0, 2, 3))))
    R.Tape.reset()
    for p in ref.parameters():
        p.zero_grad()
    return out, n_units


@pytest.mark.parametrize("name,builder,spec,batch,shape,full", [
    ("cfg2", lambda r: R.build_mlp([784, 128, 10], r), "MLP_784_128_10", 512, (784,), 0),
    ("example", lambda r: R.build_mlp([784, 128, 64, 10], r), "MLP_EXAMPLE", 256, (784,), 0),
    ("cfg4", lambda r: R.build_mlp([784, 1024, 1024, 10], r), "MLP_784_1024_1024_10", 1024, (784,), 0),
    ("cnn2_full", R.build_cnn2, "CNN2", 16, (1, 28, 28), 1),
    ("cnn5_strict", R.build_cnn5, "CNN5", 8, (1, 28, 28), 0),
    ("cnn5_full", R.build_cnn5, "CNN5", 8, (1, 28, 28), 1),
])
def test_teacher_forced_gradients(name, builder, spec, batch, shape, full):
    """At every step of an oracle Adam trajectory the CUDA tape, started from the oracle's current parameters, must give the
    oracle's loss, correct count and every parameter gradient (None pattern included) within 1e-4 of ||g||_inf
    (plus relu_tie_slack for the samples that sit on a ReLU threshold)."""
    from taper_b200 import host
    host.config(conv_full_adjoint=full)
    R.Config.strict_reference_conv = not full
    ref, m = make_pair(builder, getattr(host, spec), 3)
    opt = R.Adam(ref.parameters(), 1e-3 if "cnn" not in name else 0.01)
    rng = np.random.default_rng(17)
    steps = 4 if batch >= 1024 or "cnn" in name else 6
    for i, (x, y) in enumerate(batches(rng, steps, batch, shape)):
        m.load_from_oracle(ref)
        m.zero_grad()
        loss, correct, _ = m.loss_backward(x, y)
        slack, n_tied = relu_tie_slack(ref, x, y, batch) if "cnn" not in name else ([0.0] * len(ref.parameters()), 0)
        if "cnn" in name and not full:
            extra, n_tied = conv_bias_tie_slack(ref, x, y)
            for j, v in extra.items():
                slack[j] += v
        R.Tape.reset()
        logits = ref.forward(R.Tensor.new(x, x.shape))
        l = R.cross_entropy_loss(logits, R.Tensor.new(y, y.shape))
        acc = R.accuracy(logits, R.Tensor.new(y, y.shape))
        l.backward()
        assert loss == pytest.approx(float(l.data()[0]), rel=1e-5), f"step {i}"
        lg = logits.numpy().astype(np.float64)
        top2 = np.sort(lg, axis=1)[:, -2:]
        near = int(np.sum(top2[:, 1] - top2[:, 0] <= 1e-4 * np.max(np.abs(lg))))
        assert abs(correct - round(float(acc) * batch)) <= near, f"step {i}"
        for j, p in enumerate(ref.parameters()):
            g = m.get_grad(j)
            if p.grad() is None:
                assert g is None, f"step {i} param {j}: the reference leaves this gradient None (SURVEY A1)"
            else:
                scale = max(float(np.max(np.abs(p.grad()))), 1e-6)
                err = float(np.max(np.abs(np.asarray(g, np.float64).reshape(-1) - p.grad().astype(np.float64).reshape(-1))))
                assert err <= 1e-4 * scale + slack[j], \
                    f"step {i} grad {j}: max abs err {err:.3e} > 1e-4 * {scale:.3e} + {slack[j]:.3e} ({n_tied} samples on a ReLU threshold)"
        opt.step()
        opt.zero_grad()


# ---- cfg3: CNNs, strict_reference (A1) and full adjoint -----------------------------------------------------------
@pytest.mark.parametrize("full", [0, 1])
def test_cfg3_cnn2_adam(full):
    from taper_b200 import host
    host.config(conv_full_adjoint=full)
    R.Config.strict_reference_conv = not full
    run_parity(R.build_cnn2, host.CNN2, "adam", 0.01, 1e-4, 16, (1, 28, 28), 4, ragged=5, tol=2e-4)


@pytest.mark.parametrize("full", [0, 1])
def test_cfg3_cnn5_example_adam(full):           # examples/train_mnist_cnn.rs:35-137: Adam(0.01, wd 1e-4)
    from taper_b200 import host
    host.config(conv_full_adjoint=full)
    R.Config.strict_reference_conv = not full
    run_parity(R.build_cnn5, host.CNN5, "adam", 0.01, 1e-4, 8, (1, 28, 28), 3, tol=2e-4)


def test_cnn5_strict_reference_trains_only_conv5_bias_and_linears():
    from taper_b200 import host
    ref, m = make_pair(R.build_cnn5, host.CNN5, 4)
    rng = np.random.default_rng(0)
    x = rng.random((4, 1, 28, 28)).astype(F32)
    y = rng.integers(0, 10, 4).astype(F32)
    m.loss_backward(x, y)
    has = [m.get_grad(i) is not None for i in range(m.num_params())]
    # params: conv1..5 (w, b) = indices 0..9, then three linears (w, b) = 10..15
    assert has == [False] * 9 + [True] + [True] * 6


def test_cfg5_cnn2_adamw_decays_gradless_params():                       # src/optim.rs:154-161 (A4)
    from taper_b200 import host
    run_parity(R.build_cnn2, host.CNN2, "adamw", 0.01, 1e-2, 8, (1, 28, 28), 3, tol=2e-4)


def test_cnn_forward_only_matches_oracle():
    from taper_b200 import host
    ref, m = make_pair(R.build_cnn5, host.CNN5, 6)
    x = np.random.default_rng(1).random((6, 1, 28, 28)).astype(F32)
    close(m.forward(x), ref.forward(R.Tensor.new(x, x.shape)).data())


# ---- device-resident dataset path == host-fed path -------------------------------------------------------------------
@pytest.mark.parametrize("fused", [False, True])
def test_resident_dataset_gather_equals_host_batches(fused):
    """fused=True: rows are gathered inside the step kernel through perm + cursor; fused=False: tp_gather_batch."""
    from taper_b200 import host
    rng = np.random.default_rng(9)
    n, b = 1000, 128
    X = rng.random((n, 784)).astype(F32)
    Y = rng.integers(0, 10, n).astype(F32)
    perm = rng.permutation(n).astype(np.uint32)
    _, m1 = make_pair(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, 1)
    _, m2 = make_pair(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, 1)
    t1, t2 = host.Trainer(m1, "adam", lr=1e-3), host.Trainer(m2, "adam", lr=1e-3)
    t1.set_use_fused(fused); t2.set_use_fused(fused)
    t2.load_dataset(X, Y, perm)
    for s in range(12):                          # wraps around the dataset (cursor modulo n)
        idx = perm[(s * b + np.arange(b)) % n]
        r1 = t1.step(X[idx], Y[idx])
        t2.step_resident(b)
        r2 = t2.fetch()
        assert r1 == r2, f"step {s}: {r1} vs {r2}"
    for i in range(4):
        np.testing.assert_array_equal(m1.get_param(i), m2.get_param(i))
    assert t2.fused_steps() == (12 if fused else 0)


def test_async_pipeline_fifo_results():
    from taper_b200 import host
    rng = np.random.default_rng(2)
    _, m1 = make_pair(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, 1)
    _, m2 = make_pair(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, 1)
    t1, t2 = host.Trainer(m1, "sgd", lr=0.01), host.Trainer(m2, "sgd", lr=0.01)
    bufs = [(host.PinnedArray((64, 784)), host.PinnedArray((64,))) for _ in range(6)]
    for px, py in bufs:
        px.array[:] = rng.random((64, 784)).astype(F32)
        py.array[:] = rng.integers(0, 10, 64).astype(F32)
    sync = [t1.step(px.array, py.array) for px, py in bufs]
    for px, py in bufs:
        t2.step_async(px.array, py.array, pinned=True)
    assert t2.pending() == 6
    got = [t2.fetch() for _ in bufs]
    assert got == sync
    with pytest.raises(Exception):
        t2.fetch()


def test_eval_and_checkpoint_roundtrip(tmp_path):
    from taper_b200 import host
    ref, m = make_pair(lambda r: R.build_mlp([784, 128, 10], r), host.MLP_784_128_10, 8)
    tr = host.Trainer(m, "adam", lr=1e-3)
    rng = np.random.default_rng(5)
    x = rng.random((100, 784)).astype(F32)
    y = rng.integers(0, 10, 100).astype(F32)
    loss, correct = tr.eval(x, y)
    logits = ref.forward(R.Tensor.new(x, x.shape))
    assert loss == pytest.approx(float(R.cross_entropy_loss(logits, R.Tensor.new(y, y.shape)).data()[0]), rel=1e-5)
    assert correct == round(float(R.accuracy(logits, R.Tensor.new(y, y.shape))) * 100)
    path = tmp_path / "ckpt.txt"
    tr.save_checkpoint(path)
    lines = open(path).read().split("\n")
    assert lines[0] == "4" and lines[1] == "2 128 784"            # src/train.rs:272-281 text format
    before = [m.get_param(i) for i in range(4)]
    tr.step(x[:64], y[:64])
    assert not np.array_equal(before[0], m.get_param(0))
    tr.load_checkpoint(path)
    for i in range(4):
        np.testing.assert_array_equal(before[i], m.get_param(i))


def test_shape_errors_are_reported_not_fatal():
    from taper_b200 import host, TaperError
    m = host.Model(host.MLP_784_128_10, 0)
    with pytest.raises(TaperError, match="inner dimensions"):
        m.forward(np.zeros((4, 100), F32))
    with pytest.raises(TaperError, match="unknown layer"):
        host.Model("linear:4:4,bogus", 0)
