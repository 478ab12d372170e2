"""Parity of the WIDE device tape (step_wide.cu: the training step of a wide MLP as a plan of tcgen05 kernels on bf16x3
pre-split operands) against the CPU oracle and against the tape + CUDA-graph path, through the C ABI.
Tolerance (north_star): max|x - ref| <= 1e-4 * max(||ref||_inf, 1e-6) per tensor; counts exact up to near-tied rows."""
import numpy as np
import pytest

from oracle import taper_ref as R
from test_step_gpu import close, close_after_adam, make_pair, batches, run_parity, relu_tie_slack, defaults  # noqa: F401

pytestmark = pytest.mark.gpu
F32 = np.float32

CFG4 = ([784, 1024, 1024, 10], "linear:784:1024,relu,linear:1024:1024,relu,linear:1024:10", 1024)
# > 1.5 GFLOP per step, so tp_step_kind picks the plan; ragged tiles in every dimension (batch 1000, widths 520 / 264)
ODD = ([784, 520, 264, 10], "linear:784:520,relu,linear:520:264,relu,linear:264:10", 1000)
NOBIAS = ([256, 1024, 512, 16], "linear:256:1024:nobias,relu,linear:1024:512,relu,linear:512:16:nobias", 1536)
LINEAR_HIDDEN = ([784, 1024, 512, 10], "linear:784:1024,linear:1024:512,relu,linear:512:10", 1024)    # a hidden layer without ReLU


def test_wide_plan_is_what_runs_for_configs3():
    from taper_b200 import host
    m = host.Model(CFG4[1], 0)
    tr = host.Trainer(m, "adam", lr=1e-3)
    rng = np.random.default_rng(0)
    x, y = rng.random((1024, 784)).astype(F32), rng.integers(0, 10, 1024).astype(F32)
    l0 = host.launches()
    for _ in range(3):
        tr.step(x, y)
    assert tr.fused_steps() == 3 and tr.fused_kind() == 2 and tr.graph_replays() == 0
    # input, 2 forward GEMMs, head, dX, one grouped launch for the 3 dW (the bias-gradient fold rides on its spare CTAs),
    # optimizer = 7 launches per step (+ one parameter split on the first)
    assert host.launches() - l0 == 3 * 7 + 1
    m2 = host.Model(host.MLP_784_128_10, 0)                                # small model: the persistent kernel
    tr2 = host.Trainer(m2, "adam", lr=1e-3)
    tr2.step(x[:64], y[:64])
    assert tr2.fused_kind() == 1


@pytest.mark.parametrize("dims,spec,batch", [CFG4, ODD, NOBIAS, LINEAR_HIDDEN])
def test_wide_teacher_forced_gradients(dims, spec, batch):
    """One SGD step with lr = 1024 (a power of two: p' = p - 1024 g is exact to an ulp of p') from the oracle's parameters
    recovers every gradient the plan computed; each must match the oracle's within 1e-4 of ||g||_inf (+ relu_tie_slack)."""
    from taper_b200 import host
    nc = dims[-1]
    ref, m = make_pair_for(dims, spec)
    rng = np.random.default_rng(17)
    opt = R.Adam(ref.parameters(), 1e-3)
    for i in range(3):
        x = rng.random((batch, dims[0])).astype(F32)
        y = rng.integers(0, nc, batch).astype(F32)
        m.load_from_oracle(ref)
        before = [m.get_param(j).astype(np.float64) for j in range(m.num_params())]
        tr = host.Trainer(m, "sgd", lr=1024.0)
        loss, correct = tr.step(x, y)
        assert tr.fused_kind() == 2
        after = [m.get_param(j).astype(np.float64) for j in range(m.num_params())]
        del tr
        slack, n_tied = relu_tie_slack(ref, x, y, batch)
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
            g = (before[j] - after[j]) / 1024.0
            gr = p.grad().astype(np.float64).reshape(g.shape)
            scale = max(float(np.max(np.abs(gr))), 1e-6)
            err = float(np.max(np.abs(g - gr)))
            ulp = 2.0 ** -23 * float(np.max(np.abs(after[j]))) / 1024.0      # what the recovery itself can lose
            assert err <= 1e-4 * scale + slack[j] + ulp, \
                f"step {i} grad {j}: max abs err {err:.3e} > 1e-4 * {scale:.3e} + {slack[j]:.3e} ({n_tied} samples on a ReLU threshold)"
        opt.step()
        opt.zero_grad()


def make_pair_for(dims, spec, seed=3):
    from taper_b200 import host
    rng = np.random.default_rng(seed)
    ref = R.build_mlp(dims, rng)
    layers_spec = [s for s in spec.split(",")]
    # mirror the spec: drop biases / ReLUs the oracle builder put in by default
    lin = [l for l in ref.layers if isinstance(l, R.Linear)]
    lin_specs = [s for s in layers_spec if s.startswith("linear")]
    new_layers = []
    li = 0
    for s in layers_spec:
        if s.startswith("linear"):
            l = lin[li]; li += 1
            if s.endswith(":nobias"):
                l.bias = None
            new_layers.append(l)
        elif s == "relu":
            new_layers.append(R.ReLU())
    ref.layers = new_layers
    assert len(lin_specs) == len(lin)
    for p in ref.parameters():
        if len(p.shape) == 1:
            p._data[:] = rng.standard_normal(p._data.size).astype(F32) * F32(0.05)
    m = host.Model(spec, seed)
    m.load_from_oracle(ref)
    return ref, m


@pytest.mark.parametrize("kind,wd", [("adam", 0.0), ("adam", 1e-4), ("adamw", 1e-2), ("sgd", 0.0)])
def test_wide_free_run_large_eps_tight(kind, wd):
    """Free-running trainer loop through the plan (t, bias corrections, lr, weight decay, operand planes refreshed by the
    optimizer kernel, a ragged batch compiling a second plan) with Adam's normaliser made benign (eps = 0.1)."""
    dims, spec, batch = ODD
    ref, m = make_pair_for(dims, spec, 5)
    from taper_b200 import host
    lr = 0.05 if kind == "sgd" else 0.1
    tr = host.Trainer(m, kind, lr=lr, weight_decay=wd, eps=0.1)
    opt = {"sgd": lambda: R.SGD(ref.parameters(), lr), "adam": lambda: R.Adam(ref.parameters(), lr, None, 0.1, wd),
           "adamw": lambda: R.AdamW(ref.parameters(), lr, None, 0.1, wd)}[kind]()
    rng = np.random.default_rng(6)
    for i, (x, y) in enumerate(batches(rng, 6, batch, (784,), ragged=392, ragged_at=3)):
        loss_ref, _ = R.train_step(ref, opt, R.Tensor.new(x, x.shape), R.Tensor.new(y, y.shape))
        loss, _ = tr.step(x, y)
        assert abs(loss - loss_ref) <= 1e-4 * abs(loss_ref), (i, loss, loss_ref)
    assert tr.fused_steps() == 6 and tr.fused_kind() == 2
    for j, p in enumerate(ref.parameters()):
        close(m.get_param(j), p.data(), 3e-4, f"param {j}")


def test_wide_cfg4_adam_default_eps_free_run():
    """configs[3] with its stated optimizer, free running (a wiring check: the parity gate is the three tight tests above —
    gradients to 1e-4 along the oracle's trajectory, the optimizer kernel on identical gradients in test_kernels_gpu.py, the
    eps = 0.1 free run).  With eps = 1e-8 Adam's update is lr * sign-like wherever |g| is within its own summation noise, so
    two trajectories that differ by rounding may move such an element in opposite directions: per step they can part by up
    to 2 * lr there.  Loss to 1e-4 per step, at most 35 % of the elements beyond 1e-4 relative, median within it, nobody
    further than 2 * lr * steps."""
    dims, spec, batch = CFG4
    from taper_b200 import host
    lr, steps = 1e-3, 3
    ref, m = make_pair(lambda r: R.build_mlp(dims, r), spec, 0)
    tr = host.Trainer(m, "adam", lr=lr)
    opt = R.Adam(ref.parameters(), lr)
    rng = np.random.default_rng(1)
    for i, (x, y) in enumerate(batches(rng, steps, batch, (784,))):
        loss_ref, _ = R.train_step(ref, opt, R.Tensor.new(x, x.shape), R.Tensor.new(y, y.shape))
        loss, _ = tr.step(x, y)
        assert abs(loss - loss_ref) <= 1e-4 * abs(loss_ref), (i, loss, loss_ref)
    assert tr.fused_steps() == steps and tr.fused_kind() == 2
    for j, p in enumerate(ref.parameters()):
        got = m.get_param(j).astype(np.float64).reshape(-1)
        want = p.data().astype(np.float64).reshape(-1)
        scale = max(np.max(np.abs(want)), 1e-6)
        err = np.abs(got - want)
        assert np.mean(err > 1e-4 * scale) <= 0.35 and np.median(err) <= 1e-4 * scale, f"param {j}"
        assert err.max() <= 1e-4 * scale + 2.1 * lr * steps, f"param {j}: max abs err {err.max():.3e}"


def test_wide_matches_tape_graph_path():
    """Plan vs one-kernel-per-op path (3xTF32 GEMMs) on the same data from the same parameters: first-step loss to 1e-6,
    parameters after an eps = 0.1 Adam run to 1e-4."""
    from taper_b200 import host
    dims, spec, batch = CFG4
    rng = np.random.default_rng(3)
    data = list(batches(rng, 4, batch, (784,)))
    outs = []
    for fused in (False, True):
        _, m = make_pair(lambda r: R.build_mlp(dims, r), spec, 5)
        tr = host.Trainer(m, "adam", lr=0.05, eps=0.1)
        tr.set_use_fused(fused)
        losses = [tr.step(x, y) for x, y in data]
        outs.append((losses, [m.get_param(i) for i in range(m.num_params())]))
    assert abs(outs[0][0][0][0] - outs[1][0][0][0]) <= 2e-6 * abs(outs[0][0][0][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        close(b, a, 1e-4)


def test_wide_u8_pixels_equal_f32_pixels_bitwise_and_resident_equals_host_fed():
    """u8 / 255 on the device (src/data/mnist.rs:225) is the same f32 the host would have produced; the in-plan gather
    through perm + cursor (src/data/mnist.rs:276-309) sees the same rows as host-side batching."""
    from taper_b200 import host
    dims, spec, batch = ODD
    rng = np.random.default_rng(9)
    n = 2500
    Xu = rng.integers(0, 256, (n, 784)).astype(np.uint8)
    Xf = (Xu.astype(F32) / F32(255.0)).astype(F32)
    Y = rng.integers(0, 10, n).astype(F32)
    perm = rng.permutation(n).astype(np.uint32)
    ms = [make_pair_for(dims, spec, 1)[1] for _ in range(4)]
    ts = [host.Trainer(m, "adam", lr=1e-3) for m in ms]
    ts[2].load_dataset(Xf, Y, perm)
    ts[3].load_dataset_u8(Xu, Y, perm)
    for s in range(4):                                   # wraps around the dataset
        idx = perm[(s * batch + np.arange(batch)) % n]
        r0 = ts[0].step(Xf[idx], Y[idx])
        ts[1].step_async_u8(np.ascontiguousarray(Xu[idx]), np.ascontiguousarray(Y[idx]), pinned=False)
        r1 = ts[1].fetch()
        ts[2].step_resident(batch); r2 = ts[2].fetch()
        ts[3].step_resident(batch); r3 = ts[3].fetch()
        assert r0 == r1 == r2 == r3, f"step {s}: {r0} {r1} {r2} {r3}"
    for i in range(ms[0].num_params()):
        for k in (1, 2, 3):
            np.testing.assert_array_equal(ms[0].get_param(i), ms[k].get_param(i))
    assert all(t.fused_kind() == 2 for t in ts)


def test_wide_plan_follows_parameters_written_behind_its_back():
    """The plan's bf16 operand planes are refreshed by its own optimizer kernel; set_param / load_checkpoint / a step on the
    tape path write the fp32 parameters only, so the next plan step must re-derive the planes."""
    from taper_b200 import host
    dims, spec, batch = ODD
    rng = np.random.default_rng(4)
    x, y = rng.random((batch, 784)).astype(F32), rng.integers(0, 10, batch).astype(F32)
    _, m = make_pair_for(dims, spec, 2)
    tr = host.Trainer(m, "sgd", lr=0.0)                  # lr 0: parameters never move, the loss only depends on them
    l0, _ = tr.step(x, y)
    w = m.get_param(0)
    m.set_param(0, (w * F32(0.5)).astype(F32))
    l1, _ = tr.step(x, y)
    assert abs(l1 - l0) > 1e-3 * abs(l0)                # the halved first layer is what the plan multiplied with
    m.set_param(0, w)
    l2, _ = tr.step(x, y)
    assert l2 == l0


def test_wide_deterministic_run_to_run():
    from taper_b200 import host
    dims, spec, batch = CFG4
    rng = np.random.default_rng(8)
    data = list(batches(rng, 3, batch, (784,)))
    outs = []
    for _ in range(2):
        _, m = make_pair(lambda r: R.build_mlp(dims, r), spec, 5)
        tr = host.Trainer(m, "adam", lr=1e-3)
        outs.append(([tr.step(x, y) for x, y in data], [m.get_param(i) for i in range(m.num_params())]))
    assert outs[0][0] == outs[1][0]
    for a, b in zip(outs[0][1], outs[1][1]):
        np.testing.assert_array_equal(a, b)
