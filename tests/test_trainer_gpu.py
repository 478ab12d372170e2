"""Epoch-level trainer calls behind the C host ABI (Trainer::train_epoch / evaluate / fit + Metrics + LR schedulers,
src/train.rs:98-261; DataLoader with its pinned prefetch pipeline, src/data/mnist.rs:276-385) against the oracle's loop on the
same batches, plus the error / capture behaviours the advisor flagged: device error words surface in fetch(), SGD::set_lr
reaches a captured step, Dropout is never captured, the multi-tensor Adam visits every slice exactly once."""
import ctypes as C

import numpy as np
import pytest

from oracle import taper_ref as R
from test_step_gpu import close, make_pair, defaults  # noqa: F401

pytestmark = pytest.mark.gpu
F32 = np.float32

SMALL = ([784, 128, 10], "linear:784:128,relu,linear:128:10")                             # persistent-kernel device tape
WIDE = ([784, 520, 264, 10], "linear:784:520,relu,linear:520:264,relu,linear:264:10")    # wide plan at batch >= 1000
CNN = (None, "conv_relu:1:8:3:1:1,maxpool:2:2,flatten,linear:1568:10")                    # tape + CUDA-graph path


def oracle_epoch(ref, opt, X, Y, order, batch):
    """Trainer::train_epoch (src/train.rs:98-144) on the oracle: (sum of batch losses / num_batches, correct / samples)."""
    n = len(order)
    nb = (n + batch - 1) // batch
    tot_loss, correct = 0.0, 0
    for b in range(nb):
        idx = order[b * batch:(b + 1) * batch]
        x, y = X[idx], Y[idx]
        loss, acc = R.train_step(ref, opt, R.Tensor.new(x, x.shape), R.Tensor.new(y, y.shape))
        tot_loss += loss
        correct += int(F32(acc) * F32(len(idx)))                                  # (acc * batch as f32) as usize, :117 (f32 round trip)
    return tot_loss / nb, correct / n


@pytest.mark.parametrize("dims,spec,n,batch,u8", [
    (SMALL[0], SMALL[1], 1000, 128, False),          # ragged last batch (104), f32 pixels
    (WIDE[0], WIDE[1], 2300, 1000, False),           # wide plan, ragged last batch (300), f32 pixels
    (WIDE[0], WIDE[1], 2300, 1000, True),            # wide plan fed the raw u8 pixels
    (SMALL[0], SMALL[1], 700, 100, True),            # u8 dataset, model without a wide plan: widened on the host
])
def test_train_epoch_and_evaluate_match_the_oracle_loop(dims, spec, n, batch, u8):
    from taper_b200 import host
    rng = np.random.default_rng(5)
    Xu = rng.integers(0, 256, (n, 784)).astype(np.uint8)
    X = (Xu.astype(F32) / F32(255.0)).astype(F32) if u8 else rng.random((n, 784)).astype(F32)
    Y = rng.integers(0, 10, n).astype(F32)
    ref, m = make_pair(lambda r: R.build_mlp(dims, r), spec, 2)
    tr = host.Trainer(m, "sgd", lr=0.05)
    opt = R.SGD(ref.parameters(), 0.05)
    ds = host.Dataset(Xu if u8 else X, Y)
    assert len(ds) == n
    loader = host.Loader(ds, batch, shuffle=False)
    assert loader.num_batches() == (n + batch - 1) // batch
    order = np.arange(n)
    for epoch in range(2):
        loss_ref, acc_ref = oracle_epoch(ref, opt, X, Y, order, batch)
        loss, acc = tr.train_epoch(loader)
        assert loss == pytest.approx(loss_ref, rel=1e-4), f"epoch {epoch}"
        assert abs(acc - acc_ref) <= 3.0 / n, f"epoch {epoch}"                    # near-tied rows may flip
    # free-running SGD: the wide plan's bf16x3 products (~1e-5 of |g|inf per step) and ReLU units within summation noise of 0
    # add up over the epochs' steps, as on the 3xTF32 tape path (test_cfg4_mlp_wide_adam_large_eps_tight): 3e-4 there
    for j, p in enumerate(ref.parameters()):
        close(m.get_param(j), p.data(), 3e-4 if dims is WIDE[0] else 1e-4, f"param {j}")
    # evaluate (src/train.rs:147-172): same sums without the optimizer
    R.Tape.reset()
    tot, correct = 0.0, 0
    nb = (n + batch - 1) // batch
    for b in range(nb):
        x, y = X[b * batch:(b + 1) * batch], Y[b * batch:(b + 1) * batch]
        lg = ref.forward(R.Tensor.new(x, x.shape))
        tot += float(R.cross_entropy_loss(lg, R.Tensor.new(y, y.shape)).data()[0])
        correct += int(F32(R.accuracy(lg, R.Tensor.new(y, y.shape))) * F32(len(y)))
        R.Tape.reset()
    loss, acc = tr.evaluate(loader)
    assert loss == pytest.approx(tot / nb, rel=1e-4) and abs(acc - correct / n) <= 3.0 / n


def test_shuffled_loader_is_a_permutation_and_reproducible():
    from taper_b200 import host
    rng = np.random.default_rng(1)
    n = 1500
    X, Y = rng.random((n, 784)).astype(F32), rng.integers(0, 10, n).astype(F32)
    outs = []
    for _ in range(2):
        _, m = make_pair(lambda r: R.build_mlp(SMALL[0], r), SMALL[1], 2)
        tr = host.Trainer(m, "adam", lr=1e-3)
        loader = host.Loader(host.Dataset(X, Y), 256, shuffle=True, seed=7)
        outs.append(([tr.train_epoch(loader) for _ in range(2)], [m.get_param(i) for i in range(4)]))
    assert outs[0][0] == outs[1][0]
    for a, b in zip(outs[0][1], outs[1][1]):
        np.testing.assert_array_equal(a, b)
    # a different seed visits the samples in another order: another trajectory
    _, m = make_pair(lambda r: R.build_mlp(SMALL[0], r), SMALL[1], 2)
    tr = host.Trainer(m, "adam", lr=1e-3)
    other = tr.train_epoch(host.Loader(host.Dataset(X, Y), 256, shuffle=True, seed=8))
    assert other != outs[0][0][0]


def test_max_batches_stops_the_epoch_early():
    from taper_b200 import host
    rng = np.random.default_rng(1)
    X, Y = rng.random((1000, 784)).astype(F32), rng.integers(0, 10, 1000).astype(F32)
    _, m = make_pair(lambda r: R.build_mlp(SMALL[0], r), SMALL[1], 2)
    tr = host.Trainer(m, "sgd", lr=0.01)
    tr.train_epoch(host.Loader(host.Dataset(X, Y), 100, shuffle=False), max_batches=3)
    assert tr.fused_steps() == 3


def test_fit_runs_the_scheduler_and_fills_the_metrics():
    """Trainer::fit (src/train.rs:175-261) with StepLR(step 2, gamma 0.5): per epoch train_epoch, evaluate,
    scheduler.step(val_loss), optimizer.set_lr (:212-216), Metrics (:10-71) — against the same loop on the oracle."""
    from taper_b200 import host
    rng = np.random.default_rng(3)
    n, nv, batch, epochs = 900, 300, 128, 5
    X, Y = rng.random((n, 784)).astype(F32), rng.integers(0, 10, n).astype(F32)
    Xv, Yv = rng.random((nv, 784)).astype(F32), rng.integers(0, 10, nv).astype(F32)
    ref, m = make_pair(lambda r: R.build_mlp(SMALL[0], r), SMALL[1], 2)
    tr = host.Trainer(m, "sgd", lr=0.08)
    tr.set_scheduler(host.Scheduler("step", 0.08, p1=0.5, n=2))
    tr.fit(host.Loader(host.Dataset(X, Y), batch, shuffle=False), host.Loader(host.Dataset(Xv, Yv), batch, shuffle=False), epochs)
    got = tr.metrics()
    opt = R.SGD(ref.parameters(), 0.08)
    sched = R.StepLR(0.08, 2, 0.5)
    for e in range(epochs):
        tl, ta = oracle_epoch(ref, opt, X, Y, np.arange(n), batch)
        R.Tape.reset()
        vl, nb = 0.0, (nv + batch - 1) // batch
        for b in range(nb):
            x, y = Xv[b * batch:(b + 1) * batch], Yv[b * batch:(b + 1) * batch]
            vl += float(R.cross_entropy_loss(ref.forward(R.Tensor.new(x, x.shape)), R.Tensor.new(y, y.shape)).data()[0])
            R.Tape.reset()
        sched.step(np.float32(vl / nb))
        opt.lr = np.float32(sched.get_lr())
        assert got["train_loss"][e] == pytest.approx(tl, rel=2e-4), f"epoch {e}"
        assert got["val_loss"][e] == pytest.approx(vl / nb, rel=2e-4), f"epoch {e}"
        assert abs(got["train_acc"][e] - ta) <= 3.0 / n
    assert [len(got[k]) for k in ("train_loss", "train_acc", "val_loss", "val_acc", "epoch_times")] == [epochs] * 5
    assert tr.get_lr() == pytest.approx(0.08 * 0.5 ** (epochs // 2), rel=1e-6)
    for j, p in enumerate(ref.parameters()):
        close(m.get_param(j), p.data(), 2e-4, f"param {j}")


@pytest.mark.parametrize("spec,batch,shape,kind", [(SMALL[1], 64, (784,), 1), (WIDE[1], 1000, (784,), 2), (CNN[1], 16, (1, 28, 28), 0)])
def test_label_outside_the_class_range_is_an_error_like_the_reference_panic(spec, batch, shape, kind):
    """`logp[i * c + target]` panics in the reference for target >= classes (src/loss.rs:160-162); here the kernel raises the
    context's sticky error word, the step's result slot carries it and fetch() throws.  All three step paths."""
    from taper_b200 import host, TaperError
    rng = np.random.default_rng(0)
    m = host.Model(spec, 0)
    tr = host.Trainer(m, "adam", lr=1e-3)
    x = rng.random((batch,) + shape).astype(F32)
    y = rng.integers(0, 10, batch).astype(F32)
    tr.step(x, y)
    assert tr.fused_kind() == kind
    bad = y.copy(); bad[3] = 12.0
    with pytest.raises(TaperError, match="label is outside"):
        tr.step(x, bad)
    assert tr.device_error() == 1                                               # reading the word clears it
    assert tr.device_error() == 0
    loss, _ = tr.step(x, y)                                                     # the context trains on after the caller handled it
    assert np.isfinite(loss)


def test_sgd_set_lr_reaches_a_captured_step():
    """SGD's learning rate is read from device memory, so the CUDA graph captured on the second step follows set_lr
    (StepLR via Trainer::fit, src/train.rs:212-216) — bitwise equal to the eager path."""
    from taper_b200 import host
    rng = np.random.default_rng(4)
    data = [(rng.random((64, 784)).astype(F32), rng.integers(0, 10, 64).astype(F32)) for _ in range(7)]
    outs = []
    for use_graph in (False, True):
        _, m = make_pair(lambda r: R.build_mlp(SMALL[0], r), SMALL[1], 2)
        tr = host.Trainer(m, "sgd", lr=0.1)
        tr.set_use_fused(False)
        tr.set_use_graph(use_graph)
        for i, (x, y) in enumerate(data):
            if i == 4:
                tr.set_lr(0.01)
            tr.step(x, y)
        assert tr.graph_replays() == (6 if use_graph else 0)          # step 0 eager, step 1 captured + launched, 5 replays
        outs.append([m.get_param(i) for i in range(4)])
    for a, b in zip(*outs):
        np.testing.assert_array_equal(a, b)


def test_dropout_model_trains_without_graph_capture():
    """Dropout draws its mask on the host every forward (src/nn.rs:799-822): such a step is never captured (a replay would
    re-use whatever the staging ring held)."""
    from taper_b200 import host
    rng = np.random.default_rng(0)
    m = host.Model("linear:784:64,relu,dropout:50,linear:64:10", 0)
    tr = host.Trainer(m, "adam", lr=1e-3)
    x, y = rng.random((32, 784)).astype(F32), rng.integers(0, 10, 32).astype(F32)
    losses = [tr.step(x, y)[0] for _ in range(6)]
    assert tr.graph_replays() == 0 and tr.fused_steps() == 0
    assert all(np.isfinite(l) for l in losses)
    assert len({round(l, 6) for l in losses}) > 1                                # a fresh mask every step


def test_adam_segments_visits_every_slice_exactly_once(ctx):
    """More than 32 slices with gradient-less ones in between (mode 0 / 2): one Adam step per mode-1 slice, one decay per
    mode-2 slice, nothing twice (the scan position, not a fixed stride of 32, carries over between launches)."""
    from taper_b200 import capi
    lib = capi.lib
    rng = np.random.default_rng(0)
    nseg, seg = 75, 12
    total = nseg * seg
    p0 = rng.standard_normal(total).astype(F32)
    g = (rng.standard_normal(total) * 0.1).astype(F32)
    modes = np.array([(1 if i % 3 else 0) if i % 7 else 2 for i in range(nseg)], np.int32)
    offs = (np.arange(nseg) * seg).astype(np.int64)
    lens = np.full(nseg, seg, np.int64)
    P, G, M, V, H = ctx.upload(p0), ctx.upload(g), ctx.zeros(total), ctx.zeros(total), ctx.alloc(8)
    capi.check(lib.tp_adam_hyper_init(ctx.h, H.h, 1e-2, 0.9, 0.999, 1e-8, 0.1))
    capi.check(lib.tp_adam_advance(ctx.h, H.h))
    capi.check(lib.tp_adam_step_segments(ctx.h, P.h, G.h, M.h, V.h, H.h, 1.0, 1, offs.ctypes.data_as(C.POINTER(C.c_int64)),
                                         lens.ctypes.data_as(C.POINTER(C.c_int64)), modes.ctypes.data_as(C.POINTER(C.c_int)), nseg))
    got = P.download()
    # reference: the same decoupled step on one flat arena, then masked per slice
    P2, M2, V2 = ctx.upload(p0), ctx.zeros(total), ctx.zeros(total)
    capi.check(lib.tp_adam_step_dev(ctx.h, P2.h, G.h, M2.h, V2.h, H.h, 1.0, 1, total))
    stepped = P2.download()
    decay = np.float32(1.0) - np.float32(1e-2) * np.float32(0.1)
    for i in range(nseg):
        s = slice(i * seg, (i + 1) * seg)
        want = stepped[s] if modes[i] == 1 else (p0[s] * decay if modes[i] == 2 else p0[s])
        np.testing.assert_array_equal(got[s], want.astype(F32), err_msg=f"slice {i} mode {modes[i]}")


@pytest.fixture()
def ctx():
    import taper_b200
    c = taper_b200.Ctx(0)
    yield c
    c.close()


def test_load_dataset_rejects_a_permutation_that_does_not_cover_the_dataset():
    """tp_trainer_load_dataset[_u8] reads one permutation entry per sample (include/taper_b200_host.h): a shorter array — e.g. a
    data-parallel shard's order handed over as it is — would be read out of bounds, so the Python layer refuses it, and a
    permutation padded to n entries by repeating the shard's order walks the shard's rows."""
    from taper_b200 import host
    from taper_b200.dp import shard_permutation
    dims, spec = SMALL
    rng = np.random.default_rng(3)
    n, b = 512, 64
    X = rng.random((n, 784)).astype(F32)
    Xu = rng.integers(0, 256, (n, 784)).astype(np.uint8)
    Y = rng.integers(0, 10, n).astype(F32)
    perm = rng.permutation(n)
    shard = shard_permutation(perm, 1, 2, b)
    assert shard.size == n // 2
    ms = [host.Model(spec, 0), host.Model(spec, 0)]
    ts = [host.Trainer(m, "sgd", lr=0.05) for m in ms]
    with pytest.raises(ValueError):
        ts[0].load_dataset(X, Y, shard)
    with pytest.raises(ValueError):
        ts[0].load_dataset_u8(Xu, Y, shard)
    ts[0].load_dataset(X, Y, np.resize(shard, n))
    for s in range(3):
        idx = shard[s * b:(s + 1) * b]
        ts[0].step_resident(b)
        r0 = ts[0].fetch()
        r1 = ts[1].step(X[idx], Y[idx])
        assert r0[0] == pytest.approx(r1[0], rel=1e-6) and r0[1] == r1[1], (s, r0, r1)
