"""The PRODUCT's LR schedulers (taper_b200/csrc/host/nn.cpp, through the C host ABI tp_scheduler_*) against the reference's own
known-answer tests (src/optim.rs:392-422) and, step for step, against the oracle's restatement of src/optim.rs:190-352.
Pure host logic: runs without a GPU."""
import numpy as np
import pytest

from oracle import taper_ref as R


def test_reference_kats_on_the_product_schedulers():        # src/optim.rs:392-422
    from taper_b200 import host
    s = host.Scheduler("step", 0.1, p1=0.5, n=3)
    for _ in range(3):
        s.step()
    assert s.get_lr() == pytest.approx(0.05, abs=1e-6)
    e = host.Scheduler("exponential", 0.1, p1=0.9)
    e.step()
    assert e.get_lr() == pytest.approx(0.09, abs=1e-6)
    p = host.Scheduler("plateau", 0.1, p1=0.5, p2=1e-6, n=2, mode="min")
    for _ in range(3):
        p.step(1.0)
    assert p.get_lr() == pytest.approx(0.05, abs=1e-6)
    c = host.Scheduler("cosine", 0.1, p1=0.0, n=10)
    prev = 0.1
    for _ in range(5):
        c.step()
        assert c.get_lr() < prev
        prev = c.get_lr()


@pytest.mark.parametrize("kind,args,oracle", [
    ("step", dict(p1=0.8, n=5), lambda: R.StepLR(0.01, 5, 0.8)),                      # examples/train_mnist_cnn.rs:132-137
    ("step", dict(p1=0.1, n=1), lambda: R.StepLR(0.01, 1, 0.1)),
    ("exponential", dict(p1=0.95), lambda: R.ExponentialLR(0.01, 0.95)),
    ("cosine", dict(p1=1e-4, n=7), lambda: R.CosineAnnealingLR(0.01, 7, 1e-4)),
    ("plateau", dict(p1=0.5, p2=1e-3, n=2, mode="min"), lambda: R.ReduceLROnPlateau(0.01, 0.5, 2, 1e-3, "min")),
    ("plateau", dict(p1=0.3, p2=1e-6, n=1, mode="max"), lambda: R.ReduceLROnPlateau(0.01, 0.3, 1, None, "max")),
])
def test_product_schedulers_follow_the_oracle_step_for_step(kind, args, oracle):
    from taper_b200 import host
    s, o = host.Scheduler(kind, 0.01, **args), oracle()
    rng = np.random.default_rng(0)
    metrics = np.abs(np.cumsum(rng.standard_normal(40))).astype(np.float32)      # wanders: improvements and plateaus
    for i in range(40):
        m = float(metrics[i])
        s.step(m)
        o.step(np.float32(m))
        assert s.get_lr() == pytest.approx(float(o.get_lr()), rel=2e-6), f"{kind} epoch {i}"


def test_plateau_without_a_metric_does_nothing():           # src/optim.rs:318-320: `if let Some(metric)`
    from taper_b200 import host
    p = host.Scheduler("plateau", 0.1, p1=0.5, p2=1e-6, n=1, mode="min")
    for _ in range(5):
        p.step(None)
    assert p.get_lr() == pytest.approx(0.1)


def test_unknown_scheduler_is_an_error():
    from taper_b200 import host, TaperError
    with pytest.raises(TaperError, match="unknown scheduler"):
        host.Scheduler("linear", 0.1)
