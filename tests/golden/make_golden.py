"""Generates the golden fixtures under tests/golden/ with the CPU oracle (oracle/taper_ref.py).

The reference is a Rust crate and no Rust toolchain exists in the build image, so these vectors come from the oracle — the
op-for-op restatement of the reference, itself pinned against the reference's own known-answer tests (reference_kats.json,
tests/test_oracle_kats.py).  Re-run with `python tests/golden/make_golden.py`; test_golden.py checks that the oracle still
reproduces the committed files (drift guard) and that the CUDA path reproduces them through the C ABI.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import taper_ref as R                       # noqa: E402

F32 = np.float32


def mlp_steps(dims, batch, kind, steps, seed, ragged=None):
    """Free-running training steps of an MLP; eps = 0.1 keeps Adam's update linear in the gradient (tight comparison)."""
    rng = np.random.default_rng(seed)
    R.Tape.reset()
    model = R.build_mlp(dims, rng)
    for p in model.parameters():
        if len(p.shape) == 1:
            p._data[:] = (rng.standard_normal(p._data.size) * 0.05).astype(F32)
    out = {"dims": np.array(dims), "kind": kind}
    for i, p in enumerate(model.parameters()):
        out[f"init_{i}"] = p.data().copy()
    params = model.parameters()
    opt = {"sgd": lambda: R.SGD(params, 0.05), "adam": lambda: R.Adam(params, 0.05, None, 0.1, 1e-3),
           "adamw": lambda: R.AdamW(params, 0.05, None, 0.1, 1e-2)}[kind]()
    losses, accs = [], []
    for s in range(steps):
        b = ragged if (ragged and s == steps - 1) else batch
        x = rng.random((b, dims[0])).astype(F32)
        y = rng.integers(0, dims[-1], b).astype(F32)
        out[f"x_{s}"], out[f"y_{s}"] = x, y
        loss, acc = R.train_step(model, opt, R.Tensor.new(x, x.shape), R.Tensor.new(y, y.shape))
        losses.append(loss); accs.append(acc * b)
    out["loss"] = np.array(losses, F32)
    out["correct"] = np.array(accs, F32)
    for i, p in enumerate(model.parameters()):
        out[f"final_{i}"] = p.data().copy()
    return out


def conv_pool(seed):
    """conv2d (A2 weight layout) + bias + ReLU, max-pool (first max, flat argmax), global average pool: forward values and,
    for the bias, the strict-reference backward (A1)."""
    rng = np.random.default_rng(seed)
    R.Tape.reset()
    R.Config.strict_reference_conv = True
    x = rng.standard_normal((2, 3, 6, 6)).astype(F32)
    w = (rng.standard_normal((4, 3, 3, 3)) * 0.3).astype(F32)
    b = (rng.standard_normal(4) * 0.1).astype(F32)
    X = R.Tensor.new(x, x.shape)
    W = R.Tensor.new(w, w.shape).requires_grad_()
    Bv = R.Tensor.new(b, b.shape).requires_grad_()
    y = X.conv2d_relu(W, Bv, (1, 1), (1, 1), (1, 1))
    mp = y.max_pool2d((2, 2), (2, 2))
    gap = mp.avg_pool2d((3, 3), (3, 3))
    loss = gap.sum()
    loss.backward()
    return {"x": x, "w": w, "b": b, "conv_relu": y.data().copy(), "maxpool": mp.data().copy(), "gap": gap.data().copy(),
            "grad_b": Bv.grad().copy(), "grad_w_is_none": np.array([W.grad() is None])}


def conv_igemm(seed):
    """A 3x3 convolution wide enough (C_in = C_out = 32, K = 288) to take the implicit-GEMM tensor-core kernel, with bias and
    ReLU, followed by the 2x2 max-pool: forward values (the A2 weight layout is what a wrong kernel gets wrong first)."""
    rng = np.random.default_rng(seed)
    R.Tape.reset()
    R.Config.strict_reference_conv = True
    x = rng.standard_normal((2, 32, 8, 8)).astype(F32)
    w = (rng.standard_normal((32, 32, 3, 3)) * 0.08).astype(F32)
    b = (rng.standard_normal(32) * 0.1).astype(F32)
    X = R.Tensor.new(x, x.shape)
    y = X.conv2d_relu(R.Tensor.new(w, w.shape), R.Tensor.new(b, b.shape), (1, 1), (1, 1), (1, 1))
    mp = y.max_pool2d((2, 2), (2, 2))
    return {"x": x, "w": w, "b": b, "conv_relu": y.data().copy(), "maxpool": mp.data().copy()}


def main():
    np.savez_compressed(os.path.join(HERE, "mlp_sgd.npz"), **mlp_steps([20, 12, 5], 8, "sgd", 4, 11, ragged=5))
    np.savez_compressed(os.path.join(HERE, "mlp_adam.npz"), **mlp_steps([24, 16, 8, 6], 16, "adam", 5, 12, ragged=7))
    np.savez_compressed(os.path.join(HERE, "mlp_adamw.npz"), **mlp_steps([784, 128, 10], 32, "adamw", 3, 13))
    np.savez_compressed(os.path.join(HERE, "conv_pool.npz"), **conv_pool(14))
    np.savez_compressed(os.path.join(HERE, "conv_igemm.npz"), **conv_igemm(15))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
