"""The tcgen05 TF32 GEMM path (taper_b200/csrc/gemm_tc.cu): proves the tensor-core kernel is the one that
runs (1xTF32 results carry TF32-sized rounding, 3xTF32 results are fp32-accurate), all four operand-major
combinations, ragged M/N/K tails served by TMA zero fill, split-K determinism, fused epilogues."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F32 = np.float32


def rel_err(got, ref):
    ref = np.asarray(ref, np.float64).reshape(-1)
    return np.max(np.abs(np.asarray(got, np.float64).reshape(-1) - ref)) / max(np.max(np.abs(ref)), 1e-30)


def run(ctx, mode, ta, tb, m, n, k, A, B, C0=None, alpha=1.0, beta=0.0):
    from taper_b200 import capi
    capi.check(capi.lib.tp_set_gemm_mode(ctx.h, mode))
    try:
        a = ctx.upload((A.T if ta else A).copy())
        b = ctx.upload((B.T if tb else B).copy())
        c = ctx.upload(C0 if C0 is not None else np.zeros((m, n), F32))
        ctx.call("sgemm_rowmajor", ta, tb, m, n, k, alpha, a, b, beta, c)
        return c.download().reshape(m, n)
    finally:
        capi.check(capi.lib.tp_set_gemm_mode(ctx.h, 1))


SHAPES = [(128, 128, 32), (128, 128, 64), (128, 64, 256), (256, 32, 128), (512, 128, 784), (128, 784, 512), (512, 128, 128),
          (1024, 1024, 784), (1024, 1024, 1024), (200, 136, 100), (132, 36, 52), (50176, 64, 288), (288, 64, 12544), (1024, 16, 1024)]


@pytest.mark.parametrize("m,n,k", SHAPES)
@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
def test_tf32_modes_all_majors(ctx, m, n, k, ta, tb):
    rng = np.random.default_rng(m + 7 * n + 13 * k + ta * 2 + tb)
    A = rng.standard_normal((m, k)).astype(F32)
    B = rng.standard_normal((k, n)).astype(F32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    e3 = rel_err(run(ctx, 1, ta, tb, m, n, k, A, B), ref)
    e1 = rel_err(run(ctx, 2, ta, tb, m, n, k, A, B), ref)
    e0 = rel_err(run(ctx, 0, ta, tb, m, n, k, A, B), ref)
    assert e0 < 2e-6, f"fp32 FFMA path: {e0:.2e}"
    assert e3 < 5e-6, f"3xTF32 must be fp32-accurate: {e3:.2e}"
    assert e1 < 2e-3, f"1xTF32: {e1:.2e}"
    # operands whose contiguous dimension is a multiple of 4 floats go through TMA + tcgen05: TF32 rounding must be visible
    a_cols, b_cols = (m if ta else k), (k if tb else n)
    if a_cols % 4 == 0 and b_cols % 4 == 0:
        assert e1 > 2e-5, f"1xTF32 error {e1:.2e} is fp32-like: the tensor-core path did not run"


def test_tf32_hi_lo_split_exactness(ctx):
    """Inputs that are exactly representable in TF32 give bit-identical results in every mode (products are exact
    in fp32 for small integers), including alpha/beta and the accumulate (beta = 1) form of the backward GEMMs."""
    rng = np.random.default_rng(0)
    m, n, k = 256, 128, 96
    A = rng.integers(-8, 9, (m, k)).astype(F32)
    B = rng.integers(-8, 9, (k, n)).astype(F32)
    C0 = rng.integers(-8, 9, (m, n)).astype(F32)
    ref = A @ B
    for mode in (0, 1, 2):
        for ta, tb in ((0, 0), (0, 1), (1, 0), (1, 1)):
            np.testing.assert_array_equal(run(ctx, mode, ta, tb, m, n, k, A, B), ref)
            np.testing.assert_array_equal(run(ctx, mode, ta, tb, m, n, k, A, B, C0, 1.0, 1.0), ref + C0)
            np.testing.assert_array_equal(run(ctx, mode, ta, tb, m, n, k, A, B, C0, 0.5, -2.0), 0.5 * ref - 2.0 * C0)


def test_split_k_is_deterministic(ctx):
    rng = np.random.default_rng(1)
    m, n, k = 128, 128, 4096                      # one output tile, split 64 ways
    A = rng.standard_normal((m, k)).astype(F32)
    B = rng.standard_normal((k, n)).astype(F32)
    first = run(ctx, 1, 1, 0, m, n, k, A, B)
    for _ in range(5):
        np.testing.assert_array_equal(run(ctx, 1, 1, 0, m, n, k, A, B), first)


@pytest.mark.parametrize("batch,fin,fout", [(512, 784, 128), (1024, 784, 1024), (96, 784, 128), (1024, 1024, 1024)])
def test_linear_epilogues_on_tensor_path(ctx, batch, fin, fout):
    rng = np.random.default_rng(batch + fin)
    x = rng.standard_normal((batch, fin)).astype(F32)
    w = (rng.standard_normal((fout, fin)) * 0.05).astype(F32)
    b = rng.standard_normal(fout).astype(F32)
    gy = rng.standard_normal((batch, fout)).astype(F32)
    X, W, Bb, G = ctx.upload(x), ctx.upload(w), ctx.upload(b), ctx.upload(gy)
    Y = ctx.alloc(batch * fout)
    z = x.astype(np.float64) @ w.astype(np.float64).T + b
    ctx.call("linear_fwd", X, W, Bb, Y, batch, fin, fout, 1)
    y = Y.download().reshape(batch, fout)
    assert rel_err(y, np.maximum(z, 0)) < 5e-6
    gz = gy * (y > 0)
    gx, gw, gb = ctx.alloc(batch * fin), ctx.alloc(fout * fin), ctx.alloc(fout)
    ctx.call("linear_bwd", X, W, G, Y, gx, gw, gb, batch, fin, fout, 0, 0, 0)
    assert rel_err(gx.download(), gz.astype(np.float64) @ w.astype(np.float64)) < 5e-6
    assert rel_err(gw.download(), gz.astype(np.float64).T @ x.astype(np.float64)) < 5e-6
    assert rel_err(gb.download(), gz.sum(axis=0, dtype=np.float64)) < 5e-6


@pytest.mark.parametrize("m,n,k,ta,tb", [(512, 128, 784, 0, 1), (128, 784, 512, 1, 0), (288, 64, 12544, 1, 0), (2560, 1024, 8192, 0, 0)])
def test_3xtf32_has_no_truncation_bias_on_nonnegative_data(ctx, m, n, k, ta, tb):
    """MNIST pixels and post-ReLU activations are non-negative.  The tensor core's own accumulator truncates, which would
    bias a long same-signed sum by ~K/3 * 2^-24; the 3xTF32 mode drains the accumulator every 128 elements of K."""
    rng = np.random.default_rng(k)
    A = rng.random((m, k)).astype(F32)
    B = rng.random((k, n)).astype(F32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    got = run(ctx, 1, ta, tb, m, n, k, A, B).astype(np.float64)
    assert rel_err(got, ref) < 5e-6
    assert abs(np.mean(got - ref)) / np.max(np.abs(ref)) < 4e-6      # bounded by the 128-element chunk: ~128/3 * 2^-24
