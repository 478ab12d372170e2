"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/*.h declares; host-only entry points behave.  No compute calls (no GPU here)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import taper_ref as R


def test_library_loads_and_exports_every_declared_symbol():
    from taper_b200 import capi
    syms = capi.declared_symbols()
    assert len(syms) >= 80
    for name in syms:
        assert hasattr(capi.lib, name), f"{name} declared in include/ but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = set(syms) - exported
    assert not missing, f"not exported: {sorted(missing)}"


def test_no_torch_or_cxx_types_in_abi():
    from taper_b200 import capi
    allowed = {"int", "float", "size_t", "uint64_t", "uint32_t", "int64_t", "char", "void", "const", "*",
               "tp_ctx", "tp_buf", "tp_graph", "tp_event", "tp_model", "tp_trainer", "tp_conv_desc", "tp_pool_desc", "tp_step", "tp_step_desc", "tp_xchg",
               "tp_dataset", "tp_loader", "tp_scheduler", "tp_tensor", "tp_optimizer"}
    for name, (ret, params) in capi.declared_symbols().items():
        for decl in [ret] + params:
            toks = set(decl.replace("*", " * ").split())
            assert toks <= allowed, f"{name}: non-plain type in signature: {decl!r}"


def test_abi_version_and_error_string():
    from taper_b200 import capi
    assert capi.lib.tp_abi_version() == 1
    # NULL out pointer -> TP_ERR_INVALID with a message, no crash (reference would panic)
    rc = capi.lib.tp_ctx_create(0, None)
    assert rc == 1
    assert b"NULL" in capi.lib.tp_last_error()


def test_adam_step_size_matches_oracle_powi():          # src/optim.rs:88-90
    from taper_b200 import capi
    for t in (1, 2, 3, 10, 100, 1000, 5000):
        adam = R.Adam([], 1e-3)
        adam.t = t
        assert capi.lib.tp_adam_step_size(1e-3, 0.9, 0.999, t) == pytest.approx(float(adam.step_size()), rel=1e-6)
        got = np.float32(capi.lib.tp_adam_step_size(1e-3, 0.9, 0.999, t))
        assert got == adam.step_size(), (t, got, adam.step_size())


def test_conv_and_pool_out_dims():                      # src/tensor.rs:1254-1255, 1406-1407
    from taper_b200 import capi
    d = capi.ConvDesc(4, 1, 28, 28, 32, 3, 3, 1, 1, 1, 1, 1, 1)
    ho, wo = C.c_int(), C.c_int()
    assert capi.lib.tp_conv2d_out_dims(C.byref(d), C.byref(ho), C.byref(wo)) == 0
    assert (ho.value, wo.value) == (28, 28)
    d = capi.ConvDesc(4, 3, 11, 9, 8, 3, 3, 2, 2, 0, 1, 1, 1)
    assert capi.lib.tp_conv2d_out_dims(C.byref(d), C.byref(ho), C.byref(wo)) == 0
    assert (ho.value, wo.value) == ((11 - 3) // 2 + 1, (9 + 2 - 3) // 2 + 1)
    p = capi.PoolDesc(4, 32, 28, 28, 2, 2, 2, 2, 0, 0)
    assert capi.lib.tp_pool_out_dims(C.byref(p), C.byref(ho), C.byref(wo)) == 0
    assert (ho.value, wo.value) == (14, 14)
    bad = capi.ConvDesc(4, 1, 2, 2, 32, 5, 5, 1, 1, 0, 0, 1, 1)
    assert capi.lib.tp_conv2d_out_dims(C.byref(bad), C.byref(ho), C.byref(wo)) == 1


def test_rust_ffi_block_covers_every_declared_symbol():
    """rust/taper-b200-sys/src/lib.rs is generated from the headers (scripts/gen_rust_ffi.py): it must not drift."""
    import os
    import re
    from taper_b200 import capi
    path = os.path.join(capi.ROOT, "rust", "taper-b200-sys", "src", "lib.rs")
    names = set(re.findall(r"pub fn (tp_\w+)\(", open(path).read()))
    assert names == set(capi.declared_symbols()), sorted(names ^ set(capi.declared_symbols()))


def test_step_supported_draws_the_line_without_a_gpu():
    """tp_step_supported / tp_step_kind only inspect the description: chains of Linear(+ReLU) with <= 16 classes; widths % 4 == 0
    below 1.5 GFLOP per step run as the persistent kernel (kind 1), wider steps with a hidden layer and widths % 8 == 0 as the
    plan of tcgen05 kernels (kind 2)."""
    import ctypes as C
    from taper_b200 import capi

    def desc(dims, batch, opt=1):
        d = capi.StepDesc()
        d.n_layers = len(dims) - 1
        off = 0
        for l in range(len(dims) - 1):
            d.dims[l] = dims[l]
            d.relu[l] = 1 if l < len(dims) - 2 else 0
            d.w_off[l] = off; off += (dims[l] * dims[l + 1] + 3) // 4 * 4
            d.b_off[l] = off; off += (dims[l + 1] + 3) // 4 * 4
        d.dims[len(dims) - 1] = dims[-1]
        d.batch, d.optimizer, d.arena_len = batch, opt, off
        return d

    ok = lambda d: capi.lib.tp_step_supported(C.byref(d))
    assert ok(desc([784, 128, 10], 512)) == 1                 # configs[1]
    assert ok(desc([784, 128, 10], 64, 0)) == 1               # configs[0]
    assert ok(desc([784, 128, 64, 10], 256)) == 1             # the reference's example MLP
    assert ok(desc([784, 10], 200)) == 1                      # softmax regression
    kind = lambda d: capi.lib.tp_step_kind(C.byref(d))
    assert ok(desc([784, 1024, 1024, 10], 1024)) == 1         # configs[3]: 9.8 GFLOP per step: the wide plan
    assert kind(desc([784, 1024, 1024, 10], 1024)) == 2
    assert kind(desc([784, 128, 10], 512)) == 1 and kind(desc([784, 10], 200)) == 1
    assert kind(desc([784, 1024, 2048, 10], 1024)) == 0       # classifier input wider than 1024: tape + graph path
    assert kind(desc([780, 1024, 1024, 10], 1024)) == 0       # 780 % 8 != 0: TMA cannot describe the bf16 planes
    assert ok(desc([784, 128, 32], 64)) == 0                  # 32 classes: not a skinny head
    assert ok(desc([30, 10], 64)) == 0                        # width not a multiple of 4
    assert kind(desc([784, 128, 10], 5000)) == 2              # batch beyond the persistent kernel's gather index buffer: the plan
    assert ok(desc([784, 128, 10], 100000)) == 0
    assert capi.lib.tp_step_supported(None) == 0
