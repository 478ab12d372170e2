"""CPU oracle: an op-for-op NumPy restatement of taper's tape-evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``taper_b200/`` imports this module; it is used by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs as the checker / timed CPU baseline, never as the product path.

Parity pin status
-----------------
The reference is a Rust crate and no Rust toolchain exists in the build image, so the reference
itself could not be executed.  The oracle is pinned against every known-answer test the reference's
own tests hold for this path (``tests/test_oracle_kats.py`` cites each
``tests/smoke.rs`` / ``src/loss.rs`` / ``src/optim.rs`` line) and, for everything that is standard
math, against PyTorch autograd (the ``torch`` cross-checks in ``tests/test_oracle_kats.py``; the ``tests/golden/*.npz``
vectors are written by ``tests/golden/make_golden.py`` from this oracle: drift guards, not pins).  Conv2d / pooling / optimizer
*values* are not pinned by any reference test ("parity unpinned" there): the
restatement following the cited lines is the only pin.

Every function cites the reference file:line it follows (paths relative to the reference root).
All arithmetic is float32.  GEMM goes through NumPy's BLAS ``sgemm`` (OpenBLAS), standing in for the
reference's ``--features blas-openblas`` path (src/gemm.rs:8-49); summation order inside sgemm is
implementation-defined in the reference as well (matrixmultiply 0.3.10 or a vendor BLAS).
"""
from __future__ import annotations

import math
import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------------------------
# Switches that select between "what the reference does" and "the evident intent".
# --------------------------------------------------------------------------------------------
class Config:
    # A1 (src/tensor.rs:1725, 2075): im2col and transpose_4d return plain tensors without tape
    # nodes, so no gradient reaches conv weights or crosses a conv layer.  True = reproduce that.
    strict_reference_conv = True
    # A3 (src/tensor.rs:524-528, src/tape.rs:65): node id 0 doubles as the "no node" sentinel, so
    # a graph whose root is the first recorded op never backprops.  True = reproduce that.
    node0_sentinel = False


# --------------------------------------------------------------------------------------------
# Tape  (src/tape.rs:6-127)
# --------------------------------------------------------------------------------------------
class Tape:
    """Thread-local Vec of closures in the reference (src/tape.rs:6-23); one global list here."""
    nodes: list = []

    @staticmethod
    def reset():                      # src/tape.rs:43-49
        Tape.nodes = []

    @staticmethod
    def new():                        # src/tape.rs:27-31
        return Tape

    @staticmethod
    def ensure_active():              # src/tape.rs:34-40
        pass

    @staticmethod
    def _push(output, fn):
        idx = len(Tape.nodes)
        Tape.nodes.append(fn)
        # reference stamps the raw index (src/tape.rs:63-74); the A3 fix stamps index+1
        output._node[0] = idx if Config.node0_sentinel else idx + 1

    @staticmethod
    def push_binary_op(a, b, output, fn):   # src/tape.rs:51-76
        if not (a.requires_grad or b.requires_grad):
            return
        Tape._push(output, fn)

    @staticmethod
    def push_unary_op(inp, output, fn):     # src/tape.rs:78-101
        if not inp.requires_grad:
            return
        Tape._push(output, fn)


def tape_backward(final_node_id):           # src/tape.rs:106-127
    if not Tape.nodes:
        return
    end = min(final_node_id, len(Tape.nodes) - 1)
    fns = list(Tape.nodes[: end + 1])        # cloned first: closures may append nodes (A5)
    for f in reversed(fns):
        f()


# --------------------------------------------------------------------------------------------
# Tensor  (src/tensor.rs:236-244, 469-541)
# --------------------------------------------------------------------------------------------
class Tensor:
    __slots__ = ("_data", "shape", "_grad", "requires_grad", "_node")

    def __init__(self, data, shape):
        arr = np.ascontiguousarray(np.asarray(data, dtype=F32).reshape(-1))
        shape = tuple(int(s) for s in shape)
        assert arr.size == int(np.prod(shape, dtype=np.int64)), (arr.size, shape)
        self._data = arr
        self.shape = shape
        self._grad = [None]           # Arc<RwLock<Option<Vec<f32>>>>
        self.requires_grad = False
        self._node = [0]              # Arc<AtomicUsize>

    # constructors -------------------------------------------------------------------------
    @staticmethod
    def new(data, shape):             # src/tensor.rs:470-478
        return Tensor(data, shape)

    @staticmethod
    def scalar(v):                    # src/tensor.rs:480-482
        return Tensor([v], (1,))

    def requires_grad_(self):         # src/tensor.rs:484-487 (`requires_grad(self) -> Self`)
        self.requires_grad = True
        return self

    # accessors ----------------------------------------------------------------------------
    def data(self):                   # src/tensor.rs:493-496
        return self._data

    def numpy(self):
        return self._data.reshape(self.shape)

    def grad(self):                   # src/tensor.rs:512-518 (clone of the grad or None)
        g = self._grad[0]
        return None if g is None else g.copy()

    def grad_ref(self):               # src/tensor.rs:505-508
        return self._grad[0]

    def set_grad(self, g):            # `grad` is a pub field (src/tensor.rs:241)
        self._grad[0] = None if g is None else np.asarray(g, dtype=F32).reshape(-1).copy()

    def _grad_slot(self):
        if self._grad[0] is None:     # lazily zero-allocated (e.g. src/ops.rs:126-128)
            self._grad[0] = np.zeros(self._data.size, dtype=F32)
        return self._grad[0]

    def zero_grad(self):              # src/tensor.rs:531-533
        self._grad[0] = None

    def backward(self):               # src/tensor.rs:520-529
        self._grad[0] = np.ones(self._data.size, dtype=F32)
        node_id = self._node[0]
        if node_id != 0:
            tape_backward(node_id if Config.node0_sentinel else node_id - 1)

    # ---- elementwise ops  (src/ops.rs:8-120, 377-496) --------------------------------------
    def __add__(self, other):         # src/ops.rs:8-52
        assert self._data.size == other._data.size, "Tensor dimensions must match"
        out = Tensor(self._data + other._data, self.shape)
        if self.requires_grad or other.requires_grad:
            out.requires_grad = True
            a, b, o = self, other, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    if a.requires_grad:
                        accumulate_grad(a, g)
                    if b.requires_grad:
                        accumulate_grad(b, g)
            Tape.push_binary_op(self, other, out, bw)
        return out

    def __sub__(self, other):         # src/ops.rs:377-420
        assert self._data.size == other._data.size, "Tensor dimensions must match"
        out = Tensor(self._data - other._data, self.shape)
        if self.requires_grad or other.requires_grad:
            out.requires_grad = True
            a, b, o = self, other, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    if a.requires_grad:
                        accumulate_grad(a, g)
                    if b.requires_grad:
                        accumulate_grad_scaled(b, g, -1.0)
            Tape.push_binary_op(self, other, out, bw)
        return out

    def __mul__(self, other):         # src/ops.rs:54-120
        assert self._data.size == other._data.size, "Tensor dimensions must match"
        out = Tensor(self._data * other._data, self.shape)
        if self.requires_grad or other.requires_grad:
            out.requires_grad = True
            a, b, o = self, other, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    if a.requires_grad:
                        ga = a._grad_slot()
                        a._grad[0] = ga + g * b._data
                    if b.requires_grad:
                        gb = b._grad_slot()
                        b._grad[0] = gb + g * a._data
            Tape.push_binary_op(self, other, out, bw)
        return out

    def __truediv__(self, other):     # src/ops.rs:440-496
        assert self._data.size == other._data.size, "Tensor dimensions must match"
        out = Tensor(self._data / other._data, self.shape)
        if self.requires_grad or other.requires_grad:
            out.requires_grad = True
            a, b, o = self, other, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    if a.requires_grad:
                        ga = a._grad_slot()
                        ga += g / b._data
                    if b.requires_grad:
                        gb = b._grad_slot()
                        gb -= g * a._data / (b._data * b._data)
            Tape.push_binary_op(self, other, out, bw)
        return out

    # ---- matmul  (src/ops.rs:200-298 over src/gemm.rs) --------------------------------------
    def matmul(self, other):
        assert len(self.shape) == 2 and len(other.shape) == 2
        m, k = self.shape
        k2, n = other.shape
        assert k == k2, "Inner dimensions must match"
        c = np.zeros(m * n, dtype=F32)
        sgemm_rowmajor(False, False, m, n, k, 1.0, self._data, other._data, 0.0, c)
        out = Tensor(c, (m, n))
        if self.requires_grad or other.requires_grad:
            out.requires_grad = True
            a, b, o = self, other, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    if a.requires_grad:          # dA += dC·Bᵀ  (N,T,β=1)  src/ops.rs:241-266
                        ga = a._grad_slot()
                        sgemm_rowmajor(False, True, m, k, n, 1.0, g, b._data, 1.0, ga)
                    if b.requires_grad:          # dB += Aᵀ·dC  (T,N,β=1)  src/ops.rs:268-292
                        gb = b._grad_slot()
                        sgemm_rowmajor(True, False, k, n, m, 1.0, a._data, g, 1.0, gb)
            Tape.push_binary_op(self, other, out, bw)
        return out

    # ---- relu  (src/ops.rs:312-374) ----------------------------------------------------------
    def relu(self):
        out = Tensor(np.maximum(self._data, F32(0.0)), self.shape)
        if self.requires_grad:
            out.requires_grad = True
            x, o = self, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    gin = x._grad_slot()
                    gin += np.where(x._data > 0, g, F32(0.0))
            Tape.push_unary_op(self, out, bw)
        return out

    # ---- transpose  (src/tensor.rs:544-591) ---------------------------------------------------
    def transpose(self):
        assert len(self.shape) == 2, "Can only transpose 2D tensors"
        r, c = self.shape
        out = Tensor(np.ascontiguousarray(self._data.reshape(r, c).T), (c, r))
        if self.requires_grad:
            out.requires_grad = True
            x, o = self, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    gin = x._grad_slot()
                    gin += np.ascontiguousarray(g.reshape(c, r).T).reshape(-1)
            Tape.push_unary_op(self, out, bw)
        return out

    # ---- add_broadcast  (src/tensor.rs:636-704) -----------------------------------------------
    def add_broadcast(self, other):
        if self.shape == other.shape:
            return self + other
        assert len(self.shape) == 2 and len(other.shape) == 1, "Unsupported broadcasting shapes"
        bsz, f = self.shape
        assert f == other.shape[0], "Last dimension must match for broadcasting"
        out = Tensor((self._data.reshape(bsz, f) + other._data[None, :]), self.shape)
        if self.requires_grad or other.requires_grad:
            out.requires_grad = True
            a, b, o = self, other, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    if a.requires_grad:
                        accumulate_grad(a, g)
                    if b.requires_grad:
                        gb = b._grad_slot()
                        gb += g.reshape(bsz, f).sum(axis=0, dtype=F32)
            Tape.push_binary_op(self, other, out, bw)
        return out

    # ---- sub_broadcast_rows  (src/tensor.rs:707-770) -------------------------------------------
    def sub_broadcast_rows(self, other):
        if self.shape == other.shape:
            return self - other
        assert len(self.shape) == 2 and other.shape == (self.shape[0], 1)
        bsz, c = self.shape
        out = Tensor(self._data.reshape(bsz, c) - other._data[:, None], self.shape)
        if self.requires_grad or other.requires_grad:
            out.requires_grad = True
            a, r, o = self, other, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    if a.requires_grad:
                        accumulate_grad(a, g)
                    if r.requires_grad:
                        accumulate_grad(r, -g.reshape(bsz, c).sum(axis=1, dtype=F32))
            Tape.push_binary_op(self, other, out, bw)
        return out

    # ---- reshape / flatten / view  (src/tensor.rs:803-858, 1214) --------------------------------
    def reshape(self, shape):
        shape = tuple(int(s) for s in shape)
        assert self._data.size == int(np.prod(shape, dtype=np.int64)), "Cannot reshape"
        out = Tensor(self._data.copy(), shape)          # copies (A11)
        if self.requires_grad:
            out.requires_grad = True
            x, o = self, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    gin = x._grad_slot()
                    gin += g
            Tape.push_unary_op(self, out, bw)
        return out

    def flatten(self, start_dim=1):
        assert start_dim < len(self.shape), "start_dim out of bounds"
        new_shape = list(self.shape[:start_dim]) + [int(np.prod(self.shape[start_dim:]))]
        return self.reshape(new_shape)

    view = reshape

    # ---- sum  (src/tensor.rs:890-1018) ------------------------------------------------------------
    def sum(self, dim=None, keepdim=False):
        if dim is None:
            out = Tensor.scalar(self._data.sum(dtype=F32))
            if self.requires_grad:
                out.requires_grad = True
                x, o = self, out

                def bw():
                    g = o._grad[0]
                    if g is not None:
                        accumulate_grad(x, np.full(x._data.size, g[0], dtype=F32))
                Tape.push_unary_op(self, out, bw)
            return out
        d = int(dim)
        assert d < len(self.shape)
        arr = self._data.reshape(self.shape)
        res = arr.sum(axis=d, keepdims=keepdim, dtype=F32)
        out = Tensor(res, res.shape if res.ndim else (1,))
        if self.requires_grad:
            out.requires_grad = True
            x, o = self, out
            kd_shape = list(self.shape)
            kd_shape[d] = 1

            def bw():
                g = o._grad[0]
                if g is not None:
                    gin = x._grad_slot()
                    gin += np.broadcast_to(g.reshape(kd_shape), x.shape).reshape(-1)
            Tape.push_unary_op(self, out, bw)
        return out

    # ---- max / argmax  (src/tensor.rs:1021-1088) ----------------------------------------------------
    def max(self, dim=None):
        if dim is None:
            # Iterator::max_by keeps the LAST of equal maxima (src/tensor.rs:1072-1080)
            d = self._data
            idx = d.size - 1 - int(np.argmax(d[::-1])) if d.size else 0
            val = d[idx] if d.size else 0.0
            return Tensor.scalar(val), Tensor.scalar(float(idx))
        d = int(dim)
        assert d < len(self.shape) and len(self.shape) <= 2, "A7: only 1-D/2-D is well defined"
        arr = self._data.reshape(self.shape)
        out_shape = list(self.shape)
        out_shape[d] = 1
        # strict '>' scanning upward from -inf / index 0: first max wins, NaN never selected
        with np.errstate(invalid="ignore"):
            cmp = np.where(np.isnan(arr), -np.inf, arr)
        idx = np.argmax(cmp, axis=d)
        vals = np.take_along_axis(cmp, np.expand_dims(idx, d), axis=d)
        # a row of all -inf/NaN keeps (-inf, 0)
        return Tensor(vals.astype(F32), out_shape), Tensor(idx.astype(F32), out_shape)

    def argmax(self, dim=None):
        return self.max(dim)[1]

    # ---- sigmoid / mean / pow / sqrt  (src/tensor.rs:594-634, 772-800, 1172-1211) ------------
    def sigmoid(self):
        x = self._data
        with np.errstate(over="ignore"):
            pos = F32(1.0) / (F32(1.0) + np.exp(-x))                  # x > 0 branch (:602-604)
            ex = np.exp(x)
            neg = ex / (F32(1.0) + ex)                                # else branch (:605-607)
        res = np.where(x > 0, pos, neg).astype(F32)
        out = Tensor(res, self.shape)
        if self.requires_grad:
            out.requires_grad = True
            xt, o = self, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    gin = xt._grad_slot()
                    gin += g * res * (F32(1.0) - res)                 # :627
            Tape.push_unary_op(self, out, bw)
        return out

    def mean(self):
        n = self._data.size
        acc = F32(0.0)
        for v in self._data:                                          # iter().sum::<f32>() (:774)
            acc = F32(acc + v)
        out = Tensor.scalar(F32(acc / F32(n)))
        if self.requires_grad:
            out.requires_grad = True
            xt, o = self, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    gin = xt._grad_slot()
                    gin += F32(g[0] / F32(n))                         # :786-795
            Tape.push_unary_op(self, out, bw)
        return out

    def pow(self, exponent):
        e = F32(exponent)
        with np.errstate(invalid="ignore", divide="ignore"):
            res = np.power(self._data, e).astype(F32)                 # powf (:1177)
        out = Tensor(res, self.shape)
        if self.requires_grad:
            out.requires_grad = True
            xt, o = self, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    gin = xt._grad_slot()
                    with np.errstate(invalid="ignore", divide="ignore"):
                        gin += g * e * np.power(xt._data, F32(e - F32(1.0))).astype(F32)      # :1199
            Tape.push_unary_op(self, out, bw)
        return out

    def sqrt(self):                                                   # :1209-1211
        return self.pow(0.5)

    # ---- exp / log  (src/tensor.rs:1091-1169) ---------------------------------------------------------
    def exp(self):
        res = np.exp(self._data)
        out = Tensor(res, self.shape)
        if self.requires_grad:
            out.requires_grad = True
            x, o = self, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    gin = x._grad_slot()
                    x._grad[0] = gin + g * res
            Tape.push_unary_op(self, out, bw)
        return out

    def log(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            out = Tensor(np.log(self._data), self.shape)
        if self.requires_grad:
            out.requires_grad = True
            x, o = self, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    gin = x._grad_slot()
                    gin += g / x._data
            Tape.push_unary_op(self, out, bw)
        return out

    # ---- conv2d  (src/tensor.rs:1221-1285, 1663-1780, 1972-2076) ----------------------------------------
    def conv2d(self, weight, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1)):
        assert len(self.shape) == 4 and len(weight.shape) == 4
        n, cin, h, w = self.shape
        cout, cin_w, kh, kw = weight.shape
        assert cin == cin_w, "Input and weight channel dimensions must match"
        sh, sw = stride
        ph, pw = padding
        dh, dw = dilation
        ho = (h + 2 * ph - dh * (kh - 1) - 1) // sh + 1
        wo = (w + 2 * pw - dw * (kw - 1) - 1) // sw + 1
        k = cin * kh * kw
        col = self._im2col(kh, kw, stride, padding, dilation, ho, wo)      # [N*Ho*Wo, K]
        w2 = weight.reshape((k, cout))                                      # reinterpret (A2)
        out2d = col.matmul(w2)                                              # [N*Ho*Wo, Cout]
        out = out2d.reshape((n, ho, wo, cout))
        out = out._transpose_4d_nhwc_to_nchw()
        if bias is not None:
            assert bias.shape == (cout,), "Bias must be 1D with C_out elements"
            out = out._add_bias_4d(bias)
        return out

    def conv2d_relu(self, weight, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1)):
        return self.conv2d(weight, bias, stride, padding, dilation).relu()   # src/tensor.rs:1379-1389

    def _im2col(self, kh, kw, stride, padding, dilation, ho, wo):
        # col[(n,oh,ow), ci*kh*kw + kr*kw + kc] = X[n,ci,oh*s+kr*d-p, ow*s+kc*d-p] or 0
        # (src/tensor.rs:1728-1780 for 3x3/s1/d1; :1806-1969 general; same column order)
        n, cin, h, w = self.shape
        sh, sw = stride
        ph, pw = padding
        dh, dw = dilation
        x = self._data.reshape(n, cin, h, w)
        xp = np.zeros((n, cin, h + 2 * ph, w + 2 * pw), dtype=F32)
        xp[:, :, ph:ph + h, pw:pw + w] = x
        col = np.empty((n, ho, wo, cin, kh, kw), dtype=F32)
        for kr in range(kh):
            for kc in range(kw):
                col[:, :, :, :, kr, kc] = xp[:, :, kr * dh: kr * dh + (ho - 1) * sh + 1: sh,
                                             kc * dw: kc * dw + (wo - 1) * sw + 1: sw].transpose(0, 2, 3, 1)
        out = Tensor(col.reshape(n * ho * wo, cin * kh * kw), (n * ho * wo, cin * kh * kw))
        if self.requires_grad and not Config.strict_reference_conv:
            # full_adjoint: restore the link the reference drops at src/tensor.rs:1725
            out.requires_grad = True
            xin, o = self, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    gc = g.reshape(n, ho, wo, cin, kh, kw)
                    gp = np.zeros((n, cin, h + 2 * ph, w + 2 * pw), dtype=F32)
                    for kr in range(kh):
                        for kc in range(kw):
                            gp[:, :, kr * dh: kr * dh + (ho - 1) * sh + 1: sh,
                               kc * dw: kc * dw + (wo - 1) * sw + 1: sw] += gc[:, :, :, :, kr, kc].transpose(0, 3, 1, 2)
                    gin = xin._grad_slot()
                    gin += gp[:, :, ph:ph + h, pw:pw + w].reshape(-1)
            Tape.push_unary_op(self, out, bw)
        return out

    def _transpose_4d_nhwc_to_nchw(self):     # src/tensor.rs:2034-2076 with axes [0,3,1,2]
        n, ho, wo, c = self.shape
        out = Tensor(np.ascontiguousarray(self._data.reshape(n, ho, wo, c).transpose(0, 3, 1, 2)),
                     (n, c, ho, wo))
        if self.requires_grad and not Config.strict_reference_conv:
            # full_adjoint: restore the link the reference drops at src/tensor.rs:2075
            out.requires_grad = True
            x, o = self, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    gin = x._grad_slot()
                    gin += np.ascontiguousarray(g.reshape(n, c, ho, wo).transpose(0, 2, 3, 1)).reshape(-1)
            Tape.push_unary_op(self, out, bw)
        return out

    def _add_bias_4d(self, bias):             # src/tensor.rs:1972-2031
        n, c, h, w = self.shape
        out = Tensor(self._data.reshape(n, c, h * w) + bias._data[None, :, None], self.shape)
        if self.requires_grad or bias.requires_grad:
            out.requires_grad = True
            x, b, o = self, bias, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    if x.requires_grad:
                        accumulate_grad(x, g)
                    if b.requires_grad:
                        gb = b._grad_slot()
                        gb += g.reshape(n, c, h * w).sum(axis=(0, 2), dtype=F32)
            Tape.push_binary_op(self, bias, out, bw)
        return out

    # ---- max_pool2d  (src/tensor.rs:1391-1521) -------------------------------------------------------------
    def max_pool2d(self, kernel_size, stride=None, padding=(0, 0)):
        assert len(self.shape) == 4
        n, c, h, w = self.shape
        kh, kw = kernel_size
        sh, sw = stride if stride is not None else kernel_size
        ph, pw = padding
        ho = (h + 2 * ph - kh) // sh + 1
        wo = (w + 2 * pw - kw) // sw + 1
        x = self._data.reshape(n * c, h, w)
        best = np.full((n * c, ho, wo), -np.inf, dtype=F32)
        plane_base = (np.arange(n * c, dtype=np.int64) * (h * w))[:, None, None]
        best_idx = np.broadcast_to(plane_base, (n * c, ho, wo)).copy()   # "any valid default"
        oh = np.arange(ho)[:, None]
        ow = np.arange(wo)[None, :]
        for r in range(kh):                                    # scan kh then kw, strict '>' (A6)
            ih = oh * sh + r - ph
            for q in range(kw):
                iw = ow * sw + q - pw
                valid = (ih >= 0) & (ih < h) & (iw >= 0) & (iw < w)
                ihc = np.clip(ih, 0, h - 1)
                iwc = np.clip(iw, 0, w - 1)
                v = x[:, ihc, iwc]
                take = valid[None] & (v > best)
                best = np.where(take, v, best)
                best_idx = np.where(take, plane_base + ihc * w + iwc, best_idx)
        out = Tensor(best, (n, c, ho, wo))
        if self.requires_grad:
            out.requires_grad = True
            xin, o = self, out
            arg = best_idx.reshape(-1)

            def bw():
                g = o._grad[0]
                if g is not None:
                    xin._grad_slot()
                    gin = np.zeros(n * c * h * w, dtype=F32)      # zeroes each plane first (A6)
                    np.add.at(gin, arg, g)
                    xin._grad[0] = gin
            Tape.push_unary_op(self, out, bw)
        return out

    # ---- avg_pool2d  (src/tensor.rs:1524-1660) -------------------------------------------------------------
    def avg_pool2d(self, kernel_size, stride=None, padding=(0, 0)):
        assert len(self.shape) == 4
        n, c, h, w = self.shape
        kh, kw = kernel_size
        sh, sw = stride if stride is not None else kernel_size
        ph, pw = padding
        ho = (h + 2 * ph - kh) // sh + 1
        wo = (w + 2 * pw - kw) // sw + 1
        pool = F32(kh * kw)                                     # count includes padding
        x = self._data.reshape(n * c, h, w)
        xp = np.zeros((n * c, h + 2 * ph, w + 2 * pw), dtype=F32)
        xp[:, ph:ph + h, pw:pw + w] = x
        acc = np.zeros((n * c, ho, wo), dtype=F32)
        for r in range(kh):
            for q in range(kw):
                acc += xp[:, r: r + (ho - 1) * sh + 1: sh, q: q + (wo - 1) * sw + 1: sw]
        out = Tensor(acc / pool, (n, c, ho, wo))
        if self.requires_grad:
            out.requires_grad = True
            xin, o = self, out

            def bw():
                g = o._grad[0]
                if g is not None:
                    gg = g.reshape(n * c, ho, wo) / pool
                    gp = np.zeros((n * c, h + 2 * ph, w + 2 * pw), dtype=F32)
                    for r in range(kh):
                        for q in range(kw):
                            gp[:, r: r + (ho - 1) * sh + 1: sh, q: q + (wo - 1) * sw + 1: sw] += gg
                    gin = xin._grad_slot()                       # accumulates (A6)
                    gin += gp[:, ph:ph + h, pw:pw + w].reshape(-1)
            Tape.push_unary_op(self, out, bw)
        return out


# --------------------------------------------------------------------------------------------
# Kernel layer  (src/gemm.rs, src/ops.rs:124-151)
# --------------------------------------------------------------------------------------------
def sgemm_rowmajor(trans_a, trans_b, m, n, k, alpha, a, b, beta, c):
    """C[m,n] = alpha*op(A)*op(B) + beta*C, row-major; lda/ldb rules of src/gemm.rs:21-29, 88-98.
    `c` is updated in place (flat float32 array)."""
    A = a.reshape(k, m).T if trans_a else a.reshape(m, k)
    B = b.reshape(n, k).T if trans_b else b.reshape(k, n)
    C = c.reshape(m, n)
    prod = A @ B
    if alpha != 1.0:
        prod = F32(alpha) * prod
    if beta == 0.0:
        C[...] = prod
    else:
        if beta != 1.0:
            C *= F32(beta)
        C += prod


def accumulate_grad(t, src):          # src/ops.rs:124-137
    g = t._grad_slot()
    t._grad[0] = g + np.asarray(src, dtype=F32).reshape(-1)   # allocates a temp, like the reference


def accumulate_grad_scaled(t, src, scale):   # src/ops.rs:140-151
    g = t._grad_slot()
    g += F32(scale) * src


# --------------------------------------------------------------------------------------------
# Losses  (src/loss.rs:82-195, 271-290)
# --------------------------------------------------------------------------------------------
def log_softmax(x, dim=-1):           # src/loss.rs:101-126
    nd = len(x.shape)
    d = nd + dim if dim < 0 else dim
    assert d == nd - 1, "Only last-dim log_softmax is supported"
    max_vals, _ = x.max(d)
    shifted = x.sub_broadcast_rows(max_vals)
    sum_exp = shifted.exp().sum(d, True)
    log_sum = sum_exp.log()
    return shifted.sub_broadcast_rows(log_sum)


def softmax(x, dim=-1):
    """Row-wise stable softmax.  The reference body (src/loss.rs:82-98) uses the non-broadcasting
    `-` and `/` and therefore panics for C>1 (A13); this restates the evident intent."""
    nd = len(x.shape)
    d = nd + dim if dim < 0 else dim
    assert d == nd - 1 and nd == 2
    b, c = x.shape
    arr = x._data.reshape(b, c)
    e = np.exp(arr - arr.max(axis=1, keepdims=True))
    return Tensor(e / e.sum(axis=1, keepdims=True, dtype=F32), x.shape)


def cross_entropy_loss(logits, targets):     # src/loss.rs:136-195
    assert len(targets.shape) == 1 or (len(targets.shape) == 2 and targets.shape[1] == 1)
    assert len(logits.shape) == 2 and logits.shape[0] == targets.shape[0]
    b, c = logits.shape
    logp = log_softmax(logits, -1)
    lp = logp._data.reshape(b, c)
    cls = targets._data.astype(np.int64)        # `t[i] as usize`
    assert (cls >= 0).all() and (cls < c).all(), "Target class out of bounds"
    acc = F32(0.0)
    picked = lp[np.arange(b), cls]
    for v in picked:                             # sequential f32 accumulation (src/loss.rs:158-164)
        acc = F32(acc - v)
    out = Tensor.scalar(F32(acc / F32(b)))
    if logits.requires_grad:
        out.requires_grad = True
        lg, o = logits, out

        def bw():
            g = o._grad[0]
            if g is not None:
                grad = np.exp(logp._data).reshape(b, c).copy()
                grad[np.arange(b), cls] -= F32(1.0)
                scale = F32(g[0] / F32(b))
                accumulate_grad(lg, (grad * scale).reshape(-1))
        Tape.push_unary_op(logits, out, bw)
    return out


def bce_loss(predictions, targets):          # src/loss.rs:6-72
    eps = F32(1e-7)
    p, t = predictions._data, targets._data
    assert p.size == t.size, "bce_loss: predictions and targets must match in length"
    n = p.size
    pc = np.clip(p, eps, F32(1.0) - eps).astype(F32)
    terms = (t * np.log(pc) + (F32(1.0) - t) * np.log(F32(1.0) - pc)).astype(F32)
    acc = F32(0.0)
    for v in terms:                              # acc -= ... sequentially (:18-22)
        acc = F32(acc - v)
    out = Tensor.scalar(F32(acc / F32(n)))
    if predictions.requires_grad or targets.requires_grad:
        out.requires_grad = True
        pr, tg, o = predictions, targets, out

        def bw():
            g = o._grad[0]
            if g is not None:
                gs = F32(g[0])
                pcl = np.clip(pr._data, eps, F32(1.0) - eps).astype(F32)
                if pr.requires_grad:
                    gp = pr._grad_slot()
                    gp += (gs * (-(tg._data / pcl - (F32(1.0) - tg._data) / (F32(1.0) - pcl))) / F32(n)).astype(F32)   # :52
                if tg.requires_grad:
                    gy = tg._grad_slot()
                    gy += (gs * (np.log(F32(1.0) - pcl) - np.log(pcl)) / F32(n)).astype(F32)                          # :66
        Tape.push_binary_op(predictions, targets, out, bw)
    return out


def mse_loss(predictions, targets):          # src/loss.rs:75-80
    diff = predictions - targets
    squared = diff * diff
    return squared.mean()


def one_hot(indices, num_classes):           # src/loss.rs:248-268
    assert len(indices.shape) == 1, "Indices must be 1D"
    b = indices.shape[0]
    cls = indices._data.astype(np.int64)
    assert (cls >= 0).all() and (cls < num_classes).all(), "Index out of bounds"
    oh = np.zeros((b, num_classes), F32)
    oh[np.arange(b), cls] = F32(1.0)
    return Tensor(oh.reshape(-1), (b, num_classes))


def cross_entropy_loss_onehot(logits, targets):      # src/loss.rs:202-245
    assert tuple(logits.shape) == tuple(targets.shape), "Logits and targets shapes must match"
    assert len(logits.shape) == 2, "Must be 2D tensors"
    b = logits.shape[0]
    rg_l, rg_t = logits.requires_grad, targets.requires_grad
    logits.requires_grad = targets.requires_grad = False      # the reference rebuilds a fresh scalar (:219-221): no tape link
    try:
        log_probs = log_softmax(logits, -1)
        total = (targets * log_probs).sum(None, False)
    finally:
        logits.requires_grad, targets.requires_grad = rg_l, rg_t
    out = Tensor.scalar(F32(-total._data[0] / F32(b)))
    if logits.requires_grad:
        out.requires_grad = True
        lg, tg, o = logits, targets, out

        def bw():
            g = o._grad[0]
            if g is not None:
                probs = np.exp(log_probs._data)
                grad = ((probs - tg._data) * F32(g[0]) / F32(b)).astype(F32)      # :233-236
                accumulate_grad(lg, grad)
        Tape.push_unary_op(logits, out, bw)
    return out


def accuracy(predictions, targets):          # src/loss.rs:271-290
    assert predictions.shape[0] == targets.shape[0]
    pred = predictions.argmax(1)._data
    t = targets._data
    correct = int((np.abs(pred - t) < 1e-6).sum())
    return F32(correct) / F32(t.size)


# --------------------------------------------------------------------------------------------
# Modules  (src/nn.rs, src/activation.rs)
# --------------------------------------------------------------------------------------------
class Module:
    def forward(self, x):
        raise NotImplementedError

    def parameters(self):
        return []

    def __call__(self, x):
        return self.forward(x)


class Linear(Module):                 # src/nn.rs:28-78
    def __init__(self, in_features, out_features, with_bias=True, rng=None):
        rng = rng if rng is not None else np.random.default_rng()
        scale = math.sqrt(2.0 / in_features)
        w = rng.uniform(-scale, scale, size=in_features * out_features).astype(F32)
        self.weight = Tensor(w, (out_features, in_features)).requires_grad_()
        self.bias = Tensor(np.zeros(out_features, F32), (out_features,)).requires_grad_() if with_bias else None

    def forward(self, x):             # src/nn.rs:54-60
        out = x.matmul(self.weight.transpose())
        if self.bias is not None:
            out = out.add_broadcast(self.bias)
        return out

    def parameters(self):
        return [self.weight] + ([self.bias] if self.bias is not None else [])


class ReLU(Module):                   # src/activation.rs:7-21
    def forward(self, x):
        return x.relu()


class Sigmoid(Module):                # src/activation.rs:37-51
    def forward(self, x):
        return x.sigmoid()


class Sequential(Module):             # src/nn.rs:130-162
    def __init__(self, layers):
        self.layers = list(layers)

    def forward(self, x):
        for l in self.layers:
            x = l.forward(x)
        return x

    def parameters(self):
        return [p for l in self.layers for p in l.parameters()]


class Conv2d(Module):                 # src/nn.rs:180-354 (groups == 1 only; groups>1 is out of scope)
    def __init__(self, cin, cout, kernel_size, stride=None, padding=None, dilation=None, groups=None,
                 bias=True, rng=None):
        rng = rng if rng is not None else np.random.default_rng()
        self.stride = stride or (1, 1)
        self.padding = padding or (0, 0)
        self.dilation = dilation or (1, 1)
        assert (groups or 1) == 1
        kh, kw = kernel_size
        fan_in = cin * kh * kw
        bound = math.sqrt(2.0 / fan_in) * math.sqrt(3.0)
        w = rng.uniform(-bound, bound, size=cout * cin * kh * kw).astype(F32)
        self.weight = Tensor(w, (cout, cin, kh, kw)).requires_grad_()
        self.bias = Tensor(np.zeros(cout, F32), (cout,)).requires_grad_() if bias else None

    def forward(self, x):
        return x.conv2d(self.weight, self.bias, self.stride, self.padding, self.dilation)

    def parameters(self):
        return [self.weight] + ([self.bias] if self.bias is not None else [])


class Conv2dReLU(Conv2d):             # src/nn.rs:433-490
    def forward(self, x):
        return x.conv2d_relu(self.weight, self.bias, self.stride, self.padding, self.dilation)


class MaxPool2d(Module):              # src/nn.rs:508-549
    def __init__(self, kernel_size, stride=None, padding=None):
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding or (0, 0)

    def forward(self, x):
        return x.max_pool2d(self.kernel_size, self.stride, self.padding)


class AdaptiveAvgPool2d(Module):      # src/nn.rs:655-697
    def __init__(self, output_size=(1, 1)):
        self.output_size = output_size

    @staticmethod
    def global_():
        return AdaptiveAvgPool2d((1, 1))

    def forward(self, x):
        _, _, h, w = x.shape
        ho, wo = self.output_size
        return x.avg_pool2d((h // ho, w // wo), (h // ho, w // wo), (0, 0))


class Flatten(Module):                # src/nn.rs:730-756
    def __init__(self, start_dim=1):
        self.start_dim = 1 if start_dim is None else start_dim

    def forward(self, x):
        return x.flatten(self.start_dim)


# --------------------------------------------------------------------------------------------
# Optimizers and LR schedulers  (src/optim.rs)
# --------------------------------------------------------------------------------------------
def powi_f32(a, n):
    """f32::powi — compiler-rt __powisf2 (square-and-multiply in f32)."""
    a = F32(a)
    r = F32(1.0)
    b = int(n)
    recip = b < 0
    b = abs(b)
    while True:
        if b & 1:
            r = F32(r * a)
        b //= 2
        if b == 0:
            break
        a = F32(a * a)
    return F32(1.0) / r if recip else r


class SGD:                            # src/optim.rs:8-40 (momentum argument ignored, :14-17)
    def __init__(self, params, lr, momentum=None):
        self.params, self.lr = list(params), F32(lr)

    def step(self):
        for p in self.params:
            g = p._grad[0]
            if g is not None:
                p._data -= self.lr * g

    def zero_grad(self):
        for p in self.params:
            p.zero_grad()


class Adam:                           # src/optim.rs:43-128
    def __init__(self, params, lr, betas=None, eps=None, weight_decay=None):
        self.params = list(params)
        self.lr = F32(lr)
        self.betas = (F32(0.9), F32(0.999)) if betas is None else (F32(betas[0]), F32(betas[1]))
        self.eps = F32(1e-8 if eps is None else eps)
        self.weight_decay = F32(0.0 if weight_decay is None else weight_decay)
        self.m = [np.zeros(p._data.size, F32) for p in self.params]
        self.v = [np.zeros(p._data.size, F32) for p in self.params]
        self.t = 0

    def step_size(self):
        bc1 = F32(F32(1.0) - powi_f32(self.betas[0], self.t))
        bc2 = F32(F32(1.0) - powi_f32(self.betas[1], self.t))
        return F32(self.lr * F32(np.sqrt(bc2) / bc1))

    def step(self):                   # src/optim.rs:83-113
        self.t += 1
        ss = self.step_size()
        b1, b2 = self.betas
        for i, p in enumerate(self.params):
            g0 = p._grad[0]
            if g0 is None:
                continue
            d = p._data
            g = g0 + self.weight_decay * d
            m = self.m[i]
            v = self.v[i]
            m[...] = b1 * m + (F32(1.0) - b1) * g
            v[...] = b2 * v + (F32(1.0) - b2) * g * g
            d -= ss * m / (np.sqrt(v) + self.eps)

    def zero_grad(self):
        for p in self.params:
            p.zero_grad()

    def get_lr(self):
        return self.lr

    def set_lr(self, lr):
        self.lr = F32(lr)


class AdamW:                          # src/optim.rs:131-181
    def __init__(self, params, lr, betas=None, eps=None, weight_decay=None):
        self.adam = Adam(params, lr, betas, eps, weight_decay)

    def step(self):                   # src/optim.rs:148-168
        wd, lr = self.adam.weight_decay, self.adam.lr
        if wd > 0.0:
            for p in self.adam.params:          # every param, grad or not (A4)
                p._data *= F32(F32(1.0) - lr * wd)
        self.adam.weight_decay = F32(0.0)
        self.adam.step()
        self.adam.weight_decay = wd

    def zero_grad(self):
        self.adam.zero_grad()

    def get_lr(self):
        return self.adam.get_lr()

    def set_lr(self, lr):
        self.adam.set_lr(lr)


class StepLR:                         # src/optim.rs:190-219
    def __init__(self, base_lr, step_size, gamma):
        self.lr, self.step_size, self.gamma, self.epoch = F32(base_lr), step_size, F32(gamma), 0

    def step(self, metrics=None):
        self.epoch += 1
        if self.epoch % self.step_size == 0:
            self.lr = F32(self.lr * self.gamma)

    def get_lr(self):
        return self.lr


class ExponentialLR:                  # src/optim.rs:221-244
    def __init__(self, base_lr, gamma):
        self.lr, self.gamma = F32(base_lr), F32(gamma)

    def step(self, metrics=None):
        self.lr = F32(self.lr * self.gamma)

    def get_lr(self):
        return self.lr


class CosineAnnealingLR:              # src/optim.rs:246-285
    def __init__(self, base_lr, t_max, min_lr=None):
        self.base_lr, self.min_lr = F32(base_lr), F32(0.0 if min_lr is None else min_lr)
        self.lr, self.t_max, self.epoch = F32(base_lr), t_max, 0

    def step(self, metrics=None):
        self.epoch += 1
        progress = F32(F32(self.epoch) / F32(self.t_max))
        cos_val = F32((F32(1.0) + F32(np.cos(F32(progress * F32(np.pi))))) / F32(2.0))
        self.lr = F32(self.min_lr + (self.base_lr - self.min_lr) * cos_val)

    def get_lr(self):
        return self.lr


class ReduceLROnPlateau:              # src/optim.rs:287-352
    def __init__(self, initial_lr, factor, patience, min_lr=None, mode=None):
        self.lr, self.factor, self.patience = F32(initial_lr), F32(factor), patience
        self.min_lr = F32(1e-6 if min_lr is None else min_lr)
        self.mode = mode or "min"
        self.best = F32(np.inf) if self.mode == "min" else F32(-np.inf)
        self.counter = 0

    def step(self, metrics=None):
        if metrics is None:
            return
        improved = metrics < self.best if self.mode == "min" else metrics > self.best
        if improved:
            self.best, self.counter = F32(metrics), 0
        else:
            self.counter += 1
            if self.counter >= self.patience:
                self.lr = max(F32(self.lr * self.factor), self.min_lr)
                self.counter = 0

    def get_lr(self):
        return self.lr


# --------------------------------------------------------------------------------------------
# One training step, exactly the loop body of src/train.rs:106-138 / examples/train_mnist.rs:89-135
# --------------------------------------------------------------------------------------------
def train_step(model, optimizer, images, labels):
    """images: Tensor [B, ...], labels: Tensor [B].  Returns (loss, accuracy) as Python floats."""
    Tape.reset()
    logits = model.forward(images)
    loss = cross_entropy_loss(logits, labels)
    acc = accuracy(logits, labels)
    loss.backward()
    optimizer.step()
    optimizer.zero_grad()
    return float(loss._data[0]), float(acc)


# Canonical model builders for BASELINE.json's configs (weights drawn by OUR seeded generator,
# SURVEY.md §8d; injected identically into the oracle and the CUDA path).
def build_mlp(sizes, rng):
    layers = []
    for i in range(len(sizes) - 1):
        layers.append(Linear(sizes[i], sizes[i + 1], True, rng))
        if i < len(sizes) - 2:
            layers.append(ReLU())
    return Sequential(layers)


def build_cnn2(rng):
    """cfg3(i): Conv3x3(1,32)-ReLU-MaxPool2 - Conv3x3(32,64)-ReLU-MaxPool2 - Flatten - Linear(3136,10)."""
    return Sequential([
        Conv2dReLU(1, 32, (3, 3), (1, 1), (1, 1), None, None, True, rng), MaxPool2d((2, 2), (2, 2)),
        Conv2dReLU(32, 64, (3, 3), (1, 1), (1, 1), None, None, True, rng), MaxPool2d((2, 2), (2, 2)),
        Flatten(1), Linear(64 * 7 * 7, 10, True, rng)])


def build_cnn5(rng):
    """cfg3(ii): the shipped example model, examples/train_mnist_cnn.rs:35-100."""
    return Sequential([
        Conv2dReLU(1, 32, (3, 3), (1, 1), (1, 1), None, None, True, rng),
        Conv2dReLU(32, 32, (3, 3), (1, 1), (1, 1), None, None, True, rng),
        MaxPool2d((2, 2), (2, 2)),
        Conv2dReLU(32, 64, (3, 3), (1, 1), (1, 1), None, None, True, rng),
        Conv2dReLU(64, 64, (3, 3), (1, 1), (1, 1), None, None, True, rng),
        MaxPool2d((2, 2), (2, 2)),
        Conv2dReLU(64, 128, (3, 3), (1, 1), (1, 1), None, None, True, rng),
        AdaptiveAvgPool2d((1, 1)), Flatten(1),
        Linear(128, 128, True, rng), ReLU(), Linear(128, 64, True, rng), ReLU(), Linear(64, 10, True, rng)])
