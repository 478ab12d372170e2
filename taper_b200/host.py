"""Python handles over the flat C entry points of the C++ host layer (include/taper_b200_host.h).

Used by tests/, bench.py and __graft_entry__.py.  All compute runs in libtaper_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import check, lib

F32 = np.float32

# Canonical layer specs for BASELINE.json's configs
MLP_784_128_10 = "linear:784:128,relu,linear:128:10"                       # cfg1 / cfg2
MLP_EXAMPLE = "linear:784:128,relu,linear:128:64,relu,linear:64:10"         # examples/train_mnist.rs:28-51
MLP_784_1024_1024_10 = "linear:784:1024,relu,linear:1024:1024,relu,linear:1024:10"   # cfg4
CNN2 = ("conv_relu:1:32:3:1:1,maxpool:2:2,conv_relu:32:64:3:1:1,maxpool:2:2,flatten,linear:3136:10")   # cfg3(i)
CNN5 = ("conv_relu:1:32:3:1:1,conv_relu:32:32:3:1:1,maxpool:2:2,conv_relu:32:64:3:1:1,conv_relu:64:64:3:1:1,maxpool:2:2,"
        "conv_relu:64:128:3:1:1,gap,flatten,linear:128:128,relu,linear:128:64,relu,linear:64:10")      # examples/train_mnist_cnn.rs:35-100


def _f32(a):
    return np.ascontiguousarray(a, dtype=F32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _shape(dims):
    return (C.c_size_t * len(dims))(*[int(d) for d in dims])


def set_device(device):
    check(lib.tp_host_set_device(int(device)))


def config(conv_full_adjoint=-1, fuse_linear_relu=-1, reference_op_sequence=-1, gemm_mode=-1):
    check(lib.tp_host_config(int(conv_full_adjoint), int(fuse_linear_relu), int(reference_op_sequence), int(gemm_mode)))


def config_conv_stack(fuse=1):
    check(lib.tp_host_config_conv_stack(int(fuse)))


def config_small_mlp(fuse=1):
    check(lib.tp_host_config_small_mlp(int(fuse)))


def host_ctx():
    h = C.c_void_p()
    check(lib.tp_host_ctx(C.byref(h)))
    return h


def launches() -> int:
    c = C.c_uint64()
    check(lib.tp_ctx_launch_count(host_ctx(), C.byref(c)))
    return c.value


def sync():
    check(lib.tp_sync(host_ctx()))


class Event:
    def __init__(self):
        self.h = C.c_void_p()
        check(lib.tp_event_create(host_ctx(), C.byref(self.h)))

    def record(self):
        check(lib.tp_event_record(host_ctx(), self.h))

    def sync(self):
        check(lib.tp_event_sync(self.h))

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float()
        check(lib.tp_event_elapsed_ms(self.h, stop.h, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            lib.tp_event_destroy(self.h)
        except Exception:
            pass


class PinnedArray:
    """float32 numpy view over cudaMallocHost memory."""

    def __init__(self, shape):
        n = int(np.prod(shape))
        self.p = C.c_void_p()
        check(lib.tp_host_alloc_pinned(max(n, 1) * 4, C.byref(self.p)))
        self.array = np.ctypeslib.as_array(C.cast(self.p, C.POINTER(C.c_float)), shape=(n,)).reshape(shape)

    def __del__(self):
        try:
            lib.tp_host_free_pinned(self.p)
        except Exception:
            pass


class Model:
    """nn::Sequential built from a layer spec (see taper_b200_host.h)."""

    def __init__(self, spec: str, seed: int = 0):
        self.h = C.c_void_p()
        self.spec = spec
        check(lib.tp_model_create(spec.encode(), int(seed), C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                lib.tp_model_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def num_params(self) -> int:
        n = C.c_int()
        check(lib.tp_model_num_params(self.h, C.byref(n)))
        return n.value

    def param_shape(self, i):
        numel, nd, dims = C.c_size_t(), C.c_int(), (C.c_size_t * 4)()
        check(lib.tp_model_param_info(self.h, i, C.byref(numel), C.byref(nd), dims))
        return tuple(dims[k] for k in range(nd.value))

    def set_param(self, i, arr):
        a = _f32(arr).reshape(-1)
        check(lib.tp_model_set_param(self.h, i, _fp(a), a.size))

    def get_param(self, i):
        shape = self.param_shape(i)
        out = np.empty(int(np.prod(shape)), F32)
        check(lib.tp_model_get_param(self.h, i, _fp(out), out.size))
        return out.reshape(shape)

    def get_grad(self, i):
        shape = self.param_shape(i)
        out = np.empty(int(np.prod(shape)), F32)
        has = C.c_int()
        check(lib.tp_model_get_grad(self.h, i, _fp(out), out.size, C.byref(has)))
        return out.reshape(shape) if has.value else None

    def zero_grad(self):
        check(lib.tp_model_zero_grad(self.h))

    def forward_tensor(self, x: "Tensor") -> "Tensor":
        """Module::forward on a device tensor (records tape nodes)."""
        h = C.c_void_p()
        check(lib.tp_module_forward(self.h, x.h, C.byref(h)))
        return Tensor(_h=h)

    def parameters(self):
        out = []
        for i in range(self.num_params()):
            h = C.c_void_p()
            check(lib.tp_model_parameter(self.h, i, C.byref(h)))
            out.append(Tensor(_h=h))
        return out

    def load_from_oracle(self, oracle_model):
        ps = oracle_model.parameters()
        assert len(ps) == self.num_params()
        for i, p in enumerate(ps):
            self.set_param(i, p.data())

    def forward(self, x, out_cap=None):
        x = _f32(x)
        cap = out_cap or max(x.shape[0] * 4096, 1 << 16)
        out = np.empty(cap, F32)
        n = C.c_size_t()
        check(lib.tp_model_forward(self.h, _fp(x), _shape(x.shape), x.ndim, _fp(out), cap, C.byref(n)))
        return out[: n.value].copy()

    def regression_backward(self, x, targets, kind):
        """forward; bce / mse / ce_onehot loss against `targets` (output-shaped); backward.  Returns the loss."""
        x, targets = _f32(x), _f32(targets)
        loss = C.c_float()
        check(lib.tp_model_regression_backward(self.h, _fp(x), _shape(x.shape), x.ndim, _fp(targets), targets.size, kind.encode(),
                                               C.byref(loss)))
        return loss.value

    def loss_backward(self, x, labels):
        x, labels = _f32(x), _f32(labels)
        loss, correct, tl = C.c_float(), C.c_float(), C.c_size_t()
        check(lib.tp_model_loss_backward(self.h, _fp(x), _shape(x.shape), x.ndim, _fp(labels), C.byref(loss), C.byref(correct), C.byref(tl)))
        return loss.value, correct.value, tl.value


class Trainer:
    """train::Trainer: one train_epoch iteration per step() (src/train.rs:106-138)."""

    def __init__(self, model: Model, optimizer="adam", lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.model = model
        self.h = C.c_void_p()
        check(lib.tp_trainer_create(model.h, optimizer.encode(), lr, betas[0], betas[1], eps, weight_decay, C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                lib.tp_trainer_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_lr(self, lr):
        check(lib.tp_trainer_set_lr(self.h, lr))

    def set_use_graph(self, on):
        check(lib.tp_trainer_set_use_graph(self.h, int(bool(on))))

    def step(self, images, labels):
        images, labels = _f32(images), _f32(labels)
        loss, correct = C.c_float(), C.c_float()
        check(lib.tp_trainer_step(self.h, _fp(images), _fp(labels), images.shape[0], _shape(images.shape[1:]),
                                  images.ndim - 1, C.byref(loss), C.byref(correct)))
        return loss.value, correct.value

    def step_async(self, images, labels, pinned=True):
        """images/labels: C-contiguous float32 arrays that stay alive (and unmodified) until fetch()."""
        check(lib.tp_trainer_step_async(self.h, _fp(images), _fp(labels), images.shape[0], _shape(images.shape[1:]),
                                        images.ndim - 1, int(pinned)))

    def load_dataset(self, images, labels, perm=None):
        images, labels = _f32(images), _f32(labels)
        p = None
        if perm is not None:
            perm = np.ascontiguousarray(perm, dtype=np.uint32)
            if perm.size != images.shape[0]:          # the C ABI reads one entry per sample
                raise ValueError(f"load_dataset: perm has {perm.size} entries for {images.shape[0]} samples")
            p = perm.ctypes.data_as(C.POINTER(C.c_uint32))
        check(lib.tp_trainer_load_dataset(self.h, _fp(images), _fp(labels), images.shape[0], _shape(images.shape[1:]),
                                          images.ndim - 1, p))

    def step_async_u8(self, images_u8, labels, pinned=True):
        """images_u8: C-contiguous uint8 [B, ...] (MNIST's on-disk pixels); labels float32 [B]; both stay alive until fetch()."""
        assert images_u8.dtype == np.uint8 and images_u8.flags["C_CONTIGUOUS"]
        check(lib.tp_trainer_step_async_u8(self.h, images_u8.ctypes.data_as(C.c_void_p), _fp(labels), images_u8.shape[0],
                                           _shape(images_u8.shape[1:]), images_u8.ndim - 1, int(pinned)))

    def load_dataset_u8(self, images_u8, labels, perm=None):
        images_u8 = np.ascontiguousarray(images_u8, dtype=np.uint8)
        labels = _f32(labels)
        p = None
        if perm is not None:
            perm = np.ascontiguousarray(perm, dtype=np.uint32)
            if perm.size != images_u8.shape[0]:       # the C ABI reads one entry per sample
                raise ValueError(f"load_dataset_u8: perm has {perm.size} entries for {images_u8.shape[0]} samples")
            p = perm.ctypes.data_as(C.POINTER(C.c_uint32))
        check(lib.tp_trainer_load_dataset_u8(self.h, images_u8.ctypes.data_as(C.c_void_p), _fp(labels), images_u8.shape[0],
                                             _shape(images_u8.shape[1:]), images_u8.ndim - 1, p))

    def fused_kind(self) -> int:
        k = C.c_int()
        check(lib.tp_trainer_fused_kind(self.h, C.byref(k)))
        return k.value

    def step_resident(self, batch):
        check(lib.tp_trainer_step_resident(self.h, int(batch)))

    def fetch(self):
        loss, correct = C.c_float(), C.c_float()
        check(lib.tp_trainer_fetch(self.h, C.byref(loss), C.byref(correct)))
        return loss.value, correct.value

    def pending(self) -> int:
        n = C.c_size_t()
        check(lib.tp_trainer_pending(self.h, C.byref(n)))
        return n.value

    def eval(self, images, labels):
        images, labels = _f32(images), _f32(labels)
        loss, correct = C.c_float(), C.c_float()
        check(lib.tp_trainer_eval(self.h, _fp(images), _fp(labels), images.shape[0], _shape(images.shape[1:]),
                                  images.ndim - 1, C.byref(loss), C.byref(correct)))
        return loss.value, correct.value

    def save_checkpoint(self, path):
        check(lib.tp_trainer_save_checkpoint(self.h, str(path).encode()))

    def load_checkpoint(self, path):
        check(lib.tp_trainer_load_checkpoint(self.h, str(path).encode()))

    def comm_init(self, rank, world, unique_id: bytes):
        assert len(unique_id) == 128
        check(lib.tp_trainer_comm_init(self.h, rank, world, unique_id))

    def broadcast_params(self, root=0):
        check(lib.tp_trainer_broadcast_params(self.h, root))

    def peer_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(lib.tp_trainer_peer_handle(self.h, buf))
        return buf.raw

    def peer_connect(self, handles: bytes):
        check(lib.tp_trainer_peer_connect(self.h, handles))

    def peer_exchange_init(self, dist):
        """All-gather the window handles over torch.distributed (plumbing only) and map the peers' windows."""
        handles = [None] * dist.get_world_size()
        dist.all_gather_object(handles, self.peer_handle())
        self.peer_connect(b"".join(handles))

    def train_epoch(self, loader: Loader, max_batches=0):
        """Trainer::train_epoch (src/train.rs:98-144): returns (mean batch loss, accuracy)."""
        loss, acc = C.c_float(), C.c_float()
        check(lib.tp_trainer_train_epoch(self.h, loader.h, int(max_batches), C.byref(loss), C.byref(acc)))
        return loss.value, acc.value

    def evaluate(self, loader: Loader):
        loss, acc = C.c_float(), C.c_float()
        check(lib.tp_trainer_evaluate(self.h, loader.h, C.byref(loss), C.byref(acc)))
        return loss.value, acc.value

    def set_scheduler(self, sched):
        self._sched = sched
        check(lib.tp_trainer_set_scheduler(self.h, sched.h if sched is not None else None))

    def fit(self, train_loader: Loader, val_loader: Loader, epochs, verbose=False):
        check(lib.tp_trainer_fit(self.h, train_loader.h, val_loader.h, int(epochs), int(bool(verbose))))

    def metrics(self):
        out = {}
        for i, name in enumerate(("train_loss", "train_acc", "val_loss", "val_acc", "epoch_times")):
            n = C.c_size_t()
            check(lib.tp_trainer_metrics(self.h, i, None, 0, C.byref(n)))
            buf = np.zeros(max(n.value, 1), F32)
            check(lib.tp_trainer_metrics(self.h, i, _fp(buf), n.value, C.byref(n)))
            out[name] = buf[: n.value].tolist()
        return out

    def get_lr(self) -> float:
        lr = C.c_float()
        check(lib.tp_trainer_get_lr(self.h, C.byref(lr)))
        return lr.value

    def device_error(self) -> int:
        """Reads (and clears) the context's sticky device error word."""
        c = C.c_int()
        check(lib.tp_trainer_device_error(self.h, C.byref(c)))
        return c.value

    def set_use_fused(self, on):
        check(lib.tp_trainer_set_use_fused(self.h, int(bool(on))))

    def fused_steps(self) -> int:
        c = C.c_uint64()
        check(lib.tp_trainer_fused_steps(self.h, C.byref(c)))
        return c.value

    def graph_replays(self) -> int:
        c = C.c_uint64()
        check(lib.tp_trainer_graph_replays(self.h, C.byref(c)))
        return c.value


class Tensor:
    """A device tensor handle of the host layer (taper::Tensor = the reference's Tensor, src/tensor.rs:236-244): every op
    records the tape node the reference records.  What a Rust shim binds (rust/taper-b200)."""

    def __init__(self, data=None, shape=None, requires_grad=False, _h=None):
        if _h is not None:
            self.h = _h
            return
        data = _f32(data)
        shape = tuple(shape) if shape is not None else data.shape
        self.h = C.c_void_p()
        check(lib.tp_tensor_new(_fp(data), _shape(shape), len(shape), int(requires_grad), C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                lib.tp_tensor_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def shape(self):
        nd = C.c_int()
        check(lib.tp_tensor_ndim(self.h, C.byref(nd)))
        dims = (C.c_size_t * nd.value)()
        check(lib.tp_tensor_shape(self.h, dims, nd.value))
        return tuple(dims)

    def data(self):
        out = np.empty(int(np.prod(self.shape)), F32)
        check(lib.tp_tensor_data(self.h, _fp(out), out.size))
        return out.reshape(self.shape)

    def grad(self):
        out = np.empty(int(np.prod(self.shape)), F32)
        has = C.c_int()
        check(lib.tp_tensor_grad(self.h, _fp(out), out.size, C.byref(has)))
        return out.reshape(self.shape) if has.value else None

    def zero_grad(self):
        check(lib.tp_tensor_zero_grad(self.h))

    def backward(self):
        check(lib.tp_tensor_backward(self.h))

    def _un(self, op, arg=0.0):
        h = C.c_void_p()
        check(lib.tp_tensor_unary(op.encode(), self.h, arg, C.byref(h)))
        return Tensor(_h=h)

    def _bin(self, op, other):
        h = C.c_void_p()
        check(lib.tp_tensor_binary(op.encode(), self.h, other.h, C.byref(h)))
        return Tensor(_h=h)

    def relu(self): return self._un("relu")
    def exp(self): return self._un("exp")
    def log(self): return self._un("log")
    def sigmoid(self): return self._un("sigmoid")
    def mean(self): return self._un("mean")
    def transpose(self): return self._un("transpose")
    def pow(self, e): return self._un("pow", float(e))
    def sqrt(self): return self._un("sqrt")
    def matmul(self, o): return self._bin("matmul", o)
    def add_broadcast(self, o): return self._bin("add_broadcast", o)
    def __add__(self, o): return self._bin("add", o)
    def __sub__(self, o): return self._bin("sub", o)
    def __mul__(self, o): return self._bin("mul", o)
    def __truediv__(self, o): return self._bin("div", o)

    def reshape(self, shape):
        h = C.c_void_p()
        check(lib.tp_tensor_reshape(self.h, _shape(shape), len(shape), C.byref(h)))
        return Tensor(_h=h)

    def sum(self, dim=None, keepdim=False):
        h = C.c_void_p()
        check(lib.tp_tensor_sum(self.h, -1 if dim is None else int(dim), int(keepdim), C.byref(h)))
        return Tensor(_h=h)

    def argmax(self, dim=None):
        h = C.c_void_p()
        check(lib.tp_tensor_argmax(self.h, -1 if dim is None else int(dim), C.byref(h)))
        return Tensor(_h=h)


def tape_reset():
    check(lib.tp_tape_reset())


def tape_len() -> int:
    n = C.c_size_t()
    check(lib.tp_tape_len(C.byref(n)))
    return n.value


def loss(kind, predictions: Tensor, targets: Tensor) -> Tensor:
    h = C.c_void_p()
    check(lib.tp_loss(kind.encode(), predictions.h, targets.h, C.byref(h)))
    return Tensor(_h=h)


def accuracy(predictions: Tensor, targets: Tensor) -> float:
    a = C.c_float()
    check(lib.tp_accuracy(predictions.h, targets.h, C.byref(a)))
    return a.value


class Optimizer:
    """optim::{SGD, Adam, AdamW} over tensor handles (src/optim.rs:8-181)."""

    def __init__(self, kind, params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = list(params)
        arr = (C.c_void_p * len(self.params))(*[p.h for p in self.params])
        self.h = C.c_void_p()
        check(lib.tp_optimizer_create(kind.encode(), arr, len(self.params), lr, betas[0], betas[1], eps, weight_decay, C.byref(self.h)))

    def step(self): check(lib.tp_optimizer_step(self.h))
    def zero_grad(self): check(lib.tp_optimizer_zero_grad(self.h))
    def set_lr(self, lr): check(lib.tp_optimizer_set_lr(self.h, lr))

    def get_lr(self):
        lr = C.c_float()
        check(lib.tp_optimizer_get_lr(self.h, C.byref(lr)))
        return lr.value

    def __del__(self):
        try:
            if self.h:
                lib.tp_optimizer_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Dataset:
    """data::MNISTDataset (src/data/mnist.rs:21-25) from arrays (f32 in [0,1] or the raw u8 pixels) or from the IDX files."""

    def __init__(self, images=None, labels=None, mnist_dir=None, train=True):
        self.h = C.c_void_p()
        if mnist_dir is not None:
            check(lib.tp_dataset_load_mnist(str(mnist_dir).encode(), int(train), C.byref(self.h)))
            return
        images = np.ascontiguousarray(images)
        labels = _f32(labels)
        is_u8 = images.dtype == np.uint8
        if not is_u8:
            images = _f32(images)
        n = images.shape[0]
        check(lib.tp_dataset_from_arrays(images.ctypes.data_as(C.c_void_p), int(is_u8), _fp(labels), n, images.size // max(n, 1), C.byref(self.h)))

    def __len__(self):
        n = C.c_size_t()
        check(lib.tp_dataset_len(self.h, C.byref(n)))
        return n.value

    def __del__(self):
        try:
            if self.h:
                lib.tp_dataset_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Loader:
    """data::DataLoader (src/data/mnist.rs:326-385) over a Dataset, with a pinned-memory prefetch pipeline."""

    def __init__(self, dataset: Dataset, batch_size, shuffle=True, seed=0, sample_shape=None):
        self.dataset = dataset
        self.h = C.c_void_p()
        check(lib.tp_loader_create(dataset.h, int(batch_size), int(bool(shuffle)), int(seed), C.byref(self.h)))
        if sample_shape is not None:
            check(lib.tp_loader_set_sample_shape(self.h, _shape(sample_shape), len(sample_shape)))

    def num_batches(self) -> int:
        n = C.c_size_t()
        check(lib.tp_loader_num_batches(self.h, C.byref(n)))
        return n.value

    def __del__(self):
        try:
            if self.h:
                lib.tp_loader_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Scheduler:
    """optim::{StepLR, ExponentialLR, CosineAnnealingLR, ReduceLROnPlateau} (src/optim.rs:190-352)."""

    def __init__(self, kind, base_lr, p1=0.0, p2=0.0, n=0, mode=None):
        self.h = C.c_void_p()
        check(lib.tp_scheduler_create(kind.encode(), base_lr, p1, p2, int(n), mode.encode() if mode else None, C.byref(self.h)))

    def step(self, metric=None):
        check(lib.tp_scheduler_step(self.h, int(metric is not None), 0.0 if metric is None else float(metric)))

    def get_lr(self) -> float:
        lr = C.c_float()
        check(lib.tp_scheduler_get_lr(self.h, C.byref(lr)))
        return lr.value

    def __del__(self):
        try:
            if self.h:
                lib.tp_scheduler_destroy(self.h)
                self.h = None
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib.tp_comm_unique_id(buf))
    return buf.raw
