//! Safe wrappers over the C ABI of libtaper_b200.so (include/taper_b200.h).
//!
//! NOT COMPILED in the build image (no Rust toolchain there) — see rust/README.md.  Line references are to
//! vaibhawvipul/taper @ aea74b46.
use std::cell::RefCell;
use std::ffi::CStr;
use std::os::raw::{c_int, c_void};
use std::ptr;

pub use taper_b200_sys as ffi;

// taper's own API surface (Tensor / Tape / nn / activation / loss / optim / train) over the host ABI: with `[lib] name = "taper"`
// the reference's examples compile against this crate as they are.
mod taper_api;
pub use taper_api::{activation, loss, nn, optim, train, Tape, Tensor};

/// Non-zero status -> panic with the library's thread-local message: the reference panics on the same conditions
/// (`assert!` / `unwrap`, e.g. src/ops.rs:11-15, 201-208).
#[inline]
pub fn check(rc: c_int) {
    if rc != ffi::TP_OK {
        let msg = unsafe { CStr::from_ptr(ffi::tp_last_error()) }.to_string_lossy().into_owned();
        panic!("taper_b200 error {}: {}", rc, msg);
    }
}

/// One host thread : one context : one device : one CUDA stream — the device-side twin of the reference's
/// `thread_local!` tape (src/tape.rs:6-9).
pub struct Ctx(*mut ffi::tp_ctx);

impl Ctx {
    pub fn new(device: i32) -> Ctx {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::tp_ctx_create(device, &mut h) });
        Ctx(h)
    }
    pub fn raw(&self) -> *mut ffi::tp_ctx { self.0 }
    pub fn sync(&self) { check(unsafe { ffi::tp_sync(self.0) }); }
}

impl Drop for Ctx {
    fn drop(&mut self) { unsafe { ffi::tp_ctx_destroy(self.0); } }
}

thread_local! {
    static CTX: RefCell<Option<Ctx>> = RefCell::new(None);
}

/// The calling thread's context (created on device 0 on first use; call `set_device` first to choose another).
pub fn with_ctx<R>(f: impl FnOnce(&Ctx) -> R) -> R {
    CTX.with(|c| {
        let mut c = c.borrow_mut();
        if c.is_none() { *c = Some(Ctx::new(0)); }
        f(c.as_ref().unwrap())
    })
}

pub fn set_device(device: i32) {
    CTX.with(|c| *c.borrow_mut() = Some(Ctx::new(device)));
}

/// Ref-counted device buffer of f32 == the reference's `Arc<RwLock<Vec<f32>>>` (src/tensor.rs:236-244).
pub struct DeviceBuf { h: *mut ffi::tp_buf, len: usize }

impl DeviceBuf {
    pub fn alloc(len: usize) -> DeviceBuf {
        let mut h = ptr::null_mut();
        with_ctx(|c| check(unsafe { ffi::tp_buf_alloc(c.raw(), len.max(1), &mut h) }));
        DeviceBuf { h, len }
    }
    /// `Tensor::new` (src/tensor.rs:470-478)
    pub fn from_slice(data: &[f32]) -> DeviceBuf {
        let b = DeviceBuf::alloc(data.len());
        with_ctx(|c| check(unsafe { ffi::tp_buf_upload(c.raw(), b.h, data.as_ptr() as *const c_void, data.len()) }));
        b
    }
    /// `Tensor::data()` (src/tensor.rs:493-496): synchronises the stream
    pub fn to_vec(&self) -> Vec<f32> {
        let mut v = vec![0.0f32; self.len];
        with_ctx(|c| check(unsafe { ffi::tp_buf_download(c.raw(), self.h, v.as_mut_ptr() as *mut c_void, self.len) }));
        v
    }
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }
    pub fn raw(&self) -> *mut ffi::tp_buf { self.h }
}

impl Clone for DeviceBuf {
    fn clone(&self) -> DeviceBuf {
        check(unsafe { ffi::tp_buf_retain(self.h) });
        DeviceBuf { h: self.h, len: self.len }
    }
}

impl Drop for DeviceBuf {
    fn drop(&mut self) { unsafe { ffi::tp_buf_release(self.h); } }
}

/// Transpose tags of the reference's operator boundary (src/gemm.rs: `n()` / `t()`).
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub enum Transpose { N, T }
pub fn n() -> Transpose { Transpose::N }
pub fn t() -> Transpose { Transpose::T }

/// The reference's operator boundary on device buffers: C = alpha * op(A) * op(B) + beta * C, row-major
/// (src/gemm.rs:8-49; call sites src/ops.rs:215-226, 254-265, 280-291).
#[allow(clippy::too_many_arguments)]
pub fn sgemm_rowmajor_dev(trans_a: Transpose, trans_b: Transpose, m: i32, n: i32, k: i32, alpha: f32, a: &DeviceBuf, b: &DeviceBuf,
                          beta: f32, c: &DeviceBuf) {
    with_ctx(|ctx| {
        check(unsafe {
            ffi::tp_sgemm_rowmajor(ctx.raw(), (trans_a == Transpose::T) as c_int, (trans_b == Transpose::T) as c_int, m, n, k, alpha,
                                   a.raw(), b.raw(), beta, c.raw())
        })
    });
}

/// Same signature as the reference's `pub fn sgemm_rowmajor` (src/gemm.rs:8-19, re-exported at src/lib.rs:13): the third
/// `#[cfg(feature = "b200")]` backend next to cblas / matrixmultiply.  Host slices in, host slice out (two uploads and one
/// download per call) — the drop-in for code that cannot hold device buffers; the `Tensor` integration keeps its data in
/// `DeviceBuf`s and calls `sgemm_rowmajor_dev`.
#[allow(clippy::too_many_arguments)]
pub fn sgemm_rowmajor(trans_a: Transpose, trans_b: Transpose, m: i32, n: i32, k: i32, alpha: f32, a: &[f32], b: &[f32], beta: f32,
                      c: &mut [f32]) {
    assert_eq!(a.len(), (m * k) as usize, "A has the wrong size");
    assert_eq!(b.len(), (k * n) as usize, "B has the wrong size");
    assert_eq!(c.len(), (m * n) as usize, "C has the wrong size");
    let da = DeviceBuf::from_slice(a);
    let db = DeviceBuf::from_slice(b);
    let dc = if beta != 0.0 { DeviceBuf::from_slice(c) } else { DeviceBuf::alloc(c.len()) };
    sgemm_rowmajor_dev(trans_a, trans_b, m, n, k, alpha, &da, &db, beta, &dc);
    c.copy_from_slice(&dc.to_vec());
}

/// `ops::accumulate_grad` / `accumulate_grad_scaled` (src/ops.rs:124-151): `first_touch` stands for the reference's lazily
/// zero-allocated gradient (`None` -> store instead of add).
pub fn accumulate_grad_scaled(dst: &DeviceBuf, src: &DeviceBuf, scale: f32, first_touch: bool) {
    with_ctx(|ctx| check(unsafe { ffi::tp_accumulate(ctx.raw(), dst.raw(), src.raw(), scale, src.len(), (!first_touch) as c_int) }));
}

/// One whole training step (Tape::reset, forward, cross_entropy_loss, accuracy, backward, optimizer.step, zero_grad —
/// the loop body of `Trainer::train_epoch`, src/train.rs:106-138) of a Sequential of Linear(+ReLU) layers as ONE persistent
/// kernel (`tp_step_*`).
pub struct FusedStep { h: *mut ffi::tp_step, result: DeviceBuf }

impl FusedStep {
    /// `params`, `grads`, `m`, `v` are the optimizer's flat arenas; `hyper` the device-resident Adam state
    /// (`tp_adam_hyper_init`); pass `None` for SGD.
    pub fn new(desc: &ffi::tp_step_desc, params: &DeviceBuf, grads: &DeviceBuf, m: Option<&DeviceBuf>, v: Option<&DeviceBuf>,
               hyper: Option<&DeviceBuf>) -> Option<FusedStep> {
        if unsafe { ffi::tp_step_supported(desc) } == 0 { return None; }
        let result = DeviceBuf::alloc(2);
        let mut h = ptr::null_mut();
        let raw = |b: Option<&DeviceBuf>| b.map_or(ptr::null_mut(), |b| b.raw());
        with_ctx(|ctx| {
            check(unsafe {
                ffi::tp_step_create(ctx.raw(), desc, params.raw(), grads.raw(), raw(m), raw(v), raw(hyper), result.raw(), ptr::null_mut(), &mut h)
            })
        });
        Some(FusedStep { h, result })
    }

    /// Host-fed step: `x` [batch, in] and `labels` [batch] already on the device.  Returns (loss, #correct) — this
    /// synchronous form reads the device result; an asynchronous caller passes a pinned slot and a sequence number instead.
    pub fn run(&self, x: &DeviceBuf, labels: &DeviceBuf, sgd_lr: f32) -> (f32, f32) {
        with_ctx(|ctx| {
            check(unsafe {
                ffi::tp_step_run(ctx.raw(), self.h, x.raw(), labels.raw(), ptr::null(), ptr::null_mut(), 0, -1, sgd_lr, 1.0, ptr::null_mut(), 0)
            })
        });
        let r = self.result.to_vec();
        (r[0], r[1])
    }
}

impl Drop for FusedStep {
    fn drop(&mut self) { unsafe { ffi::tp_step_destroy(self.h); } }
}
