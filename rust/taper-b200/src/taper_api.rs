//! taper's public API surface (`Tensor`, `Tape`, `nn::Module`, `loss`, `optim`, `train::Trainer`) over libtaper_b200.so, so
//! that `examples/train_mnist.rs` and `examples/train_mnist_cnn.rs` of the reference compile against this crate unchanged
//! (`[lib] name = "taper"` in Cargo.toml): same module paths, type names, method names and argument meaning.
//!
//! NOT COMPILED in the build image (no Rust toolchain there).  Every item cites the reference definition it stands for
//! (vaibhawvipul/taper @ aea74b46) and the C entry point it calls (include/taper_b200_host.h).
//!
//! What stays the reference's own code: `data::mnist` (IDX parsing, shuffling, batching — src/data/mnist.rs).  It only
//! needs `Tensor::new`, so it is re-used verbatim on top of this `Tensor`.
//!
//! Ownership: a `Tensor` is a handle on a ref-counted device tensor (`tp_tensor`), exactly like the reference's
//! `Arc<RwLock<Vec<f32>>>` fields (src/tensor.rs:236-244): `clone()` shares storage, gradient and tape node.
//! The tape is thread-local on the C++ side (one host thread : one CUDA stream), like src/tape.rs:6-9.
use std::ffi::CString;
use std::os::raw::c_int;
use std::ptr;

use crate::check;
use crate::ffi;

// =================================================================================================
// Tensor  (src/tensor.rs:236-244, 469-541; ops: src/ops.rs, src/tensor.rs)
// =================================================================================================
pub struct Tensor {
    h: *mut ffi::tp_tensor,
    shape: Vec<usize>,
}

impl Tensor {
    fn from_raw(h: *mut ffi::tp_tensor) -> Tensor {
        let mut nd: c_int = 0;
        check(unsafe { ffi::tp_tensor_ndim(h, &mut nd) });
        let mut shape = vec![0usize; nd as usize];
        check(unsafe { ffi::tp_tensor_shape(h, shape.as_mut_ptr(), nd) });
        Tensor { h, shape }
    }
    pub(crate) fn raw(&self) -> *mut ffi::tp_tensor { self.h }

    /// `Tensor::new(data, shape)` (src/tensor.rs:470-478): uploads to the device.
    pub fn new(data: Vec<f32>, shape: &[usize]) -> Tensor {
        assert_eq!(data.len(), shape.iter().product::<usize>(), "Data length must match shape");
        let mut h = ptr::null_mut();
        check(unsafe { ffi::tp_tensor_new(data.as_ptr(), shape.as_ptr(), shape.len() as c_int, 0, &mut h) });
        Tensor { h, shape: shape.to_vec() }
    }
    /// `Tensor::scalar` (src/tensor.rs:480-482)
    pub fn scalar(v: f32) -> Tensor { Tensor::new(vec![v], &[1]) }
    /// `fn requires_grad(mut self) -> Self` (src/tensor.rs:484-487)
    pub fn requires_grad(self) -> Tensor {
        let data = self.data();
        let mut h = ptr::null_mut();
        check(unsafe { ffi::tp_tensor_new(data.as_ptr(), self.shape.as_ptr(), self.shape.len() as c_int, 1, &mut h) });
        Tensor { h, shape: self.shape.clone() }
    }
    /// `shape()` (src/tensor.rs:489-491)
    pub fn shape(&self) -> &[usize] { &self.shape }
    /// `data()` (src/tensor.rs:493-496): the reference returns a read guard on the Vec; here a host copy (synchronises the stream).
    /// `loss.data()[0]`, `p.data().len()`, `labels.data()[i]` read the same either way.
    pub fn data(&self) -> Vec<f32> {
        let n: usize = self.shape.iter().product();
        let mut v = vec![0.0f32; n];
        check(unsafe { ffi::tp_tensor_data(self.h, v.as_mut_ptr(), n) });
        v
    }
    /// `data_mut()` writes (src/tensor.rs:499-501)
    pub fn set_data(&self, data: &[f32]) { check(unsafe { ffi::tp_tensor_set_data(self.h, data.as_ptr(), data.len()) }); }
    /// `grad()` (src/tensor.rs:512-518): `None` until a backward pass has written it.
    pub fn grad(&self) -> Option<Vec<f32>> {
        let n: usize = self.shape.iter().product();
        let mut v = vec![0.0f32; n];
        let mut has: c_int = 0;
        check(unsafe { ffi::tp_tensor_grad(self.h, v.as_mut_ptr(), n, &mut has) });
        if has != 0 { Some(v) } else { None }
    }
    /// `zero_grad()` (src/tensor.rs:531-533): the gradient becomes `None`.
    pub fn zero_grad(&self) { check(unsafe { ffi::tp_tensor_zero_grad(self.h) }); }
    /// `backward()` (src/tensor.rs:520-529): seeds ones and replays the tape in reverse (src/tape.rs:106-127).
    pub fn backward(&self) { check(unsafe { ffi::tp_tensor_backward(self.h) }); }

    fn unary(&self, op: &str, arg: f32) -> Tensor {
        let c = CString::new(op).unwrap();
        let mut h = ptr::null_mut();
        check(unsafe { ffi::tp_tensor_unary(c.as_ptr(), self.h, arg, &mut h) });
        Tensor::from_raw(h)
    }
    fn binary(&self, op: &str, other: &Tensor) -> Tensor {
        let c = CString::new(op).unwrap();
        let mut h = ptr::null_mut();
        check(unsafe { ffi::tp_tensor_binary(c.as_ptr(), self.h, other.h, &mut h) });
        Tensor::from_raw(h)
    }
    pub fn matmul(&self, other: &Tensor) -> Tensor { self.binary("matmul", other) }               // src/ops.rs:200-298
    pub fn add_broadcast(&self, other: &Tensor) -> Tensor { self.binary("add_broadcast", other) } // src/tensor.rs:636-704
    pub fn relu(&self) -> Tensor { self.unary("relu", 0.0) }                                       // src/ops.rs:312-374
    pub fn sigmoid(&self) -> Tensor { self.unary("sigmoid", 0.0) }                                 // src/tensor.rs:594-634
    pub fn transpose(&self) -> Tensor { self.unary("transpose", 0.0) }                             // src/tensor.rs:544-591
    pub fn exp(&self) -> Tensor { self.unary("exp", 0.0) }                                         // src/tensor.rs:1091-1133
    pub fn log(&self) -> Tensor { self.unary("log", 0.0) }                                         // src/tensor.rs:1136-1169
    pub fn pow(&self, e: f32) -> Tensor { self.unary("pow", e) }                                   // src/tensor.rs:1172-1206
    pub fn sqrt(&self) -> Tensor { self.pow(0.5) }                                                 // src/tensor.rs:1209-1211
    pub fn mean(&self) -> Tensor { self.unary("mean", 0.0) }                                       // src/tensor.rs:772-800
    /// `reshape` (src/tensor.rs:803-840; copies)
    pub fn reshape(&self, shape: &[usize]) -> Tensor {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::tp_tensor_reshape(self.h, shape.as_ptr(), shape.len() as c_int, &mut h) });
        Tensor::from_raw(h)
    }
    pub fn view(&self, shape: &[usize]) -> Tensor { self.reshape(shape) }                          // src/tensor.rs:1214
    /// `flatten(start_dim)` (src/tensor.rs:842-858)
    pub fn flatten(&self, start_dim: usize) -> Tensor {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::tp_tensor_flatten(self.h, start_dim, &mut h) });
        Tensor::from_raw(h)
    }
    /// `sum(dim, keepdim)` (src/tensor.rs:890-1018)
    pub fn sum(&self, dim: Option<usize>, keepdim: bool) -> Tensor {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::tp_tensor_sum(self.h, dim.map_or(-1, |d| d as c_int), keepdim as c_int, &mut h) });
        Tensor::from_raw(h)
    }
    /// `argmax(dim)` (src/tensor.rs:1086-1088): indices as f32, first maximum wins.
    pub fn argmax(&self, dim: Option<usize>) -> Tensor {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::tp_tensor_argmax(self.h, dim.map_or(-1, |d| d as c_int), &mut h) });
        Tensor::from_raw(h)
    }
}

impl Clone for Tensor {
    /// `Tensor: Clone` shares data, grad and tape node (src/tensor.rs:236-244).
    fn clone(&self) -> Tensor {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::tp_tensor_clone(self.h, &mut h) });
        Tensor { h, shape: self.shape.clone() }
    }
}
impl Drop for Tensor {
    fn drop(&mut self) { unsafe { ffi::tp_tensor_free(self.h); } }
}

macro_rules! binop {
    ($tr:ident, $f:ident, $name:expr) => {
        impl<'a> std::ops::$tr<&'a Tensor> for &'a Tensor {
            type Output = Tensor;
            fn $f(self, rhs: &'a Tensor) -> Tensor { self.binary($name, rhs) }
        }
    };
}
binop!(Add, add, "add"); // src/ops.rs:8-52
binop!(Sub, sub, "sub"); // src/ops.rs:377-420
binop!(Mul, mul, "mul"); // src/ops.rs:54-120
binop!(Div, div, "div"); // src/ops.rs:440-496

// =================================================================================================
// Tape  (src/tape.rs)
// =================================================================================================
pub struct Tape;
impl Tape {
    /// `Tape::reset()` (src/tape.rs:43-49): drops the recorded closures (and the activations they own).
    pub fn reset() { check(unsafe { ffi::tp_tape_reset() }); }
    pub fn len() -> usize {
        let mut n = 0usize;
        check(unsafe { ffi::tp_tape_len(&mut n) });
        n
    }
}

// =================================================================================================
// nn  (src/nn.rs, src/activation.rs)
// =================================================================================================
pub mod nn {
    use super::*;

    /// `trait Module` (src/nn.rs:10-18).  `spec()` is the one addition: the layer's constructor arguments in the layer-list
    /// grammar of `tp_model_create` (taper_b200_host.h), which lets `Sequential::new` rebuild its children as ONE device model —
    /// the form the fused paths need (Linear + ReLU epilogues, the conv stack, the device tape behind `Trainer`).
    pub trait Module {
        fn forward(&self, input: &Tensor) -> Tensor;
        fn parameters(&self) -> Vec<Tensor> { Vec::new() }
        fn spec(&self) -> String;
        fn raw_model(&self) -> *mut ffi::tp_model;
    }

    /// A device model built from a layer list; every concrete layer below is a one-element list.
    pub struct DeviceModel { h: *mut ffi::tp_model, spec: String }
    impl DeviceModel {
        pub fn new(spec: &str, seed: u64) -> DeviceModel {
            let c = CString::new(spec).unwrap();
            let mut h = ptr::null_mut();
            check(unsafe { ffi::tp_model_create(c.as_ptr(), seed, &mut h) });
            DeviceModel { h, spec: spec.to_string() }
        }
        fn forward(&self, x: &Tensor) -> Tensor {
            let mut out = ptr::null_mut();
            check(unsafe { ffi::tp_module_forward(self.h, x.raw(), &mut out) });
            Tensor::from_raw(out)
        }
        fn parameters(&self) -> Vec<Tensor> {
            let mut n: c_int = 0;
            check(unsafe { ffi::tp_model_num_params(self.h, &mut n) });
            (0..n).map(|i| {
                let mut t = ptr::null_mut();
                check(unsafe { ffi::tp_model_parameter(self.h, i, &mut t) });
                Tensor::from_raw(t)
            }).collect()
        }
    }
    impl Drop for DeviceModel {
        fn drop(&mut self) { unsafe { ffi::tp_model_destroy(self.h); } }
    }

    fn fresh_seed() -> u64 {
        // the reference draws from an unseeded thread_rng (src/nn.rs:39, 224): any seed is as good
        use std::time::{SystemTime, UNIX_EPOCH};
        SystemTime::now().duration_since(UNIX_EPOCH).map(|d| d.as_nanos() as u64).unwrap_or(0)
    }

    macro_rules! layer {
        ($name:ident) => {
            impl Module for $name {
                fn forward(&self, input: &Tensor) -> Tensor { self.0.forward(input) }
                fn parameters(&self) -> Vec<Tensor> { self.0.parameters() }
                fn spec(&self) -> String { self.0.spec.clone() }
                fn raw_model(&self) -> *mut ffi::tp_model { self.0.h }
            }
        };
    }

    /// `Linear::new(in, out, bias)` (src/nn.rs:28-78): W [out, in] ~ U[+-sqrt(2/in)], b = 0; forward = x * W^T + b.
    pub struct Linear(DeviceModel);
    impl Linear {
        pub fn new(in_features: usize, out_features: usize, bias: bool) -> Linear {
            Linear(DeviceModel::new(&format!("linear:{}:{}{}", in_features, out_features, if bias { "" } else { ":nobias" }), fresh_seed()))
        }
    }
    layer!(Linear);

    /// `Conv2d::new` / `Conv2dReLU::new(cin, cout, (kh, kw), stride, padding, dilation, groups, bias)` (src/nn.rs:180-354, 433-490).
    /// Square kernels, stride and padding as the examples use them; dilation 1, groups 1 (grouped conv is outside the hot path).
    pub struct Conv2d(DeviceModel);
    pub struct Conv2dReLU(DeviceModel);
    fn conv_spec(kind: &str, cin: usize, cout: usize, k: (usize, usize), stride: Option<(usize, usize)>, padding: Option<(usize, usize)>,
                 dilation: Option<(usize, usize)>, groups: Option<usize>, bias: bool) -> String {
        assert_eq!(k.0, k.1, "square kernels only");
        assert!(dilation.map_or(true, |d| d == (1, 1)) && groups.map_or(true, |g| g == 1), "dilation / groups are outside the hot path");
        assert!(bias, "the layer-list grammar builds conv layers with a bias (as every reference example does)");
        let s = stride.unwrap_or((1, 1));
        let p = padding.unwrap_or((0, 0));
        assert!(s.0 == s.1 && p.0 == p.1, "symmetric stride / padding only");
        format!("{}:{}:{}:{}:{}:{}", kind, cin, cout, k.0, s.0, p.0)
    }
    impl Conv2d {
        #[allow(clippy::too_many_arguments)]
        pub fn new(cin: usize, cout: usize, k: (usize, usize), stride: Option<(usize, usize)>, padding: Option<(usize, usize)>,
                   dilation: Option<(usize, usize)>, groups: Option<usize>, bias: bool) -> Conv2d {
            Conv2d(DeviceModel::new(&conv_spec("conv", cin, cout, k, stride, padding, dilation, groups, bias), fresh_seed()))
        }
    }
    impl Conv2dReLU {
        #[allow(clippy::too_many_arguments)]
        pub fn new(cin: usize, cout: usize, k: (usize, usize), stride: Option<(usize, usize)>, padding: Option<(usize, usize)>,
                   dilation: Option<(usize, usize)>, groups: Option<usize>, bias: bool) -> Conv2dReLU {
            Conv2dReLU(DeviceModel::new(&conv_spec("conv_relu", cin, cout, k, stride, padding, dilation, groups, bias), fresh_seed()))
        }
    }
    layer!(Conv2d);
    layer!(Conv2dReLU);

    /// `MaxPool2d::new(kernel, stride, padding)` (src/nn.rs:508-549)
    pub struct MaxPool2d(DeviceModel);
    impl MaxPool2d {
        pub fn new(kernel: (usize, usize), stride: Option<(usize, usize)>, padding: Option<(usize, usize)>) -> MaxPool2d {
            assert!(padding.map_or(true, |p| p == (0, 0)) && kernel.0 == kernel.1, "square, unpadded pooling windows");
            MaxPool2d(DeviceModel::new(&format!("maxpool:{}:{}", kernel.0, stride.unwrap_or(kernel).0), 0))
        }
    }
    layer!(MaxPool2d);

    /// `AdaptiveAvgPool2d::global()` (src/nn.rs:655-697)
    pub struct AdaptiveAvgPool2d(DeviceModel);
    impl AdaptiveAvgPool2d {
        pub fn global() -> AdaptiveAvgPool2d { AdaptiveAvgPool2d(DeviceModel::new("gap", 0)) }
    }
    layer!(AdaptiveAvgPool2d);

    /// `Flatten::new(Some(1))` (src/nn.rs:730-756)
    pub struct Flatten(DeviceModel);
    impl Flatten {
        pub fn new(start_dim: Option<usize>) -> Flatten {
            assert_eq!(start_dim.unwrap_or(1), 1, "Flatten(1) is what the layer-list grammar builds");
            Flatten(DeviceModel::new("flatten", 0))
        }
    }
    layer!(Flatten);

    /// `Sequential::new(Vec<Box<dyn Module>>)` (src/nn.rs:130-162).  The children are rebuilt as ONE device model (their
    /// specs joined) and their parameter values copied over, so `forward` is a single `Sequential::forward` on the C++ side
    /// with its peepholes, and `Trainer` can hand the whole model to the fused step.
    pub struct Sequential { model: DeviceModel, _children: Vec<Box<dyn Module>> }
    impl Sequential {
        pub fn new(layers: Vec<Box<dyn Module>>) -> Sequential {
            let spec = layers.iter().map(|l| l.spec()).collect::<Vec<_>>().join(",");
            let model = DeviceModel::new(&spec, fresh_seed());
            let mut dst = model.parameters().into_iter();
            for l in &layers {
                for p in l.parameters() {
                    dst.next().expect("parameter count mismatch").set_data(&p.data());
                }
            }
            Sequential { model, _children: layers }
        }
    }
    impl Module for Sequential {
        fn forward(&self, input: &Tensor) -> Tensor { self.model.forward(input) }           // src/nn.rs:149-151
        fn parameters(&self) -> Vec<Tensor> { self.model.parameters() }                      // src/nn.rs:153-155
        fn spec(&self) -> String { self.model.spec.clone() }
        fn raw_model(&self) -> *mut ffi::tp_model { self.model.h }
    }
}

/// `activation::ReLU` (src/activation.rs:7-21): a unit struct in the reference (`Box::new(ReLU)`).
pub mod activation {
    use super::nn::Module;
    use super::*;
    pub struct ReLU;
    impl Module for ReLU {
        fn forward(&self, input: &Tensor) -> Tensor { input.relu() }
        fn spec(&self) -> String { "relu".to_string() }
        fn raw_model(&self) -> *mut ffi::tp_model { ptr::null_mut() }
    }
    pub struct Sigmoid;
    impl Module for Sigmoid {
        fn forward(&self, input: &Tensor) -> Tensor { input.sigmoid() }
        fn spec(&self) -> String { "sigmoid".to_string() }
        fn raw_model(&self) -> *mut ffi::tp_model { ptr::null_mut() }
    }
}

// =================================================================================================
// loss  (src/loss.rs)
// =================================================================================================
pub mod loss {
    use super::*;
    fn loss_of(kind: &str, predictions: &Tensor, targets: &Tensor) -> Tensor {
        let c = CString::new(kind).unwrap();
        let mut h = ptr::null_mut();
        check(unsafe { ffi::tp_loss(c.as_ptr(), predictions.raw(), targets.raw(), &mut h) });
        Tensor::from_raw(h)
    }
    /// `cross_entropy_loss(logits, targets)` (src/loss.rs:136-195): class indices stored as f32, direct gradient.
    pub fn cross_entropy_loss(logits: &Tensor, targets: &Tensor) -> Tensor { loss_of("cross_entropy", logits, targets) }
    pub fn cross_entropy_loss_onehot(logits: &Tensor, targets: &Tensor) -> Tensor { loss_of("cross_entropy_onehot", logits, targets) } // :202-245
    pub fn mse_loss(predictions: &Tensor, targets: &Tensor) -> Tensor { loss_of("mse", predictions, targets) }                       // :75-80
    pub fn bce_loss(predictions: &Tensor, targets: &Tensor) -> Tensor { loss_of("bce", predictions, targets) }                       // :6-72
    /// `accuracy(predictions, targets)` (src/loss.rs:271-290): first-max argmax compared with the f32 class index.
    pub fn accuracy(predictions: &Tensor, targets: &Tensor) -> f32 {
        let mut acc = 0.0f32;
        check(unsafe { ffi::tp_accuracy(predictions.raw(), targets.raw(), &mut acc) });
        acc
    }
}

// =================================================================================================
// optim  (src/optim.rs)
// =================================================================================================
pub mod optim {
    use super::*;

    /// What `Trainer` needs to know about an optimizer to rebuild it next to the fused step (kind and hyper-parameters).
    #[derive(Clone, Debug)]
    pub struct OptimizerConfig { pub kind: &'static str, pub lr: f32, pub beta1: f32, pub beta2: f32, pub eps: f32, pub weight_decay: f32 }

    /// `trait Optimizer` (src/optim.rs:3-6) + `set_lr` as the examples call it on the concrete types.
    pub trait Optimizer {
        fn step(&mut self);
        fn zero_grad(&mut self);
        fn set_lr(&mut self, lr: f32);
        fn config(&self) -> OptimizerConfig;
    }

    pub struct DeviceOptimizer { h: *mut ffi::tp_optimizer, cfg: OptimizerConfig, _params: Vec<Tensor> }
    impl DeviceOptimizer {
        fn new(cfg: OptimizerConfig, params: Vec<Tensor>) -> DeviceOptimizer {
            let kind = CString::new(cfg.kind).unwrap();
            let raw: Vec<*mut ffi::tp_tensor> = params.iter().map(|p| p.raw()).collect();
            let mut h = ptr::null_mut();
            check(unsafe {
                ffi::tp_optimizer_create(kind.as_ptr(), raw.as_ptr() as *const *mut ffi::tp_tensor, raw.len() as c_int, cfg.lr, cfg.beta1,
                                         cfg.beta2, cfg.eps, cfg.weight_decay, &mut h)
            });
            DeviceOptimizer { h, cfg, _params: params }
        }
    }
    impl Drop for DeviceOptimizer {
        fn drop(&mut self) { unsafe { ffi::tp_optimizer_destroy(self.h); } }
    }
    macro_rules! optimizer {
        ($name:ident) => {
            impl Optimizer for $name {
                fn step(&mut self) { check(unsafe { ffi::tp_optimizer_step(self.0.h) }); }             // src/optim.rs:21-33, 83-113, 148-168
                fn zero_grad(&mut self) { check(unsafe { ffi::tp_optimizer_zero_grad(self.0.h) }); }   // :35-39
                fn set_lr(&mut self, lr: f32) {                                                       // :125-127
                    self.0.cfg.lr = lr;
                    check(unsafe { ffi::tp_optimizer_set_lr(self.0.h, lr) });
                }
                fn config(&self) -> OptimizerConfig { self.0.cfg.clone() }
            }
        };
    }
    /// `SGD::new(params, lr, momentum)` (src/optim.rs:8-39; the momentum argument is ignored by the reference, :14-17)
    pub struct SGD(DeviceOptimizer);
    impl SGD {
        pub fn new(params: Vec<Tensor>, lr: f32, _momentum: Option<f32>) -> SGD {
            SGD(DeviceOptimizer::new(OptimizerConfig { kind: "sgd", lr, beta1: 0.9, beta2: 0.999, eps: 1e-8, weight_decay: 0.0 }, params))
        }
    }
    /// `Adam::new(params, lr, betas, eps, weight_decay)` (src/optim.rs:42-127; eps placement and L2 form: SURVEY A4)
    pub struct Adam(DeviceOptimizer);
    impl Adam {
        pub fn new(params: Vec<Tensor>, lr: f32, betas: Option<(f32, f32)>, eps: Option<f32>, weight_decay: Option<f32>) -> Adam {
            let (b1, b2) = betas.unwrap_or((0.9, 0.999));
            Adam(DeviceOptimizer::new(OptimizerConfig { kind: "adam", lr, beta1: b1, beta2: b2, eps: eps.unwrap_or(1e-8),
                                                        weight_decay: weight_decay.unwrap_or(0.0) }, params))
        }
    }
    /// `AdamW::new` (src/optim.rs:130-181): decoupled decay on ALL parameters, grad-less ones included
    pub struct AdamW(DeviceOptimizer);
    impl AdamW {
        pub fn new(params: Vec<Tensor>, lr: f32, betas: Option<(f32, f32)>, eps: Option<f32>, weight_decay: Option<f32>) -> AdamW {
            let (b1, b2) = betas.unwrap_or((0.9, 0.999));
            AdamW(DeviceOptimizer::new(OptimizerConfig { kind: "adamw", lr, beta1: b1, beta2: b2, eps: eps.unwrap_or(1e-8),
                                                         weight_decay: weight_decay.unwrap_or(0.01) }, params))
        }
    }
    optimizer!(SGD);
    optimizer!(Adam);
    optimizer!(AdamW);

    /// LR schedulers (src/optim.rs:184-352): host scalars, `tp_scheduler_*`.
    pub struct LRScheduler { h: *mut ffi::tp_scheduler }
    impl LRScheduler {
        fn make(kind: &str, base_lr: f32, p1: f32, p2: f32, n: usize, mode: &str) -> LRScheduler {
            let (k, m) = (CString::new(kind).unwrap(), CString::new(mode).unwrap());
            let mut h = ptr::null_mut();
            check(unsafe { ffi::tp_scheduler_create(k.as_ptr(), base_lr, p1, p2, n, m.as_ptr(), &mut h) });
            LRScheduler { h }
        }
        pub fn step_lr(base_lr: f32, step_size: usize, gamma: f32) -> LRScheduler { LRScheduler::make("step", base_lr, gamma, 0.0, step_size, "") }  // :190-219
        pub fn exponential(base_lr: f32, gamma: f32) -> LRScheduler { LRScheduler::make("exponential", base_lr, gamma, 0.0, 0, "") }                 // :222-246
        pub fn cosine(base_lr: f32, t_max: usize, eta_min: f32) -> LRScheduler { LRScheduler::make("cosine", base_lr, eta_min, 0.0, t_max, "") }     // :249-285
        pub fn step(&mut self, metric: Option<f32>) { check(unsafe { ffi::tp_scheduler_step(self.h, metric.is_some() as c_int, metric.unwrap_or(0.0)) }); }
        pub fn get_lr(&self) -> f32 {
            let mut lr = 0.0f32;
            check(unsafe { ffi::tp_scheduler_get_lr(self.h, &mut lr) });
            lr
        }
        pub(crate) fn raw(&self) -> *mut ffi::tp_scheduler { self.h }
    }
    impl Drop for LRScheduler {
        fn drop(&mut self) { unsafe { ffi::tp_scheduler_destroy(self.h); } }
    }
}

// =================================================================================================
// train  (src/train.rs)
// =================================================================================================
pub mod train {
    use super::nn::Module;
    use super::optim::{LRScheduler, Optimizer};
    use super::*;

    /// `Trainer { pub model, pub optimizer, .. }` (src/train.rs:73-95).  The examples drive the loop themselves through the two
    /// public fields (examples/train_mnist.rs:89-121): that path is the op-by-op tape.  `step` is the fused form of the same
    /// loop body (src/train.rs:106-138): Tape::reset, forward, loss, accuracy, backward, optimizer.step, zero_grad as one device
    /// step (persistent-kernel tape, tcgen05 kernel plan or CUDA graph, whichever the model qualifies for).
    pub struct Trainer {
        pub model: Box<dyn Module>,
        pub optimizer: Box<dyn Optimizer>,
        pub scheduler: Option<LRScheduler>,
        fused: *mut ffi::tp_trainer,
    }
    impl Trainer {
        pub fn new<O: Optimizer + 'static>(model: Box<dyn Module>, optimizer: O, scheduler: Option<LRScheduler>) -> Trainer {
            Trainer { model, optimizer: Box::new(optimizer), scheduler, fused: ptr::null_mut() }
        }
        fn fused(&mut self) -> *mut ffi::tp_trainer {
            if self.fused.is_null() {
                let cfg = self.optimizer.config();
                let kind = CString::new(cfg.kind).unwrap();
                let m = self.model.raw_model();
                assert!(!m.is_null(), "Trainer::step needs a device model (Sequential or a single layer)");
                check(unsafe { ffi::tp_trainer_create(m, kind.as_ptr(), cfg.lr, cfg.beta1, cfg.beta2, cfg.eps, cfg.weight_decay, &mut self.fused) });
                if let Some(s) = &self.scheduler { check(unsafe { ffi::tp_trainer_set_scheduler(self.fused, s.raw()) }); }
            }
            self.fused
        }
        /// One fused training step on a host batch; returns (loss, correct count).
        pub fn step(&mut self, images: &[f32], labels: &[f32], sample_shape: &[usize]) -> (f32, f32) {
            let t = self.fused();
            let (mut loss, mut correct) = (0.0f32, 0.0f32);
            check(unsafe {
                ffi::tp_trainer_step(t, images.as_ptr(), labels.as_ptr(), labels.len(), sample_shape.as_ptr(), sample_shape.len() as c_int,
                                     &mut loss, &mut correct)
            });
            (loss, correct)
        }
        /// `save_checkpoint` (src/train.rs:264-292): the reference's text format.
        pub fn save_checkpoint(&mut self, path: &str) {
            let t = self.fused();
            let c = CString::new(path).unwrap();
            check(unsafe { ffi::tp_trainer_save_checkpoint(t, c.as_ptr()) });
        }
    }
    impl Drop for Trainer {
        fn drop(&mut self) { if !self.fused.is_null() { unsafe { ffi::tp_trainer_destroy(self.fused); } } }
    }
}
