// Internal definitions of the host layer (not part of the public taper.hpp surface).
#pragma once
#include "taper.hpp"

#include <algorithm>

namespace taper {

// == the shared part of the reference's Tensor: Arc<RwLock<Vec<f32>>> data, Arc<RwLock<Option<Vec<f32>>>> grad,
// Arc<AtomicUsize> tape_node (src/tensor.rs:236-244).  `requires_grad` is by-value and lives in the handle.
struct TensorImpl {
    tp_buf* buf = nullptr;
    Shape shape;
    size_t n = 0;
    tp_buf* grad = nullptr;          // storage (possibly a slice of an optimizer's flat gradient arena)
    bool has_grad = false;           // grad is Some(..)
    size_t tape_node = 0;            // 0 = no node; otherwise tape index + 1
    uint64_t version = 1;            // bumped whenever the device data changes
    uint64_t grad_version = 0;
    mutable std::vector<float> host; // lazily synchronised mirror behind Tensor::data()
    mutable uint64_t host_version = 0;
    ~TensorImpl();
    // gradient buffer to write into and whether to add (1) or store (0): `None -> zeros; +=` semantics
    tp_buf* grad_for_write(int* accumulate);
};

[[noreturn]] void panic(const char* fmt, ...);
size_t shape_numel(const Shape& s);

}  // namespace taper
