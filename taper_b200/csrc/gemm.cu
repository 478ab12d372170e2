// GEMM dispatch + the C ABI of the reference's operator boundary (src/gemm.rs) and the fused Linear.
#include "common.cuh"

namespace tp {

int gemm_rowmajor(tp_ctx* ctx, int ta, int tb, int m, int n, int k, float alpha, const float* a, const float* b,
                  float beta, float* c, const Epilogue& ep) {
    if (ctx->gemm_mode == 3) {
        int rc = gemm_bx3(ctx, ta, tb, m, n, k, alpha, a, b, beta, c, ep);
        if (rc != TP_ERR_UNSUPPORTED) return rc;
    }
    if (ctx->gemm_mode != 0) {
        // bf16x3 shapes TMA cannot describe fall back to the fp32-accurate 3xTF32 kernel
        int rc = gemm_tc(ctx, ta, tb, m, n, k, alpha, a, b, beta, c, ep, ctx->gemm_mode == 3 ? 1 : ctx->gemm_mode);
        if (rc != TP_ERR_UNSUPPORTED) return rc;
    }
    return gemm_simt(ctx, ta, tb, m, n, k, alpha, a, b, beta, c, ep);
}

}  // namespace tp

extern "C" {

int tp_set_gemm_mode(tp_ctx* ctx, int mode) {
    TP_CHECK_ARG(ctx && mode >= 0 && mode <= 3, "tp_set_gemm_mode: mode must be 0 (fp32), 1 (3xTF32), 2 (1xTF32) or 3 (bf16x3)");
    ctx->gemm_mode = mode;
    return TP_OK;
}

int tp_get_gemm_mode(tp_ctx* ctx, int* mode) {
    TP_CHECK_ARG(ctx && mode, "tp_get_gemm_mode: NULL argument");
    *mode = ctx->gemm_mode;
    return TP_OK;
}

int tp_sgemm_rowmajor(tp_ctx* ctx, int trans_a, int trans_b, int m, int n, int k, float alpha, const tp_buf* a,
                      const tp_buf* b, float beta, tp_buf* c) {
    TP_CHECK_ARG(ctx && m >= 0 && n >= 0 && k >= 0, "tp_sgemm_rowmajor: negative dimension (m=%d n=%d k=%d)", m, n, k);
    TP_NEED(a, (size_t)m * k, "a"); TP_NEED(b, (size_t)k * n, "b"); TP_NEED(c, (size_t)m * n, "c");
    tp::Epilogue ep;
    return tp::gemm_rowmajor(ctx, trans_a != 0, trans_b != 0, m, n, k, alpha, a->ptr, b->ptr, beta, c->ptr, ep);
}

int tp_linear_fwd(tp_ctx* ctx, const tp_buf* x, const tp_buf* w, const tp_buf* b, tp_buf* y, int batch, int in_features,
                  int out_features, int relu) {
    TP_CHECK_ARG(ctx && batch >= 0 && in_features > 0 && out_features > 0, "tp_linear_fwd: bad dims");
    TP_NEED(x, (size_t)batch * in_features, "x"); TP_NEED(w, (size_t)out_features * in_features, "w");
    TP_NEED(y, (size_t)batch * out_features, "y");
    if (b) TP_NEED(b, out_features, "b");
    if (tp::linear_skinny_ok(batch, in_features, out_features))
        return tp::linear_skinny_fwd(ctx, x->ptr, w->ptr, b ? b->ptr : nullptr, y->ptr, batch, in_features, out_features, relu);
    tp::Epilogue ep;
    ep.bias = b ? b->ptr : nullptr;
    ep.relu = relu;
    // Y = X * W^T : op(A)=X (N), op(B)=W^T (T; W stored [out,in] = n x k)
    return tp::gemm_rowmajor(ctx, 0, 1, batch, out_features, in_features, 1.0f, x->ptr, w->ptr, 0.0f, y->ptr, ep);
}

int tp_linear_bwd(tp_ctx* ctx, const tp_buf* x, const tp_buf* w, const tp_buf* dy, const tp_buf* relu_mask_y, tp_buf* dx,
                  tp_buf* dw, tp_buf* db, int batch, int in_features, int out_features, int acc_dx, int acc_dw, int acc_db) {
    TP_CHECK_ARG(ctx && batch >= 0 && in_features > 0 && out_features > 0, "tp_linear_bwd: bad dims");
    size_t ny = (size_t)batch * out_features;
    TP_NEED(dy, ny, "dy");
    if (relu_mask_y) TP_NEED(relu_mask_y, ny, "relu_mask_y");
    if (tp::linear_skinny_ok(batch, in_features, out_features)) {
        const size_t nx = (size_t)batch * in_features, nw = (size_t)out_features * in_features;
        if (dx) { TP_NEED(dx, nx, "dx"); TP_NEED(w, nw, "w"); }
        if (dw) { TP_NEED(dw, nw, "dw"); TP_NEED(x, nx, "x"); }
        if (db) TP_NEED(db, out_features, "db");
        return tp::linear_skinny_bwd(ctx, x ? x->ptr : nullptr, w ? w->ptr : nullptr, dy->ptr, relu_mask_y ? relu_mask_y->ptr : nullptr,
                                     dx ? dx->ptr : nullptr, dw ? dw->ptr : nullptr, db ? db->ptr : nullptr, batch, in_features,
                                     out_features, acc_dx, acc_dw, acc_db);
    }
    const tp_buf* g = dy;
    tp_buf* tmp = nullptr;
    int rc = TP_OK;
    if (relu_mask_y) {
        // dZ = dY * [Y>0]  (Y = relu(Z) > 0  <=>  Z > 0, src/ops.rs:367)
        rc = tp_buf_alloc(ctx, ny, &tmp);
        if (rc) return rc;
        rc = tp_relu_bwd(ctx, relu_mask_y, dy, tmp, ny, 0);
        g = tmp;
    }
    tp::Epilogue ep;
    if (!rc && dx) {
        // dX[B,in] (+)= dZ[B,out] * W[out,in]       (N,N)   src/ops.rs:254-265 composed with transpose bwd
        if (!x) { /* x is not needed for dX */ }
        if (dx->n < (size_t)batch * in_features || !w || w->n < (size_t)out_features * in_features) {
            tp::set_error("tp_linear_bwd: dx/w too short");
            rc = TP_ERR_INVALID;
        } else {
            rc = tp::gemm_rowmajor(ctx, 0, 0, batch, in_features, out_features, 1.0f, g->ptr, w->ptr, acc_dx ? 1.0f : 0.0f, dx->ptr, ep);
        }
    }
    if (!rc && dw) {
        // dW[out,in] (+)= dZ^T[out,B] * X[B,in]     (T,N)   src/ops.rs:280-291 + transpose bwd src/tensor.rs:575-586
        if (!x || x->n < (size_t)batch * in_features || dw->n < (size_t)out_features * in_features) {
            tp::set_error("tp_linear_bwd: x/dw missing or too short");
            rc = TP_ERR_INVALID;
        } else {
            rc = tp::gemm_rowmajor(ctx, 1, 0, out_features, in_features, batch, 1.0f, g->ptr, x->ptr, acc_dw ? 1.0f : 0.0f, dw->ptr, ep);
        }
    }
    if (!rc && db) {
        // db[out] (+)= sum_b dZ[b,:]                          src/tensor.rs:680-691
        rc = tp_colsum(ctx, g, db, batch, out_features, 1.0f, acc_db);
    }
    if (tmp) tp_buf_release(tmp);
    return rc;
}

}  // extern "C"
