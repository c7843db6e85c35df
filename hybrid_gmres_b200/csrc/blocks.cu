// Building-block entry points on HOST vectors (tests and setup): each copies its
// operands to the device, runs the same kernels the solvers use, and copies back.
#include <algorithm>

#include "common.cuh"

namespace {
struct DBuf {
    double* p = nullptr;
    ~DBuf() { hg_dfree(p); }
    int alloc(size_t n) {
        cudaError_t e = hg_dmalloc_cur((void**)&p, std::max<size_t>(n, 1) * sizeof(double));
        if (e != cudaSuccess) {
            hg_set_error("device allocation of %zu doubles failed: %s", n, cudaGetErrorString(e));
            return HG_ERR_NOMEM;
        }
        return HG_OK;
    }
};
inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
}  // namespace

extern "C" int hg_spmv(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y) {
    HG_REQUIRE(ctx && m && x && y, "hg_spmv: NULL argument");
    HG_CUDA(cudaSetDevice(ctx->device));
    hg_alloc_scope alloc_scope(ctx);  // RAII buffers below come from / return to this context's cache
    DBuf dx, dy;
    HG_TRY(dx.alloc((size_t)m->cols));
    HG_TRY(dy.alloc((size_t)m->rows));
    HG_CUDA(cudaMemcpyAsync(dx.p, x, (size_t)m->cols * 8, cudaMemcpyHostToDevice, ctx->stream));
    hg_spmv_epilogue ep;
    HG_TRY(hg_k_spmv(ctx, m, dx.p, dy.p, ep, nullptr));
    HG_CUDA(cudaMemcpyAsync(y, dy.p, (size_t)m->rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    return HG_OK;
}

static int upload_basis(hg_ctx* ctx, int64_t n, int k, const double* V, int64_t ld, DBuf& dV,
                        int64_t* ldd) {
    *ldd = round_up(std::max<int64_t>(n, 1), 32);
    HG_TRY(dV.alloc((size_t)(*ldd) * std::max(k, 1)));
    if (k > 0 && n > 0)
        HG_CUDA(cudaMemcpy2DAsync(dV.p, (size_t)(*ldd) * 8, V, (size_t)ld * 8, (size_t)n * 8, (size_t)k,
                                  cudaMemcpyHostToDevice, ctx->stream));
    return HG_OK;
}

extern "C" int hg_multidot(hg_ctx* ctx, int64_t n, int k, const double* V, int64_t ld,
                           const double* w, double* h) {
    HG_REQUIRE(ctx && V && w && h, "hg_multidot: NULL argument");
    HG_REQUIRE(k >= 0 && n >= 0 && ld >= n, "hg_multidot: bad shape");
    HG_CUDA(cudaSetDevice(ctx->device));
    hg_alloc_scope alloc_scope(ctx);  // RAII buffers below come from / return to this context's cache
    DBuf dV, dw, dh, dp;
    int64_t ldd = 0;
    HG_TRY(upload_basis(ctx, n, k, V, ld, dV, &ldd));
    HG_TRY(dw.alloc((size_t)n));
    HG_TRY(dh.alloc((size_t)k));
    HG_TRY(dp.alloc((size_t)(k + 1) * ((size_t)n / 256 + 2)));
    HG_CUDA(cudaMemcpyAsync(dw.p, w, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    int ns = 0;
    HG_TRY(hg_k_multidot(ctx, dV.p, ldd, n, k, dw.p, dp.p, &ns));
    HG_TRY(hg_k_reduce(ctx, dp.p, ns, k, dh.p, false, nullptr, false));
    if (k > 0) HG_CUDA(cudaMemcpyAsync(h, dh.p, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    return HG_OK;
}

extern "C" int hg_lincomb(hg_ctx* ctx, int64_t n, int k, const double* V, int64_t ld,
                          const double* c, double s, const double* z, double* out,
                          double* out_norm2) {
    HG_REQUIRE(ctx && V && c && out, "hg_lincomb: NULL argument");
    HG_REQUIRE(k >= 0 && n >= 0 && ld >= n, "hg_lincomb: bad shape");
    HG_CUDA(cudaSetDevice(ctx->device));
    hg_alloc_scope alloc_scope(ctx);  // RAII buffers below come from / return to this context's cache
    DBuf dV, dc, dz, dout, dstat;
    int64_t ldd = 0;
    HG_TRY(upload_basis(ctx, n, k, V, ld, dV, &ldd));
    HG_TRY(dc.alloc((size_t)k));
    HG_TRY(dz.alloc((size_t)n));
    HG_TRY(dout.alloc((size_t)n));
    HG_TRY(dstat.alloc((size_t)n / 256 + 16));
    if (k > 0) HG_CUDA(cudaMemcpyAsync(dc.p, c, (size_t)k * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (z) HG_CUDA(cudaMemcpyAsync(dz.p, z, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    int np = 0;
    HG_TRY(hg_k_lincomb(ctx, dV.p, ldd, n, k, dc.p, s, z ? dz.p : nullptr, dout.p, nullptr,
                        out_norm2 ? dstat.p : nullptr, &np));
    if (out_norm2) {
        HG_TRY(hg_k_reduce(ctx, dstat.p, np, 1, ctx->d_scalars, false, nullptr, false));
        HG_CUDA(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    HG_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out_norm2) *out_norm2 = ctx->h_scalars[0];
    return HG_OK;
}


// CGS2 middle stage on host arrays: w1 = w0 - V h, d = V' w1.  `fused` != 0 uses the shared-memory
// staged one-pass kernel (cgs_staged.cu; HG_ERR_INVALID when (n, k) is outside its range), 0 the
// separate update + multi-dot kernels.
extern "C" int hg_cgs_mid(hg_ctx* ctx, int64_t n, int k, const double* V, int64_t ld, const double* h,
                          const double* w0, int fused, double* w1, double* d) {
    HG_REQUIRE(ctx && V && h && w0 && w1 && d, "hg_cgs_mid: NULL argument");
    HG_REQUIRE(k >= 1 && n >= 1 && ld >= n, "hg_cgs_mid: bad shape");
    HG_CUDA(cudaSetDevice(ctx->device));
    hg_alloc_scope alloc_scope(ctx);  // RAII buffers below come from / return to this context's cache
    DBuf dV, dh, dw0, dw1, dd, dp;
    int64_t ldd = 0;
    HG_TRY(upload_basis(ctx, n, k, V, ld, dV, &ldd));
    HG_TRY(dh.alloc((size_t)k));
    HG_TRY(dw0.alloc((size_t)ldd));
    HG_TRY(dw1.alloc((size_t)ldd));
    HG_TRY(dd.alloc((size_t)k));
    HG_TRY(dp.alloc((size_t)(k + 1) * (size_t)(std::max<int64_t>(n / 256 + 2, ctx->sm_count) + 1)));
    HG_CUDA(cudaMemcpyAsync(dh.p, h, (size_t)k * 8, cudaMemcpyHostToDevice, ctx->stream));
    HG_CUDA(cudaMemcpyAsync(dw0.p, w0, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    int ns = 0;
    if (fused) {
        HG_REQUIRE(hg_cgs_staged_nparts(ctx, n, k) > 0, "hg_cgs_mid: (n=%lld, k=%d) is outside the staged kernel's range",
                   (long long)n, k);
        HG_TRY(hg_k_cgs_mid_staged(ctx, dV.p, ldd, n, k, dh.p, dw0.p, dw1.p, dp.p, &ns));
    } else {
        HG_TRY(hg_k_lincomb(ctx, dV.p, ldd, n, k, dh.p, -1.0, dw0.p, dw1.p, nullptr, nullptr, nullptr));
        HG_TRY(hg_k_multidot(ctx, dV.p, ldd, n, k, dw1.p, dp.p, &ns));
    }
    HG_TRY(hg_k_reduce(ctx, dp.p, ns, k, dd.p, false, nullptr, false));
    HG_CUDA(cudaMemcpyAsync(w1, dw1.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaMemcpyAsync(d, dd.p, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    return HG_OK;
}


// One whole CGS2 step on host arrays (cgs2_step.cu): hcol[0..k) = h1 + h2, hcol[k] = ||v||, q = v / ||v||.
extern "C" int hg_cgs2_step(hg_ctx* ctx, int64_t n, int k, const double* V, int64_t ld, const double* w0,
                            double* hcol, double* q) {
    HG_REQUIRE(ctx && V && w0 && hcol && q, "hg_cgs2_step: NULL argument");
    HG_REQUIRE(k >= 1 && n >= 1 && ld >= n, "hg_cgs2_step: bad shape");
    HG_REQUIRE(k <= 208, "hg_cgs2_step: k = %d is outside the kernel's range (1..208)", k);
    HG_CUDA(cudaSetDevice(ctx->device));
    hg_alloc_scope alloc_scope(ctx);
    DBuf dV, dw0, dw1, dq, dH, dhc, dp;
    int64_t ldd = 0;
    HG_TRY(upload_basis(ctx, n, k, V, ld, dV, &ldd));
    HG_TRY(dw0.alloc((size_t)ldd));
    HG_TRY(dw1.alloc((size_t)ldd));
    HG_TRY(dq.alloc((size_t)ldd));
    HG_TRY(dH.alloc((size_t)k + 2));
    HG_TRY(dhc.alloc((size_t)k + 2));
    HG_TRY(dp.alloc((size_t)(k + 2) * (size_t)(ctx->sm_count + 1)));
    HG_CUDA(cudaMemcpyAsync(dw0.p, w0, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    HG_TRY(hg_k_cgs2_step(ctx, dV.p, ldd, n, k, dw0.p, dw1.p, dq.p, dH.p, dhc.p, dp.p));
    HG_CUDA(cudaMemcpyAsync(hcol, dH.p, (size_t)(k + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaMemcpyAsync(q, dq.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    return HG_OK;
}
