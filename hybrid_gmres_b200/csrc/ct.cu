// On-device generation of the synthetic CT system matrices (SURVEY.md §8d
// "Synthetic inputs", §8f-3): the line-intersection projector A (rows = rays,
// angle-major / detector-minor) and the pixel-driven interpolating
// back-projector B (rows = pixels).  The arithmetic is the exact sequence of
// IEEE operations in oracle/ct.py (explicit round-to-nearest intrinsics, no FMA
// contraction), driven by the same host-computed trig tables, so the projector
// is bit-identical to the NumPy generator.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kBlock = 128;
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }

struct Ray {
    double ox, oy, dx, dy;
};

__device__ __forceinline__ Ray make_ray(int geometry, double R, double c, double s, double ra,
                                        double rb) {
    Ray r;
    if (geometry == 0) {
        r.ox = mul(c, ra);
        r.oy = mul(s, ra);
        r.dx = -s;
        r.dy = c;
    } else {
        r.ox = mul(R, c);
        r.oy = mul(R, s);
        r.dx = -sub(mul(c, ra), mul(s, rb));
        r.dy = -add(mul(s, ra), mul(c, rb));
    }
    return r;
}

__device__ __forceinline__ void slab(double o, double d, double half, double& inv, double& tlo,
                                     double& thi) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (d == 0.0) {
        inv = inf;
        const bool inside = (o >= -half) && (o < half);
        tlo = inside ? -inf : inf;
        thi = inf;
    } else {
        inv = dvd(1.0, d);
        const double t1 = mul(sub(-half, o), inv);
        const double t2 = mul(sub(half, o), inv);
        tlo = fmin(t1, t2);
        thi = fmax(t1, t2);
    }
}

// Walks one ray through the N x N grid.  EMIT=false counts entries only.
template <bool EMIT>
__device__ __forceinline__ int trace(const Ray& r, int N, int32_t* __restrict__ cols,
                                     double* __restrict__ vals) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double half = N / 2.0;
    double invdx, txlo, txhi, invdy, tylo, tyhi;
    slab(r.ox, r.dx, half, invdx, txlo, txhi);
    slab(r.oy, r.dy, half, invdy, tylo, tyhi);
    const double tmin = fmax(txlo, tylo);
    const double tmax = fmin(txhi, tyhi);
    if (!(tmax > tmin)) return 0;
    const double ex = add(r.ox, mul(tmin, r.dx));
    const double ey = add(r.oy, mul(tmin, r.dy));
    double ix = fmin(fmax(floor(add(ex, half)), 0.0), (double)(N - 1));
    double iy = fmin(fmax(floor(add(ey, half)), 0.0), (double)(N - 1));
    const double offx = r.dx > 0 ? 1.0 : 0.0, offy = r.dy > 0 ? 1.0 : 0.0;
    const double sgx = r.dx > 0 ? 1.0 : -1.0, sgy = r.dy > 0 ? 1.0 : -1.0;
    const bool zx = r.dx == 0.0, zy = r.dy == 0.0;
    double t = tmin;
    int count = 0;
    for (int it = 0; it < 2 * N + 2; ++it) {
        const double tmx = zx ? inf : mul(sub(sub(add(ix, offx), half), r.ox), invdx);
        const double tmy = zy ? inf : mul(sub(sub(add(iy, offy), half), r.oy), invdy);
        const double tn = fmin(tmx, tmy);
        const double ln = sub(tn, t);
        if (ln > 0) {
            if (EMIT) {
                cols[count] = (int32_t)(((double)(N - 1) - iy) + (double)N * ix);
                vals[count] = ln;
            }
            ++count;
        }
        if (tmx <= tmy) ix += sgx;
        if (tmy <= tmx) iy += sgy;
        t = tn;
        if (!(ix >= 0 && ix < N && iy >= 0 && iy < N)) break;
    }
    return count;
}

template <bool EMIT>
__global__ void __launch_bounds__(kBlock)
projector_kernel(int N, int n_views, int p, int geometry, double R, const double* __restrict__ cos_th,
                 const double* __restrict__ sin_th, const double* __restrict__ ray_a,
                 const double* __restrict__ ray_b, int64_t row_lo, int64_t row_hi,
                 int64_t* __restrict__ rowptr, int32_t* __restrict__ colind, double* __restrict__ vals) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // local row of the block
    if (row >= row_hi - row_lo) return;
    const int64_t grow = row_lo + row;  // global ray index = view*p + i
    const int v = (int)(grow / p), i = (int)(grow % p);
    const Ray r = make_ray(geometry, R, cos_th[v], sin_th[v], ray_a[i], ray_b[i]);
    if (EMIT) {
        const int64_t s = rowptr[row];
        trace<true>(r, N, colind + s, vals + s);
    } else {
        rowptr[row] = trace<false>(r, N, nullptr, nullptr);  // counts, scanned afterwards
    }
}

// pixel-driven back-projector: row = pixel, up to two entries per view
template <bool EMIT>
__global__ void __launch_bounds__(kBlock)
backprojector_kernel(int N, int n_views, int p, int geometry, double R, double gmax, double dg,
                     const double* __restrict__ cos_th, const double* __restrict__ sin_th,
                     int64_t col_lo, int64_t col_hi, int64_t* __restrict__ rowptr,
                     int32_t* __restrict__ colind, double* __restrict__ vals) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (int64_t)N * N) return;
    // row = (N-1-iy) + N*ix
    const int ixp = (int)(row / N);
    const int iyp = (N - 1) - (int)(row % N);
    const double half = N / 2.0;
    const double xc = sub(add((double)ixp, 0.5), half);
    const double yc = sub(add((double)iyp, 0.5), half);
    int64_t pos = EMIT ? rowptr[row] : 0;
    int count = 0;
    for (int v = 0; v < n_views; ++v) {
        const double c = cos_th[v], s = sin_th[v];
        double f, scale;
        if (geometry == 0) {
            f = add(add(mul(xc, c), mul(yc, s)), (p - 1) / 2.0);
            scale = 1.0;
        } else {
            const double rx = sub(xc, mul(R, c));
            const double ry = sub(yc, mul(R, s));
            const double ecx = -c, ecy = -s;
            const double cr = sub(mul(ecx, ry), mul(ecy, rx));
            const double dt = add(mul(ecx, rx), mul(ecy, ry));
            const double g = atan2(cr, dt);
            f = dvd(add(g, gmax), dg);
            scale = dvd(1.0, mul(sqrt(add(mul(rx, rx), mul(ry, ry))), dg));
        }
        const double i0 = floor(f);
        const double w1 = sub(f, i0);
        const double w0 = sub(1.0, w1);
        const int64_t c0 = (int64_t)v * p + (int64_t)i0;  // global sinogram index of bin i0
        if (i0 >= 0 && i0 < p && c0 >= col_lo && c0 < col_hi) {
            if (EMIT) {
                colind[pos + count] = (int32_t)(c0 - col_lo);
                vals[pos + count] = mul(w0, scale);
            }
            ++count;
        }
        if (i0 + 1 >= 0 && i0 + 1 < p && c0 + 1 >= col_lo && c0 + 1 < col_hi) {
            if (EMIT) {
                colind[pos + count] = (int32_t)(c0 + 1 - col_lo);
                vals[pos + count] = mul(w1, scale);
            }
            ++count;
        }
    }
    if (!EMIT) rowptr[row] = count;
}

// counts (stored in rowptr[0..rows)) -> exclusive scan, done on the host (setup path)
int scan_counts(hg_ctx* ctx, int64_t rows, int64_t* d_rowptr, int64_t* total) {
    std::vector<int64_t> h((size_t)rows + 1);
    HG_CUDA(cudaMemcpyAsync(h.data(), d_rowptr, (size_t)rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    int64_t acc = 0;
    for (int64_t i = 0; i < rows; ++i) {
        const int64_t c = h[(size_t)i];
        h[(size_t)i] = acc;
        acc += c;
    }
    h[(size_t)rows] = acc;
    HG_CUDA(cudaMemcpyAsync(d_rowptr, h.data(), (size_t)(rows + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    *total = acc;
    return HG_OK;
}

struct Tables {
    double *c = nullptr, *s = nullptr, *a = nullptr, *b = nullptr;
    ~Tables() {
        cudaFree(c);
        cudaFree(s);
        cudaFree(a);
        cudaFree(b);
    }
};

int upload(hg_ctx* ctx, double** d, const double* h, int n) {
    HG_CUDA(cudaMalloc(d, (size_t)std::max(n, 1) * 8));
    if (h) HG_CUDA(cudaMemcpyAsync(*d, h, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    else HG_CUDA(cudaMemsetAsync(*d, 0, (size_t)std::max(n, 1) * 8, ctx->stream));
    return HG_OK;
}

}  // namespace

extern "C" int hg_ct_projector_rows(hg_ctx* ctx, int N, int n_views, int p, int geometry, double R,
                                    const double* cos_th, const double* sin_th, const double* ray_a,
                                    const double* ray_b, int64_t row_lo, int64_t row_hi, hg_matrix** out) {
    HG_REQUIRE(ctx && cos_th && sin_th && ray_a && out, "hg_ct_projector: NULL argument");
    HG_REQUIRE(row_lo >= 0 && row_lo <= row_hi && row_hi <= (int64_t)n_views * p, "hg_ct_projector: bad row range");
    HG_REQUIRE(N >= 1 && n_views >= 1 && p >= 1, "hg_ct_projector: bad sizes");
    HG_REQUIRE(geometry == 0 || (geometry == 1 && ray_b), "hg_ct_projector: bad geometry");
    HG_REQUIRE((int64_t)N * N <= 2147483647LL, "hg_ct_projector: image too large for int32 columns");
    HG_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    const int64_t rows = row_hi - row_lo;
    Tables t;
    HG_TRY(upload(ctx, &t.c, cos_th, n_views));
    HG_TRY(upload(ctx, &t.s, sin_th, n_views));
    HG_TRY(upload(ctx, &t.a, ray_a, p));
    HG_TRY(upload(ctx, &t.b, geometry == 1 ? ray_b : nullptr, p));
    int64_t* d_ptr = nullptr;
    HG_CUDA(cudaMalloc(&d_ptr, (size_t)(rows + 1) * 8));
    const unsigned grid = (unsigned)std::max<int64_t>(cdiv(rows, kBlock), 1);
    {
        hg_launch_scope scope(ctx, HG_K_SETUP, 8.0 * (double)rows);
        projector_kernel<false><<<grid, kBlock, 0, ctx->stream>>>(N, n_views, p, geometry, R, t.c, t.s,
                                                                   t.a, t.b, row_lo, row_hi, d_ptr, nullptr, nullptr);
    }
    int64_t nnz = 0;
    int st = cudaGetLastError() == cudaSuccess ? scan_counts(ctx, rows, d_ptr, &nnz) : HG_ERR_CUDA;
    hg_matrix* m = nullptr;
    if (st == HG_OK) st = hg_matrix_alloc(ctx, rows, (int64_t)N * N, nnz, &m);
    if (st == HG_OK) {
        cudaMemcpyAsync(m->rowptr, d_ptr, (size_t)(rows + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream);
        hg_launch_scope scope(ctx, HG_K_SETUP, 12.0 * (double)nnz);
        projector_kernel<true><<<grid, kBlock, 0, ctx->stream>>>(N, n_views, p, geometry, R, t.c, t.s,
                                                                  t.a, t.b, row_lo, row_hi, m->rowptr, m->colind, m->vals);
    }
    if (st == HG_OK && (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess)) {
        hg_set_error("hg_ct_projector: kernel failed");
        st = HG_ERR_CUDA;
    }
    cudaFree(d_ptr);
    if (st != HG_OK) {
        hg_matrix_destroy(m);
        return st;
    }
    hg_matrix_pick_tpr(m);
    *out = m;
    return HG_OK;
}

extern "C" int hg_ct_projector(hg_ctx* ctx, int N, int n_views, int p, int geometry, double R,
                               const double* cos_th, const double* sin_th, const double* ray_a,
                               const double* ray_b, hg_matrix** out) {
    return hg_ct_projector_rows(ctx, N, n_views, p, geometry, R, cos_th, sin_th, ray_a, ray_b, 0,
                                (int64_t)n_views * p, out);
}

extern "C" int hg_ct_backprojector_cols(hg_ctx* ctx, int N, int n_views, int p, int geometry, double R,
                                        const double* cos_th, const double* sin_th, int64_t col_lo,
                                        int64_t col_hi, hg_matrix** out) {
    HG_REQUIRE(ctx && cos_th && sin_th && out, "hg_ct_backprojector: NULL argument");
    HG_REQUIRE(col_lo >= 0 && col_lo <= col_hi && col_hi <= (int64_t)n_views * p, "hg_ct_backprojector: bad column range");
    HG_REQUIRE(N >= 1 && n_views >= 1 && p >= 2, "hg_ct_backprojector: bad sizes");
    HG_REQUIRE(geometry == 0 || geometry == 1, "hg_ct_backprojector: bad geometry");
    HG_REQUIRE((int64_t)n_views * p <= 2147483647LL, "hg_ct_backprojector: sinogram too large for int32 columns");
    HG_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    const int64_t rows = (int64_t)N * N;
    double gmax = 0.0, dg = 1.0;
    if (geometry == 1) {
        gmax = asin(sqrt(2.0) / 2.0 * N / R);
        dg = 2.0 * gmax / (p - 1);
    }
    Tables t;
    HG_TRY(upload(ctx, &t.c, cos_th, n_views));
    HG_TRY(upload(ctx, &t.s, sin_th, n_views));
    int64_t* d_ptr = nullptr;
    HG_CUDA(cudaMalloc(&d_ptr, (size_t)(rows + 1) * 8));
    const unsigned grid = (unsigned)cdiv(rows, kBlock);
    {
        hg_launch_scope scope(ctx, HG_K_SETUP, 8.0 * (double)rows);
        backprojector_kernel<false><<<grid, kBlock, 0, ctx->stream>>>(N, n_views, p, geometry, R, gmax, dg,
                                                                       t.c, t.s, col_lo, col_hi, d_ptr, nullptr, nullptr);
    }
    int64_t nnz = 0;
    int st = cudaGetLastError() == cudaSuccess ? scan_counts(ctx, rows, d_ptr, &nnz) : HG_ERR_CUDA;
    hg_matrix* m = nullptr;
    if (st == HG_OK) st = hg_matrix_alloc(ctx, rows, col_hi - col_lo, nnz, &m);
    if (st == HG_OK) {
        cudaMemcpyAsync(m->rowptr, d_ptr, (size_t)(rows + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream);
        hg_launch_scope scope(ctx, HG_K_SETUP, 12.0 * (double)nnz);
        backprojector_kernel<true><<<grid, kBlock, 0, ctx->stream>>>(N, n_views, p, geometry, R, gmax, dg,
                                                                      t.c, t.s, col_lo, col_hi, m->rowptr, m->colind, m->vals);
    }
    if (st == HG_OK && (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess)) {
        hg_set_error("hg_ct_backprojector: kernel failed");
        st = HG_ERR_CUDA;
    }
    cudaFree(d_ptr);
    if (st != HG_OK) {
        hg_matrix_destroy(m);
        return st;
    }
    hg_matrix_pick_tpr(m);
    *out = m;
    return HG_OK;
}

extern "C" int hg_ct_backprojector(hg_ctx* ctx, int N, int n_views, int p, int geometry, double R,
                                   const double* cos_th, const double* sin_th, hg_matrix** out) {
    return hg_ct_backprojector_cols(ctx, N, n_views, p, geometry, R, cos_th, sin_th, 0, (int64_t)n_views * p, out);
}
