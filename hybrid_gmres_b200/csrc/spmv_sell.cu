// Row-per-lane SpMV over a 32-row sliced copy of a CSR matrix (SELL-32).
//
// Why: the row-per-warp CSR kernel (kernels.cu) gathers x for 32 CONSECUTIVE ENTRIES OF ONE
// ROW per instruction.  For a pixel-driven back-projector B (row = pixel, two detector bins
// per view) those 32 entries belong to 16 different views, i.e. 16 different 128-byte lines:
// ncu (profiles/r01_spmv_variants.md) shows the L1 data pipe at 81-85 % of its wavefront peak
// and the kernel slowing down with the SM clock under the power cap.  Here a lane owns a ROW
// and the warp walks 32 adjacent rows in lock step: entry j of 32 adjacent pixels is the same
// view and a window of neighbouring bins, i.e. 1-2 lines per gather, and the matrix stream
// stays perfectly coalesced because the slice is stored column-major (entry j of lane l at
// slice_base + 32 j + l).  Same entries, same per-row order as the CSR matrix; only slices
// whose rows have (nearly) equal length are worth it, so the form is built only when padding
// costs <= 8 % (B of every CT configuration: exactly 2*views entries per row).
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kBlock = 256;
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ double ld_stream(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// widths[s] = max row length in slice s
__global__ void slice_width_kernel(int64_t rows, const int64_t* __restrict__ rowptr,
                                   int32_t* __restrict__ widths) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int len = row < rows ? (int)(rowptr[row + 1] - rowptr[row]) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && (row >> 5) < ((rows + 31) >> 5)) widths[row >> 5] = len;
}

// Gather locality of the two traversals, sampled: distinct 128-byte lines of x per gathered entry
// when (a) the 32 lanes hold entry j of 32 adjacent rows (sliced form) and (b) the 32 lanes hold 32
// consecutive entries of one row (row per warp).  counts = {lines_a, entries_a, lines_b, entries_b}.
__device__ __forceinline__ int distinct_lines(bool active, int col) {
    const unsigned mask = __ballot_sync(0xffffffffu, active);
    int n = 0;
    if (active) {
        const unsigned same = __match_any_sync(mask, col >> 4);
        const bool leader = (__ffs(same) - 1) == (int)(threadIdx.x & 31);
        n = __popc(__ballot_sync(mask, leader));
    }
    return __shfl_sync(0xffffffffu, n, mask ? __ffs(mask) - 1 : 0);
}

__global__ void __launch_bounds__(kBlock)
gather_lines_kernel(int64_t rows, int64_t nslices, int64_t stride, const int64_t* __restrict__ rowptr,
                    const int32_t* __restrict__ colind, unsigned long long* __restrict__ counts) {
    const int64_t slice = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * stride;
    const int lane = threadIdx.x & 31;
    if (slice >= nslices) return;
    const int64_t row = slice * 32 + lane;
    const int64_t s = row < rows ? rowptr[row] : 0;
    const int len = row < rows ? (int)(rowptr[row + 1] - s) : 0;
    int width = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) width = max(width, __shfl_xor_sync(0xffffffffu, width, o));
    unsigned long long la = 0, ea = 0, lb = 0, eb = 0;
    for (int j = 0; j < width; j += 3) {  // (a) sampled columns of the slice
        const bool act = j < len;
        const int c = act ? colind[s + j] : 0;
        la += distinct_lines(act, c);
        ea += __popc(__ballot_sync(0xffffffffu, act));
    }
    const int64_t s0 = __shfl_sync(0xffffffffu, s, 0);  // (b) the slice's first row
    const int len0 = __shfl_sync(0xffffffffu, len, 0);
    for (int i0 = 0; i0 < len0; i0 += 32) {
        const bool act = i0 + lane < len0;
        const int c = act ? colind[s0 + i0 + lane] : 0;
        lb += distinct_lines(act, c);
        eb += __popc(__ballot_sync(0xffffffffu, act));
    }
    if (lane == 0) {
        atomicAdd(counts + 0, la);
        atomicAdd(counts + 1, ea);
        atomicAdd(counts + 2, lb);
        atomicAdd(counts + 3, eb);
    }
}

// one warp per slice, 32 x 32 tiles through shared memory: a row's entries are read coalesced
// (32 consecutive entries per trip), the slice columns are written coalesced (32 consecutive rows
// per trip); padding slots repeat the row's last column with a zero value so padded gathers stay
// inside the row's own footprint
constexpr int kFillBlock = 96;  // 3 warps: 3 x 32 x 33 x 12 B = 38 KB of static shared memory

__global__ void __launch_bounds__(kFillBlock)
sell_fill_kernel(int64_t rows, int64_t nslices, const int64_t* __restrict__ rowptr,
                 const int32_t* __restrict__ colind, const double* __restrict__ vals,
                 const int64_t* __restrict__ sptr, int32_t* __restrict__ scol,
                 double* __restrict__ sval) {
    __shared__ int t_col[kFillBlock / 32][32][33];
    __shared__ double t_val[kFillBlock / 32][32][33];
    const int64_t slice = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (slice >= nslices) return;  // whole warps leave together; no block-wide barrier below
    const int64_t base = sptr[slice];
    const int width = (int)((sptr[slice + 1] - base) >> 5);
    // lane r keeps the extent of row r of the slice
    const int64_t my_row = slice * 32 + lane;
    const int64_t my_s = my_row < rows ? rowptr[my_row] : 0;
    const int my_len = my_row < rows ? (int)(rowptr[my_row + 1] - my_s) : 0;
    const int my_pad = my_len > 0 ? colind[my_s + my_len - 1] : 0;
    for (int j0 = 0; j0 < width; j0 += 32) {
        for (int r = 0; r < 32; ++r) {
            const int64_t s = __shfl_sync(0xffffffffu, my_s, r);
            const int len = __shfl_sync(0xffffffffu, my_len, r);
            const int pad = __shfl_sync(0xffffffffu, my_pad, r);
            const int j = j0 + lane;
            const bool in = j < len;
            const bool exists = slice * 32 + r < rows;
            // rows past the end of the matrix copy the columns of the slice's first row (zero values),
            // so a group's column span stays that of real rows (16-bit offsets, spmv_idx16.cu)
            t_col[w][r][lane] = in ? colind[s + j] : (exists ? pad : t_col[w][0][lane]);
            t_val[w][r][lane] = in ? vals[s + j] : 0.0;
        }
        __syncwarp();
        const int jn = min(32, width - j0);
        for (int j = 0; j < jn; ++j) {
            const int64_t dst = base + (int64_t)(j0 + j) * 32 + lane;
            scol[dst] = t_col[w][lane][j];
            sval[dst] = t_val[w][lane][j];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// y[row] = alpha * sum_j val[row,j] * x[col[row,j]] + g1 z1[row] + g2 z2[row]
// lane = row, warp = slice; software pipelined like the CSR kernel: the stream loads of the
// next batch are in flight while this batch's gathers resolve.
// ---------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(kBlock)
spmv_sell32_kernel(int64_t rows, int64_t nslices, const int64_t* __restrict__ sptr,
                   const int* __restrict__ scol, const double* __restrict__ sval,
                   const double* __restrict__ x, double* __restrict__ y, double alpha,
                   const double* __restrict__ z1, double g1, const double* __restrict__ z2,
                   double g2, const double* __restrict__ ref, double* __restrict__ stat) {
    const int lane = threadIdx.x & 31;
    const int64_t slice = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const bool live = slice < nslices;
    const int64_t s = live ? sptr[slice] : 0;
    const int64_t e = live ? sptr[slice + 1] : 0;
    double a[U];
    int c[U];
    double v[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = 0.0;
    int64_t i = s + lane;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        ok[u] = i + u * 32 < e;
        c[u] = ok[u] ? ld_stream(scol + i + u * 32) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ok[u] ? ld_stream(sval + i + u * 32) : 0.0;
    while (i - lane < e) {  // warp uniform
        double xg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) xg[u] = ok[u] ? __ldg(x + c[u]) : 0.0;
        const int64_t in = i + U * 32;
        int cn[U];
        double vn[U];
        bool okn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            okn[u] = in + u * 32 < e;
            cn[u] = okn[u] ? ld_stream(scol + in + u * 32) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) vn[u] = okn[u] ? ld_stream(sval + in + u * 32) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            a[u] = fma(v[u], xg[u], a[u]);
            c[u] = cn[u];
            v[u] = vn[u];
            ok[u] = okn[u];
        }
        i = in;
    }
    double sum = 0.0;
    if (U == 4) sum = (a[0] + a[1]) + (a[2] + a[3]);
    else
#pragma unroll
        for (int u = 0; u < U; ++u) sum += a[u];
    const int64_t row = slice * 32 + lane;
    double sq = 0.0;
    if (live && row < rows) {
        double out = alpha * sum;
        if (z1) out += g1 * z1[row];
        if (z2) out += g2 * z2[row];
        if (y) y[row] = out;
        if (stat) {
            const double d = ref ? out - ref[row] : out;
            sq = d * d;
        }
    }
    if (stat) {
        __shared__ double s_red[kBlock / 32];
        sq = warp_sum(sq);
        if (lane == 0) s_red[threadIdx.x >> 5] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kBlock / 32; ++w) t += s_red[w];
            stat[blockIdx.x] = t;
        }
    }
}

}  // namespace

// Builds the sliced form when it pays; returns true when m->sell_* are usable.
// sell_state: 0 not examined, 1 built, -1 not eligible / build failed (CSR kernel is used).
bool hg_sell_ready(hg_ctx* ctx, const hg_matrix* cm) {
    hg_matrix* m = const_cast<hg_matrix*>(cm);  // lazily built cache, like unit_row
    if (m->sell_state != 0) return m->sell_state > 0;
    std::lock_guard<std::mutex> lk(hg_matrix_form_mutex());
    if (m->sell_state != 0) return m->sell_state > 0;  // another thread built it meanwhile
    m->sell_state = -1;
    if (m->rows < 64 || m->nnz < 8 * m->rows) return false;  // short rows: TPR<32 CSR kernels do fine
    const int64_t nslices = cdiv(m->rows, 32);
    int32_t* d_w = nullptr;
    if (hg_dmalloc(ctx, &d_w, (size_t)nslices * 4) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    {
        hg_launch_scope scope(ctx, HG_K_SETUP, 8.0 * (double)m->rows);
        slice_width_kernel<<<(unsigned)cdiv(nslices * 32, kBlock), kBlock, 0, ctx->stream>>>(
            m->rows, m->rowptr, d_w);
    }
    std::vector<int32_t> w((size_t)nslices);
    cudaError_t e = cudaMemcpyAsync(w.data(), d_w, (size_t)nslices * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    hg_dfree(d_w);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    std::vector<int64_t> ptr((size_t)nslices + 1);
    int64_t acc = 0;
    for (int64_t s = 0; s < nslices; ++s) {
        ptr[(size_t)s] = acc;
        acc += (int64_t)((w[(size_t)s] + 3) / 4 * 4) * 32;  // 4-column groups (128 entries) never straddle slices
    }
    ptr[(size_t)nslices] = acc;
    const int mode = hg_spmv_mode();
    if (mode != 3) {
        // padding streams extra bytes; the gathers must be cheaper by more than that.  Pixel-driven
        // back-projectors: 360 entries in every row, 1-2 lines per gather against ~16 -> sliced form.
        // Ray-driven projectors: adjacent rays drift apart entry by entry -> row per warp.
        if ((double)acc > 1.10 * (double)m->nnz) return false;
        unsigned long long* d_cnt = nullptr;
        unsigned long long h_cnt[4] = {0, 0, 0, 0};
        if (hg_dmalloc(ctx, &d_cnt, 32) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        const int64_t stride = std::max<int64_t>(1, nslices / 2048);  // ~2048 sampled slices
        const int64_t nwarps = cdiv(nslices, stride);
        cudaError_t g = cudaMemsetAsync(d_cnt, 0, 32, ctx->stream);
        if (g == cudaSuccess) {
            hg_launch_scope scope(ctx, HG_K_SETUP, 12.0 * (double)m->nnz / (double)stride);
            gather_lines_kernel<<<(unsigned)cdiv(nwarps * 32, kBlock), kBlock, 0, ctx->stream>>>(
                m->rows, nslices, stride, m->rowptr, m->colind, d_cnt);
            g = cudaGetLastError();
        }
        if (g == cudaSuccess) g = cudaMemcpyAsync(h_cnt, d_cnt, 32, cudaMemcpyDeviceToHost, ctx->stream);
        if (g == cudaSuccess) g = cudaStreamSynchronize(ctx->stream);
        hg_dfree(d_cnt);
        if (g != cudaSuccess || h_cnt[1] == 0 || h_cnt[3] == 0) {
            cudaGetLastError();
            return false;
        }
        const double lines_sliced = (double)h_cnt[0] / (double)h_cnt[1] * ((double)acc / (double)m->nnz);
        const double lines_rowwarp = (double)h_cnt[2] / (double)h_cnt[3];
        if (lines_sliced > 0.75 * lines_rowwarp) return false;
    }
    cudaError_t a = hg_dmalloc(ctx, &m->sell_ptr, (size_t)(nslices + 1) * 8);
    if (a == cudaSuccess) a = hg_dmalloc(ctx, &m->sell_col, (size_t)(acc + kNnzPad) * 4);
    if (a == cudaSuccess) a = hg_dmalloc(ctx, &m->sell_val, (size_t)(acc + kNnzPad) * 8);
    if (a == cudaSuccess)
        a = cudaMemcpyAsync(m->sell_ptr, ptr.data(), (size_t)(nslices + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (a == cudaSuccess) {
        hg_launch_scope scope(ctx, HG_K_SETUP, 24.0 * (double)m->nnz);
        sell_fill_kernel<<<(unsigned)cdiv(nslices * 32, kFillBlock), kFillBlock, 0, ctx->stream>>>(
            m->rows, nslices, m->rowptr, m->colind, m->vals, m->sell_ptr, m->sell_col, m->sell_val);
        a = cudaGetLastError();
    }
    if (a == cudaSuccess) a = cudaStreamSynchronize(ctx->stream);  // `ptr` is pageable host memory
    if (a != cudaSuccess) {
        cudaGetLastError();
        hg_dfree(m->sell_ptr);
        hg_dfree(m->sell_col);
        hg_dfree(m->sell_val);
        m->sell_ptr = nullptr;
        m->sell_col = nullptr;
        m->sell_val = nullptr;
        return false;  // out of memory: keep the CSR path
    }
    m->sell_slices = nslices;
    m->sell_entries = acc;
    m->sell_state = 1;
    if (hg_idx16_enabled()) hg_sell_compress(ctx, m);  // 16-bit column offsets when every group spans < 65536
    return true;
}

int hg_k_spmv_sell(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y,
                   const hg_spmv_epilogue& ep, double bytes, int* nparts) {
    if (m->sell_col16 || m->sell_col8) return hg_k_spmv_sell16(ctx, m, x, y, ep, bytes, nparts);
    const int64_t grid = cdiv(m->sell_slices, kBlock / 32);
    HG_REQUIRE(grid < (int64_t)2147483647, "spmv: too many rows for one launch");
    if (nparts && ep.stat) *nparts = (int)grid;
    hg_launch_scope scope(ctx, HG_K_SPMV, bytes);
    spmv_sell32_kernel<4><<<(unsigned)grid, kBlock, 0, ctx->stream>>>(
        m->rows, m->sell_slices, m->sell_ptr, m->sell_col, m->sell_val, x, y, ep.alpha, ep.z1, ep.g1,
        ep.z2, ep.g2, ep.ref, ep.stat);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}
