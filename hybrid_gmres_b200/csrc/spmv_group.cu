// Row-group interleaved SpMV for ray-driven projectors (the matrix A of the hot path).
//
// The row-per-warp CSR kernel (kernels.cu) streams A at the HBM rate but is bound by the L1 data pipe of its
// gathers: a warp-wide gather takes 32 consecutive entries of ONE ray, which lie in ~7 different 128-byte lines
// of x (a ray crosses 4-5 pixels of every 4x4 pixel tile) — ncu: L1 wavefronts 66 % of peak at 1.97 GHz, the
// limiter once the 1 kW power cap pulls the SM clock to ~1.55 GHz.  Adjacent rays of a view run through the
// same tiles, so here a warp-wide gather takes E = 32/G consecutive entries of each of G ADJACENT rays: the
// same lines serve G rays.  To keep the matrix stream perfectly coalesced the entries are stored interleaved:
// a group of G rows is cut into rounds of 32 entries, round t holding entries [tE, tE+E) of row 0, then of
// row 1, ... (lane l: row l / E, entry tE + l % E); rows shorter than the longest of their group are padded
// with zero values whose column is the row's last valid one (no extra line).  Same entries, per-row traversal
// order kept; the per-row summation order differs from the CSR kernel (E interleaved partial sums).
//
// 16-bit column stream (default, option "spmv_group16"): lane l of a group visits entries l%E, l%E + E, l%E + 2E ...
// of its ray, E pixels further along the ray each round, so the column index moves by a few image rows per round:
// the stream holds that per-lane DIFFERENCE as int16 (first round: absolute int32 per lane), 10 instead of 12 bytes
// per non-zero.  The decode is one integer add per entry; sums, order and results are bit-identical to the 32-bit
// form.  A matrix with a difference outside int16 keeps the 32-bit form.
#include <algorithm>
#include <vector>

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
__device__ __forceinline__ int ld_stream(const short* p) {
    short v;
    asm volatile("ld.global.nc.L1::no_allocate.s16 %0, [%1];" : "=h"(v) : "l"(p));
    return (int)v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// rounds[g] = ceil(longest row of group g / E)
template <int G>
__global__ void group_rounds_kernel(int64_t rows, int64_t ngroups, const int64_t* __restrict__ rowptr,
                                    int32_t* __restrict__ rounds) {
    constexpr int E = 32 / G;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    int64_t mx = 0;
#pragma unroll
    for (int r = 0; r < G; ++r) {
        const int64_t row = g * G + r;
        if (row < rows) mx = max(mx, rowptr[row + 1] - rowptr[row]);
    }
    rounds[g] = (int32_t)((mx + E - 1) / E);
}

// one warp per group writes its rounds
template <int G>
__global__ void group_fill_kernel(int64_t rows, int64_t ngroups, const int64_t* __restrict__ rowptr,
                                  const int32_t* __restrict__ colind, const double* __restrict__ vals,
                                  const int64_t* __restrict__ gptr, int32_t* __restrict__ gcol,
                                  double* __restrict__ gval) {
    constexpr int E = 32 / G;
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= ngroups) return;
    const int64_t row = g * G + lane / E;
    const int e0 = lane % E;
    const int64_t rs = row < rows ? rowptr[row] : 0;
    const int64_t len = row < rows ? rowptr[row + 1] - rs : 0;
    const int pad_col = len > 0 ? colind[rs + len - 1] : 0;
    const int64_t s = gptr[g], e = gptr[g + 1];
    int64_t t = 0;
    for (int64_t i = s + lane; i < e; i += 32, ++t) {
        const int64_t k = t * E + e0;
        const bool ok = k < len;
        gcol[i] = ok ? colind[rs + k] : pad_col;
        gval[i] = ok ? vals[rs + k] : 0.0;
    }
}

// the same rounds with the column stream as per-lane differences: d16[i] = column(i) - column(i - 32) (0 in the first
// round); col0 holds kChunks checkpoints per group — the lane's column on entering round (c * rounds) / kChunks — so
// that up to kChunks warps can share a group, each starting in the middle of the chain; *overflow is set when a
// difference leaves int16
constexpr int kChunks = 4;
template <int G>
__global__ void group_fill16_kernel(int64_t rows, int64_t ngroups, const int64_t* __restrict__ rowptr,
                                    const int32_t* __restrict__ colind, const double* __restrict__ vals,
                                    const int64_t* __restrict__ gptr, int32_t* __restrict__ col0,
                                    short* __restrict__ d16, double* __restrict__ gval, int* __restrict__ overflow) {
    constexpr int E = 32 / G;
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= ngroups) return;
    const int64_t row = g * G + lane / E;
    const int e0 = lane % E;
    const int64_t rs = row < rows ? rowptr[row] : 0;
    const int64_t len = row < rows ? rowptr[row + 1] - rs : 0;
    const int pad_col = len > 0 ? colind[rs + len - 1] : 0;
    const int64_t s = gptr[g], e = gptr[g + 1];
    const int64_t R = (e - s) >> 5;
    int prev = e0 < len ? colind[rs + e0] : pad_col;
    bool bad = false;
    int64_t t = 0;
    int next_chunk = 0;
    for (int64_t i = s + lane; i < e; i += 32, ++t) {
        while (next_chunk < kChunks && t == (next_chunk * R) / kChunks) {
            col0[(g * kChunks + next_chunk) * 32 + lane] = prev;
            ++next_chunk;
        }
        const int64_t k = t * E + e0;
        const bool ok = k < len;
        const int c = ok ? colind[rs + k] : pad_col;
        const int d = c - prev;
        bad |= d < -32768 || d > 32767;
        d16[i] = (short)d;
        gval[i] = ok ? vals[rs + k] : 0.0;
        prev = c;
    }
    for (; next_chunk < kChunks; ++next_chunk) col0[(g * kChunks + next_chunk) * 32 + lane] = prev;  // empty chunks
    if (bad) atomicExch(overflow, 1);
}

// S warps share a group (S = 1, 2 or 4): warp `part` walks the rounds of its kChunks / S chunks from the chunk's
// checkpoint; the S partial sums of a row are added in a fixed order through shared memory.  Few rows (a rank's
// shard of A on 8 GPUs: 8 326 groups) otherwise leave the machine under-filled with 4-row warps.
template <int G, int S>
__global__ void __launch_bounds__(kBlock)
spmv_group16_kernel(int64_t rows, int64_t ngroups, const int64_t* __restrict__ gptr, const int* __restrict__ col0,
                    const short* __restrict__ d16, const double* __restrict__ gval, const double* __restrict__ x,
                    double* __restrict__ y, double alpha, const double* __restrict__ z1, double g1,
                    const double* __restrict__ z2, double g2, const double* __restrict__ ref,
                    double* __restrict__ stat) {
    constexpr int E = 32 / G;
    constexpr int U = 4;
    constexpr int NW = kBlock / 32;
    static_assert(kChunks % S == 0 && NW % S == 0, "warps per group");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int part = warp % S;
    const int64_t g = (int64_t)blockIdx.x * (NW / S) + warp / S;
    const bool live = g < ngroups;
    const int64_t gs = live ? gptr[g] : 0;
    const int64_t R = live ? (gptr[g + 1] - gs) >> 5 : 0;
    const int c0 = part * (kChunks / S);
    const int64_t s = gs + ((c0 * R) / kChunks) * 32;
    const int64_t e = gs + (((c0 + kChunks / S) * R) / kChunks) * 32;
    // same software pipeline as spmv_group_kernel; the running column of the lane is carried through the rounds
    double a[U];
    int d[U];
    double v[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = 0.0;
    int col = live ? col0[(g * kChunks + c0) * 32 + lane] : 0;
    int64_t i = s + lane;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        ok[u] = i + u * 32 < e;
        d[u] = ok[u] ? ld_stream(d16 + i + u * 32) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ok[u] ? ld_stream(gval + i + u * 32) : 0.0;
    while (i - lane < e) {  // warp uniform
        double xg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            col += d[u];  // 0 past the end of the chunk: the column stays on its last line
            xg[u] = ok[u] ? __ldg(x + col) : 0.0;
        }
        const int64_t in = i + U * 32;
        int dn[U];
        double vn[U];
        bool okn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            okn[u] = in + u * 32 < e;
            dn[u] = okn[u] ? ld_stream(d16 + in + u * 32) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) vn[u] = okn[u] ? ld_stream(gval + in + u * 32) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            a[u] = fma(v[u], xg[u], a[u]);
            d[u] = dn[u];
            v[u] = vn[u];
            ok[u] = okn[u];
        }
        i = in;
    }
    double sum = (a[0] + a[1]) + (a[2] + a[3]);
#pragma unroll
    for (int o = E / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);  // the E lanes of a row
    if (S > 1) {
        __shared__ double s_part[NW][32];
        s_part[warp][lane] = sum;
        __syncthreads();
        if (part == 0) {
            sum = 0.0;
#pragma unroll
            for (int t = 0; t < S; ++t) sum += s_part[warp + t][lane];
        }
    }
    double sq = 0.0;
    const int64_t row = g * G + lane / E;
    if (live && part == 0 && (lane % E) == 0 && row < rows) {
        double out = alpha * sum;
        if (z1) out += g1 * z1[row];
        if (z2) out += g2 * z2[row];
        if (y) y[row] = out;
        if (stat) {
            const double dd = ref ? out - ref[row] : out;
            sq = dd * dd;
        }
    }
    if (stat) {
        __shared__ double s_red[NW];
        sq = warp_sum(sq);
        if (lane == 0) s_red[warp] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) t += s_red[w];
            stat[blockIdx.x] = t;
        }
    }
}

template <int G>
__global__ void __launch_bounds__(kBlock)
spmv_group_kernel(int64_t rows, int64_t ngroups, const int64_t* __restrict__ gptr, const int* __restrict__ gcol,
                  const double* __restrict__ gval, const double* __restrict__ x, double* __restrict__ y,
                  double alpha, const double* __restrict__ z1, double g1, const double* __restrict__ z2, double g2,
                  const double* __restrict__ ref, double* __restrict__ stat) {
    constexpr int E = 32 / G;
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const bool live = g < ngroups;
    const int64_t s = live ? gptr[g] : 0;
    const int64_t e = live ? gptr[g + 1] : 0;
    // software-pipelined batches of U rounds (U * 32 entries), as in spmv_csr_kernel: the (col, val) loads of the
    // next batch are in flight while this batch's dependent gathers resolve
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
        c[u] = ok[u] ? ld_stream(gcol + i + u * 32) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ok[u] ? ld_stream(gval + i + u * 32) : 0.0;
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
            cn[u] = okn[u] ? ld_stream(gcol + in + u * 32) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) vn[u] = okn[u] ? ld_stream(gval + in + u * 32) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            a[u] = fma(v[u], xg[u], a[u]);
            c[u] = cn[u];
            v[u] = vn[u];
            ok[u] = okn[u];
        }
        i = in;
    }
    double sum = (a[0] + a[1]) + (a[2] + a[3]);
#pragma unroll
    for (int o = E / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);  // the E lanes of a row
    double sq = 0.0;
    const int64_t row = g * G + lane / E;
    if (live && (lane % E) == 0 && row < rows) {
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

int g_group = -1;
bool g_group_explicit = false;  // set by the caller (option / env): applies to every eligible matrix

template <int G>
bool build(hg_ctx* ctx, hg_matrix* m) {
    const int64_t ngroups = cdiv(m->rows, G);
    int32_t* d_r = nullptr;
    if (hg_dmalloc(ctx, &d_r, (size_t)ngroups * 4) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    {
        hg_launch_scope scope(ctx, HG_K_SETUP, 8.0 * (double)m->rows);
        group_rounds_kernel<G><<<(unsigned)cdiv(ngroups, kBlock), kBlock, 0, ctx->stream>>>(m->rows, ngroups, m->rowptr, d_r);
    }
    std::vector<int32_t> r((size_t)ngroups);
    cudaError_t e = cudaMemcpyAsync(r.data(), d_r, (size_t)ngroups * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    hg_dfree(d_r);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    std::vector<int64_t> ptr((size_t)ngroups + 1);
    int64_t acc = 0;
    for (int64_t g = 0; g < ngroups; ++g) {
        ptr[(size_t)g] = acc;
        acc += (int64_t)r[(size_t)g] * 32;
    }
    ptr[(size_t)ngroups] = acc;
    if ((double)acc > 1.06 * (double)m->nnz) return false;  // ragged groups: the padding would cost more than the gathers save
    cudaError_t a = hg_dmalloc(ctx, &m->grp_ptr, (size_t)(ngroups + 1) * 8);
    if (a == cudaSuccess) a = hg_dmalloc(ctx, &m->grp_val, (size_t)(acc + kNnzPad) * 8);
    if (a == cudaSuccess)
        a = cudaMemcpyAsync(m->grp_ptr, ptr.data(), (size_t)(ngroups + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
    bool have16 = false;
    if (a == cudaSuccess && hg_spmv_group16()) {
        // 16-bit column differences; a difference outside int16 anywhere sends the matrix to the 32-bit form
        int* d_flag = nullptr;
        int flag = 1;
        cudaError_t b = hg_dmalloc(ctx, &d_flag, sizeof(int));
        if (b == cudaSuccess) b = cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream);
        if (b == cudaSuccess) b = hg_dmalloc(ctx, &m->grp_col0, (size_t)ngroups * kChunks * 32 * 4);
        if (b == cudaSuccess) b = hg_dmalloc(ctx, &m->grp_d16, (size_t)(acc + kNnzPad) * 2);
        if (b == cudaSuccess) {
            hg_launch_scope scope(ctx, HG_K_SETUP, 22.0 * (double)m->nnz);
            group_fill16_kernel<G><<<(unsigned)cdiv(ngroups * 32, kBlock), kBlock, 0, ctx->stream>>>(
                m->rows, ngroups, m->rowptr, m->colind, m->vals, m->grp_ptr, m->grp_col0, m->grp_d16, m->grp_val, d_flag);
            b = cudaGetLastError();
        }
        if (b == cudaSuccess) b = cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (b == cudaSuccess) b = cudaStreamSynchronize(ctx->stream);
        if (d_flag) hg_dfree(d_flag);
        have16 = b == cudaSuccess && flag == 0;
        if (!have16) {
            cudaGetLastError();
            hg_dfree(m->grp_col0);
            hg_dfree(m->grp_d16);
            m->grp_col0 = nullptr;
            m->grp_d16 = nullptr;
        }
    }
    if (a == cudaSuccess && !have16) {
        a = hg_dmalloc(ctx, &m->grp_col, (size_t)(acc + kNnzPad) * 4);
        if (a == cudaSuccess) {
            hg_launch_scope scope(ctx, HG_K_SETUP, 24.0 * (double)m->nnz);
            group_fill_kernel<G><<<(unsigned)cdiv(ngroups * 32, kBlock), kBlock, 0, ctx->stream>>>(
                m->rows, ngroups, m->rowptr, m->colind, m->vals, m->grp_ptr, m->grp_col, m->grp_val);
            a = cudaGetLastError();
        }
    }
    if (a == cudaSuccess) a = cudaStreamSynchronize(ctx->stream);  // `ptr` is pageable host memory
    if (a != cudaSuccess) {
        cudaGetLastError();
        hg_dfree(m->grp_ptr);
        hg_dfree(m->grp_col);
        hg_dfree(m->grp_val);
        hg_dfree(m->grp_col0);
        hg_dfree(m->grp_d16);
        m->grp_ptr = nullptr;
        m->grp_col = nullptr;
        m->grp_val = nullptr;
        m->grp_col0 = nullptr;
        m->grp_d16 = nullptr;
        return false;
    }
    m->grp_groups = ngroups;
    m->grp_entries = acc;
    m->grp_G = G;
    return true;
}

}  // namespace

// 16-bit column differences in the interleaved form: option "spmv_group16" / env HG_SPMV_GROUP16 (default 1)
static int g_group16 = -1;
int hg_spmv_group16() {
    if (g_group16 < 0) {
        const char* e = getenv("HG_SPMV_GROUP16");
        g_group16 = (e && e[0] == '0') ? 0 : 1;
    }
    return g_group16;
}
void hg_spmv_group16_set(int v) { g_group16 = v != 0 ? 1 : 0; }
// warps per group of the 16-bit form: option "spmv_group_split" / env HG_SPMV_GROUP_SPLIT = 1, 2, 4; 0 (default): by size
static int g_group_split = -1;
int hg_spmv_group_split() {
    if (g_group_split < 0) {
        const char* e = getenv("HG_SPMV_GROUP_SPLIT");
        const int v = e ? atoi(e) : 0;
        g_group_split = (v == 1 || v == 2 || v == 4) ? v : 0;
    }
    return g_group_split;
}
void hg_spmv_group_split_set(int v) { g_group_split = (v == 1 || v == 2 || v == 4) ? v : 0; }
// smallest matrix that takes the row-group form by default: option "spmv_group_min_rows" / env HG_SPMV_GROUP_MIN_ROWS
static int64_t g_group_min_rows = -1;
int64_t hg_spmv_group_min_rows() {
    if (g_group_min_rows < 0) {
        const char* e = getenv("HG_SPMV_GROUP_MIN_ROWS");
        g_group_min_rows = e ? atoll(e) : 16384;
    }
    return g_group_min_rows;
}
void hg_spmv_group_min_rows_set(int v) { g_group_min_rows = v < 0 ? 16384 : v; }

// rows per group of the interleaved form: option "spmv_group" / env HG_SPMV_GROUP = 0 (off), 2, 4 or 8
int hg_spmv_group() {
    if (g_group < 0) {
        const char* e = getenv("HG_SPMV_GROUP");
        const int v = e ? atoi(e) : HG_SPMV_GROUP_DEFAULT;
        g_group_explicit = e != nullptr;
        g_group = (v == 2 || v == 4 || v == 8) ? v : 0;
    }
    return g_group;
}
void hg_spmv_group_set(int v) {
    if (v < 0) {  // back to the built-in default (large matrices only)
        g_group = HG_SPMV_GROUP_DEFAULT;
        g_group_explicit = false;
        return;
    }
    g_group = (v == 2 || v == 4 || v == 8) ? v : 0;
    g_group_explicit = true;
}

// Lazily builds the interleaved copy for long-row matrices that run the row-per-warp kernel.
bool hg_group_ready(hg_ctx* ctx, const hg_matrix* cm) {
    hg_matrix* m = const_cast<hg_matrix*>(cm);
    const int G = hg_spmv_group();
    if (G == 0) return false;
    if (m->grp_state != 0) return m->grp_state > 0 && m->grp_G == G;
    std::lock_guard<std::mutex> lk(hg_matrix_form_mutex());
    if (m->grp_state != 0) return m->grp_state > 0 && m->grp_G == G;
    m->grp_state = -1;
    // rows of >= 128 entries on average.  Small matrices run with several warps per group (hg_k_spmv_group); below
    // ~16 000 rows even that under-fills the machine (a rank's 8 326-row shard of the 256^2 problem: 15.4 vs 15.0 us)
    if (m->rows < (g_group_explicit ? 1024 : hg_spmv_group_min_rows()) || m->nnz < 128 * m->rows) return false;
    bool ok = false;
    if (G == 2) ok = build<2>(ctx, m);
    else if (G == 4) ok = build<4>(ctx, m);
    else ok = build<8>(ctx, m);
    if (ok) m->grp_state = 1;
    return ok;
}

void hg_group_free(hg_matrix* m) {
    hg_dfree(m->grp_ptr);
    hg_dfree(m->grp_col);
    hg_dfree(m->grp_val);
    hg_dfree(m->grp_col0);
    hg_dfree(m->grp_d16);
    m->grp_ptr = nullptr;
    m->grp_col = nullptr;
    m->grp_val = nullptr;
    m->grp_col0 = nullptr;
    m->grp_d16 = nullptr;
    m->grp_state = 0;
}

int hg_k_spmv_group(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y, const hg_spmv_epilogue& ep,
                    double bytes, int* nparts) {
    const int64_t grid = cdiv(m->grp_groups, kBlock / 32);
    HG_REQUIRE(grid < (int64_t)2147483647, "spmv: too many rows for one launch");
    if (nparts && ep.stat) *nparts = (int)grid;
    hg_launch_scope scope(ctx, HG_K_SPMV, bytes);
    if (m->grp_d16) {
        // several waves of warps: with fewer groups than ~4 x 48 warps per SM, S warps share a group
        int S = hg_spmv_group_split();
        if (S <= 0) {
            const int64_t target = (int64_t)ctx->sm_count * 48 * 4;
            S = 1;
            while (S < kChunks && m->grp_groups * S < target) S *= 2;
        }
        while (S > m->grp_G) S /= 2;  // at most one warp per row: the grid (= residual partials) stays <= rows / 8
        const int64_t grid16 = cdiv(m->grp_groups * S, kBlock / 32);
        if (nparts && ep.stat) *nparts = (int)grid16;
#define HG_GRP16_ARGS m->rows, m->grp_groups, m->grp_ptr, m->grp_col0, m->grp_d16, m->grp_val, x, y, ep.alpha, ep.z1, ep.g1, \
                      ep.z2, ep.g2, ep.ref, ep.stat
#define HG_GRP16_LAUNCH(G)                                                                                          \
    do {                                                                                                            \
        if (S == 1) spmv_group16_kernel<G, 1><<<(unsigned)grid16, kBlock, 0, ctx->stream>>>(HG_GRP16_ARGS);         \
        else if (S == 2) spmv_group16_kernel<G, 2><<<(unsigned)grid16, kBlock, 0, ctx->stream>>>(HG_GRP16_ARGS);    \
        else spmv_group16_kernel<G, 4><<<(unsigned)grid16, kBlock, 0, ctx->stream>>>(HG_GRP16_ARGS);                \
    } while (0)
        if (m->grp_G == 2) HG_GRP16_LAUNCH(2);
        else if (m->grp_G == 4) HG_GRP16_LAUNCH(4);
        else HG_GRP16_LAUNCH(8);
#undef HG_GRP16_LAUNCH
#undef HG_GRP16_ARGS
        HG_CUDA(cudaGetLastError());
        return HG_OK;
    }
#define HG_GRP_ARGS m->rows, m->grp_groups, m->grp_ptr, m->grp_col, m->grp_val, x, y, ep.alpha, ep.z1, ep.g1, ep.z2, ep.g2, \
                    ep.ref, ep.stat
    if (m->grp_G == 2) spmv_group_kernel<2><<<(unsigned)grid, kBlock, 0, ctx->stream>>>(HG_GRP_ARGS);
    else if (m->grp_G == 4) spmv_group_kernel<4><<<(unsigned)grid, kBlock, 0, ctx->stream>>>(HG_GRP_ARGS);
    else spmv_group_kernel<8><<<(unsigned)grid, kBlock, 0, ctx->stream>>>(HG_GRP_ARGS);
#undef HG_GRP_ARGS
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}
