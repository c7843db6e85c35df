// 16-bit column offsets for the SpMV streams.
//
// Both SpMV kernels are HBM-bound on the matrix stream (12 B per entry: 8 value + 4 column index;
// ncu: DRAM traffic = algorithmic bytes, profiles/r01_spmv_sell_tile.md), so the only way to make
// them faster is to move fewer bytes.  The entries a warp consumes together are close in column
// space — 128 consecutive entries of a slice of the back-projector are two views of 32 adjacent
// pixels, 32 consecutive entries of a ray are ~25 pixels of path — so their columns are stored as
// 16-bit offsets from a per-group 32-bit base: 10.03 (sliced form) / 10.13 (CSR) bytes per entry.
// The values, the entry order and the arithmetic are those of the 32-bit kernels (bit-identical
// results); a matrix with any group spanning >= 65536 columns keeps its 32-bit indices.
//
// 8-bit offsets (sliced form, default when they fit; option "spmv_idx8"): the 32 entries of ONE column of a slice
// are the same view of 32 adjacent pixels — a dozen detector bins — so with a base per slice column the offsets
// fit a byte: 1 + 4/32 = 1.125 index bytes per entry instead of 2.03 (9.1 against 10.03 bytes per entry).  The
// four offsets a lane needs for a batch of 128 entries are stored together and arrive as one 32-bit load, the
// four bases as one 16-byte broadcast load.  Entry order and arithmetic unchanged: bit-identical results.
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
__device__ __forceinline__ int ld_stream(const uint16_t* p) {
    unsigned short v;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return (int)v;
}
__device__ __forceinline__ unsigned ld_stream(const unsigned* p) {
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// absolute columns of the four entries lane `lane` owns in the batch of 128 entries that starts at i0
// (i0 multiple of 128): entry u is i0 + u*32 + lane.  I8: byte offsets stored [lane][u] + one base per slice
// column; else 16-bit offsets in entry order + one base per batch.
template <bool I8>
__device__ __forceinline__ void load_cols(const void* __restrict__ scol, const int* __restrict__ sbase, int64_t i0,
                                          int lane, bool ok, int (&c)[4]) {
    if (I8) {
        const unsigned w = ok ? ld_stream(reinterpret_cast<const unsigned*>(static_cast<const uint8_t*>(scol) + i0) + lane) : 0u;
        const int4 b = ok ? __ldg(reinterpret_cast<const int4*>(sbase) + (i0 >> 7)) : make_int4(0, 0, 0, 0);
        c[0] = b.x + (int)(w & 255u);
        c[1] = b.y + (int)((w >> 8) & 255u);
        c[2] = b.z + (int)((w >> 16) & 255u);
        c[3] = b.w + (int)(w >> 24);
    } else {
        const uint16_t* p = static_cast<const uint16_t*>(scol) + i0 + lane;
        const int b = ok ? __ldg(sbase + (i0 >> 7)) : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = b + (ok ? ld_stream(p + u * 32) : 0);
    }
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- builders ---------------------------------------------------------------
// sliced form: one warp per group of 128 consecutive entries
__global__ void __launch_bounds__(kBlock)
sell_compress_kernel(int64_t ngroups, const int32_t* __restrict__ col, uint16_t* __restrict__ col16,
                     int32_t* __restrict__ base, int* __restrict__ too_wide) {
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= ngroups) return;
    const int32_t* p = col + g * 128 + lane;
    const int c0 = p[0], c1 = p[32], c2 = p[64], c3 = p[96];
    const int lo = warp_min(min(min(c0, c1), min(c2, c3)));
    const int hi = warp_max(max(max(c0, c1), max(c2, c3)));
    if (hi - lo > 65535) {
        if (lane == 0) atomicOr(too_wide, 1);
        return;
    }
    uint16_t* q = col16 + g * 128 + lane;
    q[0] = (uint16_t)(c0 - lo);
    q[32] = (uint16_t)(c1 - lo);
    q[64] = (uint16_t)(c2 - lo);
    q[96] = (uint16_t)(c3 - lo);
    if (lane == 0) base[g] = lo;
}

// sliced form, 8-bit: one warp per group of 128 entries; a base per slice column (32 entries), offsets stored
// [lane][u] so that a lane's four offsets are one 32-bit word
__global__ void __launch_bounds__(kBlock)
sell_compress8_kernel(int64_t ngroups, const int32_t* __restrict__ col, uint8_t* __restrict__ col8,
                      int32_t* __restrict__ base, int* __restrict__ too_wide) {
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= ngroups) return;
    const int32_t* p = col + g * 128 + lane;
    const int c0 = p[0], c1 = p[32], c2 = p[64], c3 = p[96];
    const int l0 = warp_min(c0), l1 = warp_min(c1), l2 = warp_min(c2), l3 = warp_min(c3);
    const int w = max(max(warp_max(c0) - l0, warp_max(c1) - l1), max(warp_max(c2) - l2, warp_max(c3) - l3));
    if (w > 255) {
        if (lane == 0) atomicOr(too_wide, 1);
        return;
    }
    reinterpret_cast<unsigned*>(col8 + g * 128)[lane] =
        (unsigned)(c0 - l0) | ((unsigned)(c1 - l1) << 8) | ((unsigned)(c2 - l2) << 16) | ((unsigned)(c3 - l3) << 24);
    if (lane == 0) reinterpret_cast<int4*>(base)[g] = make_int4(l0, l1, l2, l3);
}

// CSR: one warp per row, a group is 32 consecutive entries of the row
__global__ void __launch_bounds__(kBlock)
csr_compress_kernel(int64_t rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                    const int64_t* __restrict__ gptr, uint16_t* __restrict__ col16,
                    int32_t* __restrict__ base, int* __restrict__ too_wide) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    int64_t g = gptr[row];
    for (int64_t i0 = s; i0 < e; i0 += 32, ++g) {
        const bool ok = i0 + lane < e;
        const int c = ok ? col[i0 + lane] : 0;
        const int lo = warp_min(ok ? c : 0x7fffffff);
        const int hi = warp_max(ok ? c : -1);
        if (hi - lo > 65535) {
            if (lane == 0) atomicOr(too_wide, 1);
            return;
        }
        if (ok) col16[i0 + lane] = (uint16_t)(c - lo);
        if (lane == 0) base[g] = lo;
    }
}

// ---- kernels: the 32-bit kernels with `base + offset` in place of the index load -------------
template <int U, bool I8>
__global__ void __launch_bounds__(kBlock)
spmv_sell16_kernel(int64_t rows, int64_t nslices, const int64_t* __restrict__ sptr,
                   const void* __restrict__ scol, const int* __restrict__ sbase,
                   const double* __restrict__ sval, const double* __restrict__ x, double* __restrict__ y,
                   double alpha, const double* __restrict__ z1, double g1, const double* __restrict__ z2,
                   double g2, const double* __restrict__ ref, double* __restrict__ stat) {
    static_assert(U == 4, "a group of the sliced form is 4 columns");
    const int lane = threadIdx.x & 31;
    const int64_t slice = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const bool live = slice < nslices;
    const int64_t s = live ? sptr[slice] : 0;  // multiples of 128
    const int64_t e = live ? sptr[slice + 1] : 0;
    double a[U];
    int c[U];
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = 0.0;
    int64_t i = s + lane;
    bool ok = i < e;  // whole groups: all U entries of a batch exist or none
    load_cols<I8>(scol, sbase, s, lane, ok, c);
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ok ? ld_stream(sval + i + u * 32) : 0.0;
    while (i - lane < e) {  // warp uniform
        double xg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) xg[u] = __ldg(x + c[u]);
        const int64_t in = i + U * 32;
        const bool okn = in < e;
        int cn[U];
        double vn[U];
        load_cols<I8>(scol, sbase, in - lane, lane, okn, cn);
#pragma unroll
        for (int u = 0; u < U; ++u) vn[u] = okn ? ld_stream(sval + in + u * 32) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            a[u] = fma(v[u], xg[u], a[u]);
            c[u] = cn[u];
            v[u] = vn[u];
        }
        i = in;
    }
    const double sum = (a[0] + a[1]) + (a[2] + a[3]);
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

// Few rows (n/32 slices < ~32 warps per SM): S warps share a slice, each walks every S-th batch of
// 128 entries, the S partial sums of a row are combined through shared memory in a fixed order
// (deterministic).  A CTA holds kBlock/32/S slices.  Same per-entry arithmetic as the kernel above;
// the per-row summation order differs (S interleaved partial sums).
template <int S, bool I8>
__global__ void __launch_bounds__(S > 8 ? S * 32 : kBlock)
spmv_sell16_split_kernel(int64_t rows, int64_t nslices, const int64_t* __restrict__ sptr,
                         const void* __restrict__ scol, const int* __restrict__ sbase,
                         const double* __restrict__ sval, const double* __restrict__ x,
                         double* __restrict__ y, double alpha, const double* __restrict__ z1, double g1,
                         const double* __restrict__ z2, double g2, const double* __restrict__ ref,
                         double* __restrict__ stat) {
    constexpr int U = 4;
    constexpr int NW = S > 8 ? S : kBlock / 32;  // warps per CTA
    constexpr int SPB = NW / S;                  // slices per CTA
    __shared__ double s_part[NW][32];
    __shared__ double s_red[NW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int part = warp % S, ls = warp / S;
    const int64_t slice = (int64_t)blockIdx.x * SPB + ls;
    const bool live = slice < nslices;
    const int64_t s = live ? sptr[slice] : 0;
    const int64_t e = live ? sptr[slice + 1] : 0;
    double a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = 0.0;
    // software pipelined like spmv_sell16_kernel: the (offset, value) loads of this warp's next batch are in
    // flight while the gathers of the current one resolve (a warp has only a handful of batches here, so an
    // un-pipelined loop exposes two dependent latencies per batch: ncu long-scoreboard 47 at 256^2)
    constexpr int64_t STEP = (int64_t)S * (U * 32);
    int64_t i = s + (int64_t)part * (U * 32) + lane;
    bool ok = i - lane < e;  // whole groups: all U entries of a batch exist or none
    int c[U];
    double v[U];
    load_cols<I8>(scol, sbase, i - lane, lane, ok, c);
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ok ? ld_stream(sval + i + u * 32) : 0.0;
    while (ok) {  // warp uniform
        double xg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) xg[u] = __ldg(x + c[u]);
        const int64_t in = i + STEP;
        const bool okn = in - lane < e;
        int cn[U];
        double vn[U];
        load_cols<I8>(scol, sbase, in - lane, lane, okn, cn);
#pragma unroll
        for (int u = 0; u < U; ++u) vn[u] = okn ? ld_stream(sval + in + u * 32) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            a[u] = fma(v[u], xg[u], a[u]);
            c[u] = cn[u];
            v[u] = vn[u];
        }
        i = in;
        ok = okn;
    }
    s_part[warp][lane] = (a[0] + a[1]) + (a[2] + a[3]);
    __syncthreads();
    double sq = 0.0;
    if (part == 0) {
        double sum = 0.0;
#pragma unroll
        for (int t = 0; t < S; ++t) sum += s_part[ls * S + t][lane];
        const int64_t row = slice * 32 + lane;
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
    }
    if (stat) {
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

__global__ void __launch_bounds__(kBlock)
spmv_csr16_kernel(int64_t rows, const int64_t* __restrict__ rowptr, const int64_t* __restrict__ gptr,
                  const uint16_t* __restrict__ col16, const int* __restrict__ base,
                  const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
                  double alpha, const double* __restrict__ z1, double g1, const double* __restrict__ z2,
                  double g2, const double* __restrict__ ref, double* __restrict__ stat) {
    constexpr int U = 4;
    constexpr int RPB = kBlock / 32;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * RPB + (threadIdx.x >> 5);
    const bool valid = row < rows;
    const int64_t s = valid ? rowptr[row] : 0;
    const int64_t e = valid ? rowptr[row + 1] : 0;
    const int* bp = base + (valid ? gptr[row] : 0);  // group t of this row: entries [s+32t, s+32t+32)
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
        c[u] = ok[u] ? __ldg(bp + u) + ld_stream(col16 + i + u * 32) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ok[u] ? ld_stream(vals + i + u * 32) : 0.0;
    while (i - lane < e) {  // warp uniform
        double xg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) xg[u] = ok[u] ? __ldg(x + c[u]) : 0.0;
        const int64_t in = i + U * 32;
        bp += U;
        int cn[U];
        double vn[U];
        bool okn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            okn[u] = in + u * 32 < e;
            cn[u] = okn[u] ? __ldg(bp + u) + ld_stream(col16 + in + u * 32) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) vn[u] = okn[u] ? ld_stream(vals + in + u * 32) : 0.0;
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
    sum = warp_sum(sum);
    double sq = 0.0;
    if (valid && lane == 0) {
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

int g_idx16 = -1;

}  // namespace

// 0: off; 1 (default): the sliced form only — measured on the 1024^2 workload: B 683 -> 609 us per launch;
// 2: also the row-per-warp CSR kernel — its gathers keep the L1 data pipe busier than the stream keeps HBM
// (A: 451.6 -> 449.1 us), so the narrower indices buy nothing there and stay opt-in.
static int idx16_level() {
    if (g_idx16 < 0) {
        const char* e = getenv("HG_IDX16");
        g_idx16 = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
    }
    return g_idx16;
}
bool hg_idx16_enabled() { return idx16_level() >= 1; }
bool hg_idx16_csr_enabled() { return idx16_level() >= 2; }
void hg_idx16_set(int v) { g_idx16 = v < 0 ? 0 : (v > 2 ? 2 : v); }

// byte offsets in the sliced form when every slice column spans < 256: option "spmv_idx8" / env HG_IDX8.
// Default (-1): matrices of >= 4096 slices (131 072 rows) — measured B 595 -> 528 us at 1024^2, 162 -> 156 us at
// 512^2, but 51 -> 55 us at 256^2, where 16 warps share a slice and each sees only ~6 batches; 1: every
// matrix that fits; 0: never.
static int g_idx8 = -2;
static int idx8_level() {
    if (g_idx8 == -2) {
        const char* e = getenv("HG_IDX8");
        g_idx8 = (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1;
    }
    return g_idx8;
}
bool hg_idx8_wanted(const hg_matrix* m) {
    const int l = idx8_level();
    return l == 1 || (l < 0 && m->sell_slices >= 4096);
}
void hg_idx8_set(int v) { g_idx8 = v < 0 ? -1 : (v != 0 ? 1 : 0); }

void hg_idx16_free(hg_matrix* m) {
    hg_dfree(m->sell_col16);
    hg_dfree(m->sell_col8);
    hg_dfree(m->sell_base);
    hg_dfree(m->csr_col16);
    hg_dfree(m->csr_base);
    hg_dfree(m->csr_gptr);
    m->sell_col16 = nullptr;
    m->sell_col8 = nullptr;
    m->sell_base = nullptr;
    m->csr_col16 = nullptr;
    m->csr_base = nullptr;
    m->csr_gptr = nullptr;
}

// Replaces the 32-bit column array of the sliced copy by 8-bit offsets (a base per slice column) when every
// slice column allows it, else by 16-bit offsets (a base per 128 entries) when every group allows that.
static bool sell_compress_try(hg_ctx* ctx, hg_matrix* m, bool bytes8) {
    const int64_t ng = m->sell_entries / 128;
    int* d_flag = nullptr;
    int h_flag = 1;
    cudaError_t e = bytes8 ? hg_dmalloc(ctx, &m->sell_col8, (size_t)(m->sell_entries + kNnzPad))
                           : hg_dmalloc(ctx, &m->sell_col16, (size_t)(m->sell_entries + kNnzPad) * 2);
    if (e == cudaSuccess) e = hg_dmalloc(ctx, &m->sell_base, (size_t)(ng + 1) * (bytes8 ? 16 : 4));
    if (e == cudaSuccess) e = hg_dmalloc(ctx, &d_flag, 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_flag, 0, 4, ctx->stream);
    if (e == cudaSuccess) {
        hg_launch_scope scope(ctx, HG_K_SETUP, (bytes8 ? 5.0 : 6.0) * (double)m->sell_entries);
        if (bytes8)
            sell_compress8_kernel<<<(unsigned)cdiv(ng * 32, kBlock), kBlock, 0, ctx->stream>>>(
                ng, m->sell_col, m->sell_col8, m->sell_base, d_flag);
        else
            sell_compress_kernel<<<(unsigned)cdiv(ng * 32, kBlock), kBlock, 0, ctx->stream>>>(
                ng, m->sell_col, m->sell_col16, m->sell_base, d_flag);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (d_flag) hg_dfree(d_flag);
    if (e != cudaSuccess || h_flag != 0) {  // a group is too wide (or no memory): keep what we had
        cudaGetLastError();
        hg_dfree(m->sell_col16);
        hg_dfree(m->sell_col8);
        hg_dfree(m->sell_base);
        m->sell_col16 = nullptr;
        m->sell_col8 = nullptr;
        m->sell_base = nullptr;
        return false;
    }
    return true;
}

void hg_sell_compress(hg_ctx* ctx, hg_matrix* m) {
    if (m->sell_state <= 0 || m->sell_col16 || m->sell_col8 || m->sell_entries == 0) return;
    if (!(hg_idx8_wanted(m) && sell_compress_try(ctx, m, true)) && !sell_compress_try(ctx, m, false)) return;
    hg_dfree(m->sell_col);  // the 32-bit copy of the sliced indices is no longer read
    m->sell_col = nullptr;
}

bool hg_csr16_ready(hg_ctx* ctx, const hg_matrix* cm) {
    hg_matrix* m = const_cast<hg_matrix*>(cm);  // lazily built cache
    if (m->csr16_state != 0) return m->csr16_state > 0;
    std::lock_guard<std::mutex> lk(hg_matrix_form_mutex());
    if (m->csr16_state != 0) return m->csr16_state > 0;
    m->csr16_state = -1;
    if (m->rows == 0 || m->nnz == 0) return false;
    std::vector<int64_t> ptr((size_t)m->rows + 1), gptr((size_t)m->rows + 1);
    cudaError_t e = cudaMemcpyAsync(ptr.data(), m->rowptr, (size_t)(m->rows + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    int64_t acc = 0;
    for (int64_t r = 0; r < m->rows; ++r) {
        gptr[(size_t)r] = acc;
        acc += (ptr[(size_t)r + 1] - ptr[(size_t)r] + 31) / 32;
    }
    gptr[(size_t)m->rows] = acc;
    int* d_flag = nullptr;
    int h_flag = 1;
    e = hg_dmalloc(ctx, &m->csr_col16, (size_t)(m->nnz + kNnzPad) * 2);
    if (e == cudaSuccess) e = hg_dmalloc(ctx, &m->csr_base, (size_t)(acc + 8) * 4);
    if (e == cudaSuccess) e = hg_dmalloc(ctx, &m->csr_gptr, (size_t)(m->rows + 1) * 8);
    if (e == cudaSuccess) e = hg_dmalloc(ctx, &d_flag, 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_flag, 0, 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->csr_base, 0, (size_t)(acc + 8) * 4, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(m->csr_gptr, gptr.data(), (size_t)(m->rows + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        hg_launch_scope scope(ctx, HG_K_SETUP, 6.0 * (double)m->nnz);
        csr_compress_kernel<<<(unsigned)cdiv(m->rows * 32, kBlock), kBlock, 0, ctx->stream>>>(
            m->rows, m->rowptr, m->colind, m->csr_gptr, m->csr_col16, m->csr_base, d_flag);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (d_flag) hg_dfree(d_flag);
    if (e != cudaSuccess || h_flag != 0) {
        cudaGetLastError();
        hg_dfree(m->csr_col16);
        hg_dfree(m->csr_base);
        hg_dfree(m->csr_gptr);
        m->csr_col16 = nullptr;
        m->csr_base = nullptr;
        m->csr_gptr = nullptr;
        return false;
    }
    m->csr_groups = acc;
    m->csr16_state = 1;
    return true;
}

int hg_k_spmv_sell16(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y,
                     const hg_spmv_epilogue& ep, double bytes, int* nparts) {
    // One warp per slice wants several waves of warps (the kernel keeps ~48 warps per SM resident): with
    // fewer slices the last, partly filled wave runs latency-bound on a few warps per SM (ncu, 256^2 / 512^2:
    // 1.15 waves, 3.5 / 5.0 TB/s).  S warps then share a slice so that there are >= 4 waves of warp tasks.
    const int64_t target = (int64_t)ctx->sm_count * 48 * 4;
    int S = 1;
    while (S < 16 && m->sell_slices * S < target && m->sell_entries / (m->sell_slices * S * 2) >= 128) S *= 2;
    const int64_t grid = S > 8 ? m->sell_slices : cdiv(m->sell_slices * S, kBlock / 32);
    HG_REQUIRE(grid < (int64_t)2147483647, "spmv: too many rows for one launch");
    if (nparts && ep.stat) *nparts = (int)grid;
    hg_launch_scope scope(ctx, HG_K_SPMV, bytes);
#define HG_SELL16_ARGS m->rows, m->sell_slices, m->sell_ptr, scol, m->sell_base, m->sell_val, x, y, ep.alpha, ep.z1, ep.g1, \
                       ep.z2, ep.g2, ep.ref, ep.stat
#define HG_SELL16_LAUNCH(I8)                                                                                      \
    do {                                                                                                          \
        if (S == 1) spmv_sell16_kernel<4, I8><<<(unsigned)grid, kBlock, 0, ctx->stream>>>(HG_SELL16_ARGS);        \
        else if (S == 2) spmv_sell16_split_kernel<2, I8><<<(unsigned)grid, kBlock, 0, ctx->stream>>>(HG_SELL16_ARGS); \
        else if (S == 4) spmv_sell16_split_kernel<4, I8><<<(unsigned)grid, kBlock, 0, ctx->stream>>>(HG_SELL16_ARGS); \
        else if (S == 8) spmv_sell16_split_kernel<8, I8><<<(unsigned)grid, kBlock, 0, ctx->stream>>>(HG_SELL16_ARGS); \
        else spmv_sell16_split_kernel<16, I8><<<(unsigned)grid, 512, 0, ctx->stream>>>(HG_SELL16_ARGS);           \
    } while (0)
    if (m->sell_col8) {
        const void* scol = m->sell_col8;
        HG_SELL16_LAUNCH(true);
    } else {
        const void* scol = m->sell_col16;
        HG_SELL16_LAUNCH(false);
    }
#undef HG_SELL16_LAUNCH
#undef HG_SELL16_ARGS
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_spmv_csr16(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y,
                    const hg_spmv_epilogue& ep, double bytes, int* nparts) {
    const int64_t grid = cdiv(m->rows, kBlock / 32);
    HG_REQUIRE(grid < (int64_t)2147483647, "spmv: too many rows for one launch");
    if (nparts && ep.stat) *nparts = (int)grid;
    hg_launch_scope scope(ctx, HG_K_SPMV, bytes);
    spmv_csr16_kernel<<<(unsigned)grid, kBlock, 0, ctx->stream>>>(
        m->rows, m->rowptr, m->csr_gptr, m->csr_col16, m->csr_base, m->vals, x, y, ep.alpha, ep.z1, ep.g1, ep.z2,
        ep.g2, ep.ref, ep.stat);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}
