// CGS2 middle stage with the basis tile staged in shared memory:
//     w1 = w0 - V h1      and      partials = V^T w1
// in ONE pass over V (the separate update + multi-dot kernels read V twice: 32kn -> 24kn bytes per
// Arnoldi step).  An earlier fused kernel (removed; numbers in profiles/r01_cgs_fusion.md) re-read
// the tile from L2 and stalled DRAM while it computed: 49 % DRAM utilisation, slower than the two
// streaming kernels.  Here a persistent CTA per SM owns two ~100 KB stages: while the warps work on
// the tile in one stage (phase A: V_tile h -> w1 tile; phase B: V_tile^T w1 tile, both from shared
// memory), asynchronous copies fill the other, so 100-200 KB per SM are always in flight whatever
// the warps are doing.  The copies are 16-byte cp.async (LDGSTS, L1 bypassed), 512 contiguous bytes
// per basis column per warp: a first version issued one cp.async.bulk (TMA) per column piece and was
// bound by the TMA unit's per-request cost at 8*TR-byte pieces (1.6 TB/s).
//
// Work split: warp w owns the columns j = w, w+8, ...; lane l owns rows l, l+32, ... of the tile, so
// every shared-memory read is conflict free and the phase-B sums stay in registers across all tiles
// of the CTA (one warp reduction per column at the very end).  Phase B takes its operands from the
// registers phase A loaded them into (KEEP), so the tile is read from shared memory once.
// Reductions are in a fixed order:
// reruns are bit-identical.
#include <algorithm>

#include <atomic>

#include "common.cuh"

namespace {

constexpr int kWarps = 16;  // 512 threads: enough warps per scheduler to hide the LDS / DFMA latencies
constexpr int kThreads = kWarps * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Shared-memory plan (bytes), stage = (k columns + w0) x TR rows, TR = 64 * NP.
struct Plan {
    int TR;
    size_t stage;  // bytes of one stage
    size_t total;  // dynamic shared memory of the kernel
};
inline Plan make_plan(int k, int NP) {
    Plan p;
    p.TR = 64 * NP;
    p.stage = (size_t)(k + 1) * p.TR * 8;
    // 2 stages | red[kWarps][TR] | w1 tile [TR] | h[k]
    p.total = 2 * p.stage + (size_t)(kWarps + 1) * p.TR * 8 + (size_t)((k + 1) / 2 * 2) * 8;
    return p;
}

// CPW: columns per warp (k <= 16*CPW); NP: row pairs per lane (tile rows TR = 64*NP; lane l owns rows
// 64p + 2l, 64p + 2l + 1: one 16-byte shared-memory load per column and pair)
template <int CPW, int NP, bool KEEP>
__global__ void __launch_bounds__(kThreads, 1)
cgs_mid_staged_kernel(const double* __restrict__ V, int64_t ld, int64_t n, int k,
                      const double* __restrict__ h, const double* __restrict__ w0,
                      double* __restrict__ w1, double* __restrict__ partials, int ntiles) {
    constexpr int TR = 64 * NP;
    extern __shared__ __align__(128) unsigned char smem[];
    const size_t stage_bytes = (size_t)(k + 1) * TR * 8;
    double* st0 = reinterpret_cast<double*>(smem);
    double* red = reinterpret_cast<double*>(smem + 2 * stage_bytes);   // [kWarps][TR]
    double* w1t = red + kWarps * TR;                                   // [TR]
    double* sh = w1t + TR;                                             // [k]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int j = threadIdx.x; j < k; j += kThreads) sh[j] = h[j];

    const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    // every thread copies its share of tile i into stage i & 1: (k basis columns + w0) x min(TR, ld - r0)
    // rows (ld is a multiple of 32 >= n: the piece never leaves its column), 16 bytes per cp.async; the
    // threads of a warp cover 512 contiguous bytes of one column.  Thread t owns the 16-byte chunk
    // t % CPC of the columns t / CPC, + CSTEP, ...  Always commits a group (possibly empty) so that "all
    // but the newest group" is tile i at iteration i.
    constexpr int CPC = TR / 2;              // chunks per column
    constexpr int CSTEP = kThreads / CPC;    // columns covered per round of the CTA
    const int my_off = (threadIdx.x % CPC) * 2;
    const int my_col0 = threadIdx.x / CPC;
    auto issue = [&](int i) {
        if (i < my_tiles) {
            const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * TR;
            if (my_off < (int)min((int64_t)TR, ld - r0)) {
                uint32_t dst = smem_u32(st0) + (uint32_t)((i & 1) * stage_bytes) + (uint32_t)(my_col0 * TR + my_off) * 8u;
                const double* src = V + (int64_t)my_col0 * ld + r0 + my_off;
                int col = my_col0;
                for (; col < k; col += CSTEP) {
                    cp_async16(dst, src);
                    dst += (uint32_t)CSTEP * TR * 8u;
                    src += (int64_t)CSTEP * ld;
                }
                if (col == k) cp_async16(dst, w0 + r0 + my_off);  // the w0 tile is "column k" of the stage
            }
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);

    double accB[CPW];
#pragma unroll
    for (int c = 0; c < CPW; ++c) accB[c] = 0.0;

    for (int i = 0; i < my_tiles; ++i) {
        const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * TR;
        cp_async_wait_1();
        __syncthreads();  // every thread's copies of tile i have landed
        const double* sv = st0 + (size_t)(i & 1) * (stage_bytes / 8);
        bool live[NP];  // rows past n hold whatever the padding of V holds: never used (n is even or the
                        // last row pair is split: handled through live + the row test below)
#pragma unroll
        for (int p = 0; p < NP; ++p) live[p] = r0 + 64 * p + 2 * lane + 1 < n;
        // ---- phase A: this warp's share of V_tile * h
        double2 pa[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) pa[p] = make_double2(0.0, 0.0);
        double2 keep[KEEP ? CPW : 1][KEEP ? NP : 1];  // KEEP: phase B re-uses the operands from registers
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const int j = warp + kWarps * c;
            if (j < k) {
                const double hj = sh[j];
                const double2* col = reinterpret_cast<const double2*>(sv + (size_t)j * TR) + lane;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const double2 v = col[32 * p];
                    if (KEEP) keep[KEEP ? c : 0][KEEP ? p : 0] = v;
                    pa[p].x = fma(hj, v.x, pa[p].x);
                    pa[p].y = fma(hj, v.y, pa[p].y);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) reinterpret_cast<double2*>(red + warp * TR)[32 * p + lane] = pa[p];
        __syncthreads();
        if (threadIdx.x < TR) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) sum += red[w * TR + threadIdx.x];
            const int64_t row = r0 + threadIdx.x;
            double out = 0.0;
            if (row < n) {
                out = sv[(size_t)k * TR + threadIdx.x] - sum;
                w1[row] = out;
            }
            w1t[threadIdx.x] = out;  // 0 for rows past n: they drop out of phase B
        }
        __syncthreads();
        // ---- phase B: V_tile^T w1_tile for the same columns, accumulated over all tiles of this CTA
        double2 wv[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) wv[p] = reinterpret_cast<const double2*>(w1t)[32 * p + lane];
        if (r0 + TR <= n) {  // CTA-uniform: every row of the tile is real (all tiles but possibly the last one)
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                const int j = warp + kWarps * c;
                if (j < k) {
                    const double2* col = reinterpret_cast<const double2*>(sv + (size_t)j * TR) + lane;
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        const double2 v = KEEP ? keep[KEEP ? c : 0][KEEP ? p : 0] : col[32 * p];
                        accB[c] = fma(v.x, wv[p].x, accB[c]);
                        accB[c] = fma(v.y, wv[p].y, accB[c]);
                    }
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                const int j = warp + kWarps * c;
                if (j < k) {
                    const double2* col = reinterpret_cast<const double2*>(sv + (size_t)j * TR) + lane;
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        const double2 v = KEEP ? keep[KEEP ? c : 0][KEEP ? p : 0] : col[32 * p];
                        // a pair straddling n: its first row is real, its second is padding
                        const bool first = r0 + 64 * p + 2 * lane < n;
                        accB[c] = fma(first ? v.x : 0.0, wv[p].x, accB[c]);
                        accB[c] = fma(live[p] ? v.y : 0.0, wv[p].y, accB[c]);
                    }
                }
            }
        }
        __syncthreads();  // every warp is done with this stage (and with red / w1t)
        issue(i + 2);
    }
    // one partial per (column, CTA); CTAs without tiles contribute zeros
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
        const int j = warp + kWarps * c;
        if (j < k) {
            const double sum = warp_sum(accB[c]);
            if (lane == 0) partials[(int64_t)j * gridDim.x + blockIdx.x] = sum;
        }
    }
}

template <int CPW, int NP, bool KEEP>
int launch(hg_ctx* ctx, const double* V, int64_t ld, int64_t n, int k, const double* h, const double* w0,
           double* w1, double* partials, int grid) {
    const Plan p = make_plan(k, NP);
    const int ntiles = (int)((n + p.TR - 1) / p.TR);
    // the opt-in is per device (and per template instance): one bit per device ordinal
    static std::atomic<unsigned long long> attr_set{0};
    const unsigned long long dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_set.load(std::memory_order_relaxed) & dev_bit)) {
        HG_CUDA(cudaFuncSetAttribute(cgs_mid_staged_kernel<CPW, NP, KEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
        attr_set.fetch_or(dev_bit, std::memory_order_relaxed);
    }
    cgs_mid_staged_kernel<CPW, NP, KEEP><<<grid, kThreads, p.total, ctx->stream>>>(V, ld, n, k, h, w0, w1, partials, ntiles);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

}  // namespace

// Number of partials per column (= CTAs) the kernel writes, or 0 when (n, k) is outside its range
// and the caller must use the separate update + multi-dot kernels.
int hg_cgs_staged_nparts(const hg_ctx* ctx, int64_t n, int k) {
    if (k < 24 || k > 208 || n < 256 * (int64_t)ctx->sm_count) return 0;
    return ctx->sm_count;
}

int hg_k_cgs_mid_staged(hg_ctx* ctx, const double* V, int64_t ld, int64_t n, int k, const double* h,
                     const double* w0, double* w1, double* partials, int* nparts) {
    const int grid = hg_cgs_staged_nparts(ctx, n, k);
    HG_REQUIRE(grid > 0, "cgs_mid_staged: (n, k) out of range");
    if (nparts) *nparts = grid;
    // algorithmic bytes: V once, w0 in, w1 out
    hg_launch_scope scope(ctx, HG_K_LINCOMB, 8.0 * (double)n * (double)k + 16.0 * (double)n);
    if (k <= 40) return launch<3, 4, true>(ctx, V, ld, n, k, h, w0, w1, partials, grid);   // 256-row tiles
    if (k <= 88) return launch<6, 2, true>(ctx, V, ld, n, k, h, w0, w1, partials, grid);   // 128-row tiles
    return launch<13, 1, true>(ctx, V, ld, n, k, h, w0, w1, partials, grid);               // 64-row tiles
}
