// Streaming CSR SpMV for long-row matrices (CT projectors / back-projectors).
//
// ncu on the row-per-warp kernel (profiles/r01_spmv_v1_rowwarp.txt) showed DRAM
// traffic equal to the algorithmic bytes but only 45-58 % DRAM utilisation with
// long-scoreboard stalls: every trip exposes DRAM latency (col/val loads) and
// then L2 latency (the dependent x gather), and row tails run without ILP.
//
// This kernel decouples the two.  The nnz stream (vals + column indices, the
// 12 B/nnz that dominate) is moved HBM -> shared memory by the TMA unit with
// cp.async.bulk (SASS UBLKCP) into a per-warp ring of STAGES chunks guarded by
// mbarriers, so STAGES-1 chunks per warp (>=140 KB per SM) are always in flight
// regardless of what the warp is doing.  The warp itself only reads shared
// memory, issues CW/32 independent x gathers per lane (L1/L2 hits), and walks
// the row boundaries.  Each warp owns one contiguous, nnz-balanced range of
// whole rows (the "unit" table, built once per matrix), so no row is split
// between warps: no atomics, no fix-up pass, bit-reproducible results.
// Row results are staged in lanes and written 32 rows at a time so the
// epilogue vectors (lambda*q, -alpha*u, b) and y are accessed coalesced.
#include <cstdlib>

#include <atomic>

#include "common.cuh"

namespace {

constexpr int kWarps = 8;  // per CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// HBM -> shared bulk copy through the TMA unit; streamed data is marked evict-first
// in L2 so the gathered x vector and the Krylov basis keep their lines.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int CW, int STAGES>
struct StreamCfg {
    static constexpr int T = CW / 32;                                // elements per lane per chunk
    static constexpr int kWarpBytes = STAGES * CW * 12 + 64;         // vals + cols + barriers
    static constexpr int kSmem = kWarps * kWarpBytes;
};

template <int CW, int STAGES>
__global__ void __launch_bounds__(kWarps * 32, 1)
spmv_stream_kernel(int64_t rows, const int64_t* __restrict__ rowptr, const int* __restrict__ colind,
                   const double* __restrict__ vals, const int64_t* __restrict__ unit_row,
                   const double* __restrict__ x, double* __restrict__ y, double alpha,
                   const double* __restrict__ z1, double g1, const double* __restrict__ z2, double g2,
                   const double* __restrict__ ref, double* __restrict__ stat) {
    using Cfg = StreamCfg<CW, STAGES>;
    constexpr int T = Cfg::T;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * kWarps + warp;
    unsigned char* base = smem + (size_t)warp * Cfg::kWarpBytes;
    double* sval = reinterpret_cast<double*>(base);
    int* scol = reinterpret_cast<int*>(base + (size_t)STAGES * CW * 8);
    const uint32_t bar0 = smem_u32(base + (size_t)STAGES * CW * 12);
    const uint32_t sval_u = smem_u32(sval), scol_u = smem_u32(scol);

    const int64_t r0 = unit_row[unit], r1 = unit_row[unit + 1];
    double stat_acc = 0.0;
    if (r0 < r1) {
        const int64_t u_lo = rowptr[r0], u_hi = rowptr[r1];
        const int64_t a_lo = u_lo & ~(int64_t)3;  // 16-byte aligned start of the bulk copies
        const int nchunks = (int)((u_hi - a_lo + CW - 1) / CW);
        const uint64_t policy = policy_evict_first();
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        auto issue = [&](int c) {
            const int s = c % STAGES;
            const int64_t g0 = a_lo + (int64_t)c * CW;
            int64_t cnt = u_hi - g0;
            cnt = (cnt + 3) & ~(int64_t)3;  // arrays are padded by 16 entries
            if (cnt > CW) cnt = CW;
            const uint32_t bar = bar0 + 8 * s;
            mbar_expect_tx(bar, (uint32_t)cnt * 12u);
            bulk_g2s(sval_u + (uint32_t)s * CW * 8u, vals + g0, (uint32_t)cnt * 8u, bar, policy);
            bulk_g2s(scol_u + (uint32_t)s * CW * 4u, colind + g0, (uint32_t)cnt * 4u, bar, policy);
        };
        if (lane == 0)
            for (int c = 0; c < nchunks && c < STAGES - 1; ++c) issue(c);

        // row pointers: lane l holds rowptr[pb + l + 1] for the current and the next batch of 32 rows
        int64_t pb = r0;
        auto load_ptrs = [&](int64_t b) {
            const int64_t i = b + lane + 1;
            return rowptr[i <= rows ? i : rows];
        };
        int64_t ptr_cur = load_ptrs(pb), ptr_next = load_ptrs(pb + 32);
        int64_t r = r0;                               // current row (warp uniform)
        int64_t re = __shfl_sync(0xffffffffu, ptr_cur, 0);  // its end pointer
        double acc = 0.0;                             // this lane's share of the current row
        double res = 0.0;                             // lane j holds the result of staged row j
        int nres = 0;

        auto flush = [&]() {  // write the nres staged rows [r - nres, r) coalesced
            if (lane < nres) {
                const int64_t rr = r - nres + lane;
                double out = alpha * res;
                if (z1) out += g1 * z1[rr];
                if (z2) out += g2 * z2[rr];
                if (y) y[rr] = out;
                if (stat) {
                    const double d = ref ? out - ref[rr] : out;
                    stat_acc = fma(d, d, stat_acc);
                }
            }
            nres = 0;
        };
        auto finish_row = [&]() {  // current row is complete: reduce, stage, advance
            const double tot = warp_sum(acc);
            acc = 0.0;
            if (lane == nres) res = tot;
            ++nres;
            ++r;
            if (nres == 32 || r == r1) flush();
            if (r - pb == 32) {
                pb += 32;
                ptr_cur = ptr_next;
                ptr_next = load_ptrs(pb + 32);
            }
            re = __shfl_sync(0xffffffffu, ptr_cur, (int)(r - pb));
        };

        for (int c = 0; c < nchunks; ++c) {
            // refill the stage consumed in the previous trip (all lanes are past it: __syncwarp below)
            if (lane == 0 && c + STAGES - 1 < nchunks) issue(c + STAGES - 1);
            const int s = c % STAGES;
            mbar_wait(bar0 + 8 * s, (uint32_t)((c / STAGES) & 1));
            const int64_t g0 = a_lo + (int64_t)c * CW;
            const int64_t chunk_lo = g0 > u_lo ? g0 : u_lo;
            const int64_t chunk_hi = (g0 + CW) < u_hi ? (g0 + CW) : u_hi;
            const double* sv = sval + (size_t)s * CW;
            const int* sc = scol + (size_t)s * CW;
            double p[T];
            {
                double xv[T];
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const int64_t idx = g0 + lane + 32 * t;
                    const bool ok = idx >= chunk_lo && idx < chunk_hi;
                    xv[t] = ok ? __ldg(x + sc[lane + 32 * t]) : 0.0;
                }
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    const int64_t idx = g0 + lane + 32 * t;
                    const bool ok = idx >= chunk_lo && idx < chunk_hi;
                    p[t] = ok ? sv[lane + 32 * t] * xv[t] : 0.0;
                }
            }
            int64_t done = chunk_lo;
            while (r < r1) {
                const int64_t lim = re < chunk_hi ? re : chunk_hi;
                if (done == chunk_lo && lim == chunk_hi) {
                    // whole chunk belongs to the current row (out-of-range slots hold 0)
#pragma unroll
                    for (int t = 0; t < T; ++t) acc += p[t];
                } else if (lim > done) {
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const int64_t idx = g0 + lane + 32 * t;
                        if (idx >= done && idx < lim) acc += p[t];
                    }
                }
                done = lim;
                if (re <= chunk_hi) finish_row();  // includes empty rows (re == done)
                else break;                        // row continues in the next chunk
                if (done == chunk_hi && re > chunk_hi) break;
            }
            __syncwarp();
        }
        // rows after the last nonzero of the unit are empty
        while (r < r1) finish_row();
    }
    if (stat) {
        const double t = warp_sum(stat_acc);
        if (lane == 0) stat[unit] = t;
    }
}

__global__ void build_units_kernel(int64_t rows, const int64_t* __restrict__ rowptr, int n_units,
                                   int64_t* __restrict__ unit_row) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u > n_units) return;
    if (u == n_units) {
        unit_row[u] = rows;
        return;
    }
    const int64_t nnz = rowptr[rows];
    // first row whose start is >= u/n_units of the nonzeros
    const int64_t target = (int64_t)(((__int128)nnz * u) / n_units);
    int64_t lo = 0, hi = rows;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (rowptr[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    unit_row[u] = u == 0 ? 0 : lo;
}

}  // namespace

// 0: auto, 1: force row-per-warp kernel, 2: force streaming kernel (HG_SPMV=v1|v2)
static int g_spmv_mode = -1;
static int spmv_mode() {
    if (g_spmv_mode < 0) {
        const char* e = getenv("HG_SPMV");
        g_spmv_mode = 0;
        if (e && (e[0] == 'v' || e[0] == 'V')) g_spmv_mode = e[1] == '1' ? 1 : (e[1] == '2' ? 2 : (e[1] == '3' ? 3 : 0));
    }
    return g_spmv_mode;
}

int hg_spmv_mode() { return spmv_mode(); }

static int g_cgs_fused = -1;
int hg_cgs_fused_mode() {
    if (g_cgs_fused < 0) {
        const char* e = getenv("HG_CGS_FUSED");
        g_cgs_fused = (e && e[0] == '0') ? 0 : 2;  // default: shared-memory-staged one-pass kernel
    }
    return g_cgs_fused;
}

void hg_dist_transport_set(int v);
void hg_idx16_set(int v);

extern "C" int hg_set_option(const char* name, int value) {
    HG_REQUIRE(name, "hg_set_option: NULL name");
    if (strcmp(name, "cgs_fused") == 0) {
        HG_REQUIRE(value == 0 || value == 2, "hg_set_option: cgs_fused must be 0 or 2");
        g_cgs_fused = value;
        return HG_OK;
    }
    if (strcmp(name, "spmv_mode") == 0) {
        HG_REQUIRE(value >= 0 && value <= 3, "hg_set_option: spmv_mode must be 0, 1, 2 or 3");
        g_spmv_mode = value;
        return HG_OK;
    }
    if (strcmp(name, "spmv_idx16") == 0) {
        hg_idx16_set(value);
        return HG_OK;
    }
    if (strcmp(name, "cgs_step_max_n") == 0) {
        hg_cgs2_step_max_n_set(value);
        return HG_OK;
    }
    if (strcmp(name, "cgs_step_max_n_dist") == 0) {
        hg_cgs2_step_max_n_dist_set(value);
        return HG_OK;
    }
    if (strcmp(name, "spmv_idx8") == 0) {
        hg_idx8_set(value);
        return HG_OK;
    }
    if (strcmp(name, "gkb_residual") == 0) {
        hg_gkb_residual_mode_set(value);
        return HG_OK;
    }
    if (strcmp(name, "spmv_group_split") == 0) {
        hg_spmv_group_split_set(value);
        return HG_OK;
    }
    if (strcmp(name, "spmv_group_min_rows") == 0) {
        hg_spmv_group_min_rows_set(value);
        return HG_OK;
    }
    if (strcmp(name, "spmv_group16") == 0) {
        hg_spmv_group16_set(value);
        return HG_OK;
    }
    if (strcmp(name, "spmv_group") == 0) {
        hg_spmv_group_set(value);
        return HG_OK;
    }
    if (strcmp(name, "dist_transport") == 0) {
        HG_REQUIRE(value >= 0 && value <= 2, "hg_set_option: dist_transport must be 0 (auto), 1 (nccl) or 2 (peer)");
        hg_dist_transport_set(value);
        return HG_OK;
    }
    hg_set_error("hg_set_option: unknown option '%s'", name);
    return HG_ERR_INVALID;
}

bool hg_spmv_stream_eligible(const hg_matrix* m) {
    const int mode = spmv_mode();
    if (mode == 1) return false;
    if (m->rows < 1 || m->nnz < 1) return false;
    // Measured on the 1024^2 fan-beam problem (profiles/r01_spmv_variants.md): the streaming
    // kernel moves the nnz stream at full rate but its per-warp private row ranges destroy the
    // L1 sharing of gathered x sectors between adjacent rows (L1 hit 3 % vs 33-50 %) and its
    // 196 KB of staging leaves 32 KB of L1, so it loses to the software-pipelined row-per-warp
    // kernel (2.0 vs 5.2 TB/s).  It stays available as spmv_mode=2 for matrices whose gathers
    // are cheap; auto mode does not select it.
    return mode == 2;
}

int hg_k_spmv_stream(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y,
                     const hg_spmv_epilogue& ep, int* nparts) {
    constexpr int CW = 512, STAGES = 4;
    using Cfg = StreamCfg<CW, STAGES>;
    hg_matrix* mm = const_cast<hg_matrix*>(m);  // the unit table is a lazily built cache
    const int n_units = ctx->sm_count * kWarps;
    if (!mm->unit_row || mm->n_units != n_units) {
        if (mm->unit_row) {
            HG_CUDA(cudaStreamSynchronize(ctx->stream));
            HG_CUDA(cudaFree(mm->unit_row));
            mm->unit_row = nullptr;
        }
        HG_CUDA(cudaMalloc(&mm->unit_row, (size_t)(n_units + 1) * sizeof(int64_t)));
        mm->n_units = n_units;
        hg_launch_scope scope(ctx, HG_K_SETUP, 8.0 * n_units * 20);
        build_units_kernel<<<(n_units + 1 + 127) / 128, 128, 0, ctx->stream>>>(m->rows, m->rowptr, n_units,
                                                                              mm->unit_row);
        HG_CUDA(cudaGetLastError());
    }
    // the opt-in is per device (and per template instance): one bit per device ordinal
    static std::atomic<unsigned long long> attr_set{0};
    const unsigned long long dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_set.load(std::memory_order_relaxed) & dev_bit)) {
        HG_CUDA(cudaFuncSetAttribute(spmv_stream_kernel<CW, STAGES>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
        attr_set.fetch_or(dev_bit, std::memory_order_relaxed);
    }
    if (nparts) *nparts = ep.stat ? n_units : 0;
    double bytes = 12.0 * (double)m->nnz + 8.0 * (double)(m->rows + 1) + 8.0 * (double)m->cols;
    if (y) bytes += 8.0 * (double)m->rows;
    if (ep.z1) bytes += 8.0 * (double)m->rows;
    if (ep.z2) bytes += 8.0 * (double)m->rows;
    if (ep.ref) bytes += 8.0 * (double)m->rows;
    hg_launch_scope scope(ctx, HG_K_SPMV, bytes);
    spmv_stream_kernel<CW, STAGES><<<ctx->sm_count, kWarps * 32, Cfg::kSmem, ctx->stream>>>(
        m->rows, m->rowptr, m->colind, m->vals, mm->unit_row, x, y, ep.alpha, ep.z1, ep.g1, ep.z2, ep.g2,
        ep.ref, ep.stat);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}
