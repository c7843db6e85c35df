// The whole CGS2 orthogonalisation of one Arnoldi step in ONE persistent cooperative kernel:
//
//     h1 = V'w0 ; w1 = w0 - V h1 ; h2 = V'w1 ; v = w1 - V h2 ; H(1:k,k) = h1 + h2 ;
//     H(k+1,k) = ||v|| ; V(:,k+1) = v / H(k+1,k)          (hybrid_ab_gmres_rtp.m:20-26 in its two-pass form)
//
// The separate-kernel path launches multi-dot, reduce, staged middle stage, reduce, update, reduce, scale:
// seven kernels whose boundaries (drain + ramp, ~5 us each) and second-stage reductions cost as much as the
// data movement once a Krylov vector is a few hundred thousand rows — BASELINE configs[1]/[2] (256^2,
// 512^2) and every rank's slice of the 1024^2 problem at 4-8 GPUs.  Here one CTA per SM sweeps its row
// tiles three times (V crosses HBM — or, at these sizes, mostly L2 — three times, as before) with grid-wide
// barriers in between; every CTA reduces the per-CTA partial sums redundantly in a fixed order, so no
// second-stage kernel, no host involvement, and bit-identical reruns.
//
// Tile pipeline: as in cgs_staged.cu — two shared-memory stages filled by 16-byte cp.async while the 16
// warps work on the other stage; warp w owns columns j = w, w+16, ...; lane l owns row pairs of the tile.
#include <cooperative_groups.h>

#include <atomic>

#include "dist_internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct StepPlan {
    int TR;
    size_t stage, total;
};
inline StepPlan step_plan(int k, int NP) {
    StepPlan p;
    p.TR = 64 * NP;
    p.stage = (size_t)(k + 1) * p.TR * 8;
    // 2 stages | red[kWarps][TR] | w1 tile [TR] | h[kpad] | scratch[64]
    p.total = 2 * p.stage + (size_t)(kWarps + 1) * p.TR * 8 + (size_t)((k + 2) / 2 * 2) * 8 + 64 * 8;
    return p;
}

// Multi-GPU (NVLink peer memory, dist_peer.cu protocol): the same kernel also does the collectives of the
// sharded step — pulls the reduce-scatter of B^p u_p, exchanges the three coefficient / norm sums as LL
// words ({half, epoch} pairs, the data carries its flag), pushes the new basis rows into every rank's
// replicated vector and scales its local copy.  One launch replaces seven kernels + three all-reduce kernels.
struct StepPeer {
    hg_peer_tbl t;
    size_t ypart_off, yflag_off, qfull_off;  // byte offsets inside the symmetric workspace
    size_t inbox_off[3];
    unsigned long long y_epoch;
    unsigned int ar_epoch[3];
    unsigned long long* err;
    int64_t row0, n_pad;    // first global row of this rank's slice; length of the replicated vector
    const double* q_slice;  // local slice of q_k (shift term), or nullptr
    double shift;
};

struct StepArgs {
    const double* V;   // basis, column-major, ld
    int64_t ld, n;
    int k;             // columns to orthogonalise against
    const double* w0;  // operator result (n)
    double* w1;        // scratch (n): w0 - V h1
    double* qnext;     // V(:,k+1): receives v / ||v||
    double* Hcol;      // device H column: k + 1 entries
    double* hcur;      // k entries: last coefficient vector (kept for callers that read it)
    double* partials;  // (k + 2) * gridDim doubles
    int ntiles;
    double* w0w;       // PEER: w0 is produced here (same buffer as w0)
    StepPeer peer;     // PEER only
};

constexpr long long kStepSpinLimit = 40000000000LL;  // ~20 s of SM clocks: a lost rank is reported, not a hang

__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// All-reduce of sh[0..ncols) over the ranks.  CTA 0 stores this rank's sums into every rank's inbox (own
// included); EVERY CTA then polls its rank's inbox and adds the P contributions in rank order, so all CTAs
// of all ranks end with identical bits.  Inbox entry (j, src) is 16 bytes at ((j * P + src) * 2) words.
__device__ __forceinline__ void peer_allreduce(const StepPeer& pr, int which, double* sh, int ncols, double* stage,
                                               int warp, int lane) {
    const int P = pr.t.P;
    const unsigned int epoch = pr.ar_epoch[which];
    __syncthreads();  // sh holds the local sums
    if (blockIdx.x == 0) {
        __threadfence_system();  // rows this rank pushed to peers are ordered before the words that announce them
        for (int idx = threadIdx.x; idx < ncols * P; idx += kThreads) {
            const int j = idx / P, dstp = idx % P;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(sh[j]);
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(pr.t.base[dstp] + pr.inbox_off[which]) +
                                      ((size_t)j * P + pr.t.rank) * 2;
            st_relaxed_sys_u64(dst, (bits << 32) | epoch);
            st_relaxed_sys_u64(dst + 1, (bits & 0xffffffff00000000ull) | epoch);
        }
    }
    __syncthreads();  // CTA 0: everything read from sh before it is overwritten
    // every thread polls one (column, rank) entry at a time — ncols * P entries over 512 threads, i.e. one or
    // two NVLink latencies per exchange instead of one per column — then thread j adds its P values in rank order
    const unsigned long long* inbox = reinterpret_cast<const unsigned long long*>(pr.t.base[pr.t.rank] + pr.inbox_off[which]);
    double* vals = stage;  // ncols * P doubles of scratch (the tile stages are idle between sweeps)
    for (int idx = threadIdx.x; idx < ncols * P; idx += kThreads) {
        const unsigned long long* src = inbox + (size_t)idx * 2;  // entry (j, src rank) = j * P + rank
        const long long t0 = clock64();
        unsigned long long w0, w1;
        bool ok = true;
        for (;;) {
            w0 = ld_relaxed_sys_u64(src);
            w1 = ld_relaxed_sys_u64(src + 1);
            if ((unsigned int)w0 == epoch && (unsigned int)w1 == epoch) break;
            if (clock64() - t0 > kStepSpinLimit) {
                *pr.err = 1ull;
                ok = false;
                break;
            }
        }
        vals[idx] = ok ? __longlong_as_double((long long)((w0 >> 32) | (w1 & 0xffffffff00000000ull))) : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < ncols; j += kThreads) {
        double s = 0.0;
        for (int r = 0; r < P; ++r) s += vals[j * P + r];  // rank order
        sh[j] = s;
    }
    (void)warp;
    (void)lane;
    // acquire, once per exchange (a system-scope fence per polled word costs microseconds): reads of
    // peer-written rows that follow come after the observed epochs
    __threadfence_system();
    __syncthreads();
}

// sum over CTAs of partials[j * G + c] for the columns of this warp -> sh[j]; every CTA does the same work
// in the same order (lane l sums c = l, l+32, ..., then the shuffle tree), so all CTAs hold identical bits.
// The partials were written by other SMs during this kernel: read through L2 (__ldcg).
__device__ __forceinline__ void reduce_partials(const double* __restrict__ partials, int G, int ncols, double* sh,
                                                int warp, int lane) {
    for (int j = warp; j < ncols; j += kWarps) {
        const double* p = partials + (size_t)j * G;
        double s = 0.0;
        for (int c = lane; c < G; c += 32) s += __ldcg(p + c);
        s = warp_sum(s);
        if (lane == 0) sh[j] = s;
    }
}

template <int CPW, int NP, bool PEER>
__global__ void __launch_bounds__(kThreads, 1) cgs2_step_kernel(StepArgs a) {
    constexpr int TR = 64 * NP;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(128) unsigned char smem[];
    const int k = a.k;
    const int64_t n = a.n, ld = a.ld;
    const size_t stage_bytes = (size_t)(k + 1) * TR * 8;
    double* st0 = reinterpret_cast<double*>(smem);
    double* red = reinterpret_cast<double*>(smem + 2 * stage_bytes);  // [kWarps][TR]
    double* w1t = red + kWarps * TR;                                  // [TR]
    double* sh = w1t + TR;                                            // [k + 1 (+pad)] coefficients
    double* scratch = sh + (k + 2) / 2 * 2;                           // [64]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = (int)gridDim.x;
    const int my_tiles = a.ntiles > (int)blockIdx.x ? (a.ntiles - 1 - (int)blockIdx.x) / G + 1 : 0;

    constexpr int CPC = TR / 2;            // 16-byte chunks per column of a tile
    constexpr int CSTEP = kThreads / CPC;  // columns covered per round of the CTA
    const int my_off = (threadIdx.x % CPC) * 2;
    const int my_col0 = threadIdx.x / CPC;
    // stage tile i of this CTA: k basis columns + one vector (`vec`, "column k" of the stage)
    auto issue = [&](int i, const double* vec) {
        if (i < my_tiles) {
            const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)i * G) * TR;
            if (my_off < (int)min((int64_t)TR, ld - r0)) {
                uint32_t dst = smem_u32(st0) + (uint32_t)((i & 1) * stage_bytes) + (uint32_t)(my_col0 * TR + my_off) * 8u;
                const double* src = a.V + (int64_t)my_col0 * ld + r0 + my_off;
                int col = my_col0;
                for (; col < k; col += CSTEP) {
                    cp_async16(dst, src);
                    dst += (uint32_t)CSTEP * TR * 8u;
                    src += (int64_t)CSTEP * ld;
                }
                if (col == k) cp_async16(dst, vec + r0 + my_off);
            }
        }
        cp_async_commit();
    };

    double acc[CPW];

    if constexpr (PEER) {
        // ---------------------------------------------------------------- sweep 0: reduce-scatter by pulling
        // w0[r] = sum_p (B^p u_p)[row0 + r] + shift * q_k[r] on this CTA's tiles, once every rank's SpMV is done
        const StepPeer& pr = a.peer;
        if (threadIdx.x < pr.t.P) {
            const unsigned long long* f = reinterpret_cast<const unsigned long long*>(pr.t.base[pr.t.rank] + pr.yflag_off) + threadIdx.x;
            const long long t0 = clock64();
            while (ld_acquire_sys_u64(f) < pr.y_epoch) {
                if (clock64() - t0 > kStepSpinLimit) {
                    *pr.err = 1ull;
                    __threadfence_system();
                    break;
                }
                __nanosleep(32);
            }
        }
        __syncthreads();
        // all threads over all row pairs of this CTA's tiles at once: the P remote loads of a pair are issued
        // together, so the sweep costs about one NVLink round trip (tile by tile it was my_tiles of them)
        const int pairs_per_tile = TR / 2;
        for (int idx = threadIdx.x; idx < my_tiles * pairs_per_tile; idx += kThreads) {
            const int i = idx / pairs_per_tile, t2 = idx % pairs_per_tile;
            const int64_t r = ((int64_t)blockIdx.x + (int64_t)i * G) * TR + 2 * t2;  // n (= n_p) is a multiple of 32
            if (r < n) {
                double2 vv[HG_MAX_PEERS];
#pragma unroll
                for (int pp = 0; pp < HG_MAX_PEERS; ++pp)
                    if (pp < pr.t.P)
                        vv[pp] = __ldcg(reinterpret_cast<const double2*>(
                            reinterpret_cast<const double*>(pr.t.base[pp] + pr.ypart_off) + pr.row0 + r));
                double2 sacc = make_double2(0.0, 0.0);
#pragma unroll
                for (int pp = 0; pp < HG_MAX_PEERS; ++pp)  // fixed rank order: every rank forms the same bits
                    if (pp < pr.t.P) {
                        sacc.x += vv[pp].x;
                        sacc.y += vv[pp].y;
                    }
                if (pr.shift != 0.0) {
                    const double2 qv = *reinterpret_cast<const double2*>(pr.q_slice + r);
                    sacc.x += pr.shift * qv.x;
                    sacc.y += pr.shift * qv.y;
                }
                *reinterpret_cast<double2*>(a.w0w + r) = sacc;
            }
        }
        __threadfence();  // staged below through cp.async (L2)
        __syncthreads();
    }

    // ------------------------------------------------------------------ sweep 1: h1 partials = V' w0
#pragma unroll
    for (int c = 0; c < CPW; ++c) acc[c] = 0.0;
    issue(0, a.w0);
    issue(1, a.w0);
    for (int i = 0; i < my_tiles; ++i) {
        const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)i * G) * TR;
        cp_async_wait_1();
        __syncthreads();
        const double* sv = st0 + (size_t)(i & 1) * (stage_bytes / 8);
        double2 wv[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            wv[p] = reinterpret_cast<const double2*>(sv + (size_t)k * TR)[32 * p + lane];
            // rows past n hold whatever the padding holds: take them out on both operands
            if (!(r0 + 64 * p + 2 * lane < n)) wv[p].x = 0.0;
            if (!(r0 + 64 * p + 2 * lane + 1 < n)) wv[p].y = 0.0;
        }
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const int j = warp + kWarps * c;
            if (j < k) {
                const double2* col = reinterpret_cast<const double2*>(sv + (size_t)j * TR) + lane;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const double2 v = col[32 * p];
                    acc[c] = fma(r0 + 64 * p + 2 * lane < n ? v.x : 0.0, wv[p].x, acc[c]);
                    acc[c] = fma(r0 + 64 * p + 2 * lane + 1 < n ? v.y : 0.0, wv[p].y, acc[c]);
                }
            }
        }
        __syncthreads();
        issue(i + 2, a.w0);
    }
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
        const int j = warp + kWarps * c;
        if (j < k) {
            const double s = warp_sum(acc[c]);
            if (lane == 0) a.partials[(size_t)j * G + blockIdx.x] = s;
        }
    }
    cp_async_wait_0();
    grid.sync();
    reduce_partials(a.partials, G, k, sh, warp, lane);
    if constexpr (PEER) peer_allreduce(a.peer, 0, sh, k, st0, warp, lane);
    __syncthreads();
    if (blockIdx.x == 0)
        for (int j = threadIdx.x; j < k; j += kThreads) a.Hcol[j] = sh[j];  // h1

    // ------------------------------------------------------------------ sweep 2: w1 = w0 - V h1, h2 partials = V' w1
#pragma unroll
    for (int c = 0; c < CPW; ++c) acc[c] = 0.0;
    issue(0, a.w0);
    issue(1, a.w0);
    for (int i = 0; i < my_tiles; ++i) {
        const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)i * G) * TR;
        cp_async_wait_1();
        __syncthreads();
        const double* sv = st0 + (size_t)(i & 1) * (stage_bytes / 8);
        double2 pa[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) pa[p] = make_double2(0.0, 0.0);
        double2 keep[CPW][NP];
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const int j = warp + kWarps * c;
            if (j < k) {
                const double hj = sh[j];
                const double2* col = reinterpret_cast<const double2*>(sv + (size_t)j * TR) + lane;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const double2 v = col[32 * p];
                    keep[c][p] = v;
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
                a.w1[row] = out;
            }
            w1t[threadIdx.x] = out;  // 0 for rows past n: they drop out of the dots
        }
        __syncthreads();
        double2 wv[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) wv[p] = reinterpret_cast<const double2*>(w1t)[32 * p + lane];
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const int j = warp + kWarps * c;
            if (j < k) {
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const double2 v = keep[c][p];
                    acc[c] = fma(r0 + 64 * p + 2 * lane < n ? v.x : 0.0, wv[p].x, acc[c]);
                    acc[c] = fma(r0 + 64 * p + 2 * lane + 1 < n ? v.y : 0.0, wv[p].y, acc[c]);
                }
            }
        }
        __syncthreads();
        issue(i + 2, a.w0);
    }
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
        const int j = warp + kWarps * c;
        if (j < k) {
            const double s = warp_sum(acc[c]);
            if (lane == 0) a.partials[(size_t)j * G + blockIdx.x] = s;
        }
    }
    cp_async_wait_0();
    __threadfence();  // w1 rows of this CTA are re-read below through cp.async (L2)
    grid.sync();
    reduce_partials(a.partials, G, k, sh, warp, lane);
    if constexpr (PEER) peer_allreduce(a.peer, 1, sh, k, st0, warp, lane);
    __syncthreads();
    if (blockIdx.x == 0)
        for (int j = threadIdx.x; j < k; j += kThreads) {
            a.Hcol[j] = a.Hcol[j] + sh[j];  // H(1:k,k) = h1 + h2
            a.hcur[j] = sh[j];
        }

    // ------------------------------------------------------------------ sweep 3: v = w1 - V h2, ||v||^2 partial
    double nrm_acc = 0.0;
    issue(0, a.w1);
    issue(1, a.w1);
    for (int i = 0; i < my_tiles; ++i) {
        const int64_t r0 = ((int64_t)blockIdx.x + (int64_t)i * G) * TR;
        cp_async_wait_1();
        __syncthreads();
        const double* sv = st0 + (size_t)(i & 1) * (stage_bytes / 8);
        double2 pa[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) pa[p] = make_double2(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
            const int j = warp + kWarps * c;
            if (j < k) {
                const double hj = sh[j];
                const double2* col = reinterpret_cast<const double2*>(sv + (size_t)j * TR) + lane;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const double2 v = col[32 * p];
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
            if (row < n) {
                const double out = sv[(size_t)k * TR + threadIdx.x] - sum;
                a.qnext[row] = out;
                if constexpr (PEER) {  // all-gather by the producer: the row goes into every rank's replicated vector
#pragma unroll 4
                    for (int pp = 0; pp < a.peer.t.P; ++pp)
                        reinterpret_cast<double*>(a.peer.t.base[pp] + a.peer.qfull_off)[a.peer.row0 + row] = out;
                }
                nrm_acc = fma(out, out, nrm_acc);
            }
        }
        __syncthreads();
        issue(i + 2, a.w1);
    }
    cp_async_wait_0();
    {   // CTA sum of nrm_acc in a fixed order: warp shuffle tree, then warps in order
        const double s = warp_sum(nrm_acc);
        if (lane == 0) scratch[warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < kWarps; ++w) t += scratch[w];
            a.partials[(size_t)k * G + blockIdx.x] = t;
        }
    }
    if constexpr (PEER) __threadfence_system();  // pushed rows are performed before anything that follows the barrier
    __threadfence();
    grid.sync();
    if (warp == 0) {
        const double* p = a.partials + (size_t)k * G;
        double s = 0.0;
        for (int c = lane; c < G; c += 32) s += __ldcg(p + c);
        s = warp_sum(s);
        if (lane == 0) scratch[32] = s;
    }
    // PEER: this exchange is also the barrier that makes every rank's pushed rows visible here
    if constexpr (PEER) peer_allreduce(a.peer, 2, scratch + 32, 1, st0, warp, lane);
    __syncthreads();
    const double nrm = sqrt(scratch[32]);  // H(k+1,k) = norm(v)
    if (blockIdx.x == 0 && threadIdx.x == 0) a.Hcol[k] = nrm;
    // ------------------------------------------------------------------ sweep 4: V(:,k+1) = v / H(k+1,k)
    // all threads over all rows of this CTA's tiles at once (tile by tile it was my_tiles dependent round trips)
    for (int idx = threadIdx.x; idx < my_tiles * TR; idx += kThreads) {
        const int64_t row = ((int64_t)blockIdx.x + (int64_t)(idx / TR) * G) * TR + idx % TR;
        if (row < n) a.qnext[row] = __ldcg(a.qnext + row) / nrm;  // division, as the reference (:26)
    }
    if constexpr (PEER) {  // ... and this rank's copy of the replicated vector (rows from all ranks)
        double2* qf = reinterpret_cast<double2*>(a.peer.t.base[a.peer.t.rank] + a.peer.qfull_off);
        const int64_t n2 = a.peer.n_pad / 2;  // n_pad is a multiple of 32
        const int64_t stride = (int64_t)G * kThreads;
        int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
        for (; i + 3 * stride < n2; i += 4 * stride) {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldcg(qf + i + u * stride);
#pragma unroll
            for (int u = 0; u < 4; ++u) qf[i + u * stride] = make_double2(v[u].x / nrm, v[u].y / nrm);
        }
        for (; i < n2; i += stride) {
            const double2 v = __ldcg(qf + i);
            qf[i] = make_double2(v.x / nrm, v.y / nrm);
        }
    }
}

template <int CPW, int NP, bool PEER>
int launch_step(hg_ctx* ctx, StepArgs& a) {
    const StepPlan p = step_plan(a.k, NP);
    a.ntiles = (int)((a.n + p.TR - 1) / p.TR);
    static std::atomic<unsigned long long> attr_set{0};
    const unsigned long long dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_set.load(std::memory_order_relaxed) & dev_bit)) {
        HG_CUDA(cudaFuncSetAttribute(cgs2_step_kernel<CPW, NP, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set.fetch_or(dev_bit, std::memory_order_relaxed);
    }
    void* params[] = {&a};
    HG_CUDA(cudaLaunchCooperativeKernel((const void*)cgs2_step_kernel<CPW, NP, PEER>, dim3(ctx->sm_count), dim3(kThreads),
                                        params, p.total, ctx->stream));
    return HG_OK;
}

int g_step_max_n = -1, g_step_max_n_dist = -1;

}  // namespace

// Largest Krylov vector length for which the whole-step kernel is used (option "cgs_step_max_n" / env
// HG_CGS_STEP_MAX_N; 0 disables).  Measured on one B200, 200-step cycles (profiles/r02_cgs_step_compare.jsonl):
// 155 vs 166 us per step at n = 65 536, 256 vs 255 at 132 496, 457 vs 432 at 262 144 — above ~130 000 rows the
// separate streaming kernels win (their multi-dot and update sweeps run at 6.5 TB/s, the staged pipeline of
// this kernel at ~4.5), below it the seven launches and three second-stage reductions they need cost more.
int64_t hg_cgs2_step_max_n() {
    if (g_step_max_n < 0) {
        const char* e = getenv("HG_CGS_STEP_MAX_N");
        g_step_max_n = e ? atoi(e) : 140000;
    }
    return g_step_max_n;
}
void hg_cgs2_step_max_n_set(int v) { g_step_max_n = v < 0 ? 0 : v; }
// The same for a rank's slice on several GPUs (option "cgs_step_max_n_dist" / env HG_CGS_STEP_MAX_N_DIST), where
// the kernel also replaces the pull, three all-reduce kernels and the scale kernel of the peer transport: 4
// launches per step instead of 10.  OFF by default (0): measured on B200s (profiles/r02_multi_gpu.md) it does not
// beat the separate kernels — 8 GPUs x 131 072-row slices 2 575 vs 3 300 it/s before the pull / poll / scale
// loops were parallelised, 2 GPUs x 131 072 rows 3 445 vs 3 692 after, a tie at 32 768 rows — because each of
// its three in-kernel exchanges costs a grid barrier + a redundant reduction + an NVLink round trip in sequence,
// where the separate all-reduce kernel works on all coefficients at once.  Kept as a tested option.
int64_t hg_cgs2_step_max_n_dist() {
    if (g_step_max_n_dist < 0) {
        const char* e = getenv("HG_CGS_STEP_MAX_N_DIST");
        g_step_max_n_dist = e ? atoi(e) : 0;
    }
    return g_step_max_n_dist;
}
void hg_cgs2_step_max_n_dist_set(int v) { g_step_max_n_dist = v < 0 ? 0 : v; }

bool hg_cgs2_step_eligible(const hg_ctx* ctx, int64_t n, int k) {
    (void)ctx;
    return k >= 1 && k <= 208 && n >= 1 && n <= hg_cgs2_step_max_n();
}
bool hg_cgs2_step_eligible_dist(const hg_ctx* ctx, int64_t n_p, int k) {
    (void)ctx;
    return k >= 1 && k <= 208 && n_p >= 1 && n_p <= hg_cgs2_step_max_n_dist();
}

// partials capacity needed: (k + 2) * sm_count doubles
int hg_k_cgs2_step(hg_ctx* ctx, const double* V, int64_t ld, int64_t n, int k, const double* w0, double* w1,
                   double* qnext, double* Hcol, double* hcur, double* partials) {
    HG_REQUIRE(k >= 1 && k <= 208 && n >= 1, "cgs2_step: (n, k) out of range");
    StepArgs a;
    a.V = V;
    a.ld = ld;
    a.n = n;
    a.k = k;
    a.w0 = w0;
    a.w1 = w1;
    a.qnext = qnext;
    a.Hcol = Hcol;
    a.hcur = hcur;
    a.partials = partials;
    a.ntiles = 0;
    a.w0w = nullptr;
    // algorithmic bytes: three sweeps over V_k, w0 twice, w1 out + in, q out + scale
    hg_launch_scope scope(ctx, HG_K_LINCOMB, 24.0 * (double)n * (double)k + 56.0 * (double)n);
    if (k <= 16) return launch_step<1, 4, false>(ctx, a);   // 256-row tiles
    if (k <= 40) return launch_step<3, 4, false>(ctx, a);   // 256-row tiles (2 x 84 KB stages)
    if (k <= 96) return launch_step<6, 2, false>(ctx, a);   // 128-row tiles
    return launch_step<13, 1, false>(ctx, a);               // 64-row tiles
}

// Sharded step over NVLink peer memory: waits for the latest HG_FLAG_Y signal of every rank, then
// w0 = sum_p ypart_p[slice] + shift * q_slice, the CGS2 step on the local slices with the three sums
// all-reduced inside the kernel, the un-normalised rows pushed into every rank's replicated vector `buf`,
// and finally qnext and the local replicated copy divided by H(k+1,k).  Uses three inbox slots.
int hg_k_cgs2_step_peer(hg_comm* c, const double* V, int64_t ld, int64_t n_p, int k, double* w0, double* w1,
                        double* qnext, double* Hcol, double* hcur, double* partials, int64_t row0,
                        const double* q_slice, double shift, int buf) {
    hg_ctx* ctx = c->ctx;
    HG_REQUIRE(k >= 1 && k <= 208 && n_p >= 1 && k + 1 <= c->lay.kpad, "cgs2_step_peer: (n, k) out of range");
    StepArgs a;
    a.V = V;
    a.ld = ld;
    a.n = n_p;
    a.k = k;
    a.w0 = w0;
    a.w0w = w0;
    a.w1 = w1;
    a.qnext = qnext;
    a.Hcol = Hcol;
    a.hcur = hcur;
    a.partials = partials;
    a.ntiles = 0;
    StepPeer& pr = a.peer;
    pr.t = c->tbl;
    pr.ypart_off = c->lay.ypart;
    pr.yflag_off = c->lay.flags + (size_t)HG_FLAG_Y * HG_MAX_PEERS * 8;
    pr.qfull_off = c->lay.qfull + (size_t)buf * c->lay.n_pad * 8;
    pr.y_epoch = c->bar_seq[HG_FLAG_Y];
    for (int i = 0; i < 3; ++i) {
        const unsigned long long seq = c->bar_seq[HG_FLAG_AR]++;
        pr.inbox_off[i] = c->lay.inbox + (size_t)(seq % kInboxSlots) * c->nranks * c->lay.kpad * 16;
        pr.ar_epoch[i] = (unsigned int)(seq % 0xfffffffeull) + 1u;
    }
    pr.err = c->d_err;
    pr.row0 = row0;
    pr.n_pad = c->lay.n_pad;
    pr.q_slice = q_slice;
    pr.shift = q_slice ? shift : 0.0;
    hg_launch_scope scope(ctx, HG_K_LINCOMB, 24.0 * (double)n_p * (double)k + 8.0 * (double)n_p * (c->nranks + 8) +
                                                 16.0 * (double)c->lay.n_pad);
    if (k <= 16) return launch_step<1, 4, true>(ctx, a);
    if (k <= 40) return launch_step<3, 4, true>(ctx, a);
    if (k <= 96) return launch_step<6, 2, true>(ctx, a);
    return launch_step<13, 1, true>(ctx, a);
}
