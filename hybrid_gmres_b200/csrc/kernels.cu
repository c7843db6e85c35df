// Hand-written sm_100a FP64 kernels of the Arnoldi / Golub-Kahan hot path.
//
// Every op on this path is a memory-bound SpMV / GEMV (<= 0.25 flop/B, two
// orders below the FP64 ridge), so the kernels are CUDA-core FP64 with
// coalesced streaming loads that bypass L1 (the L1 is reserved for the gathered
// x-vector sectors), and deterministic two-stage reductions (no float atomics):
// reruns are bit-identical, which the GCV/fminbnd parity rule needs.
//
//   spmv_csr_kernel   K1/K2/K3/K8/K9/K11 of SURVEY.md §2b  (A*v, B*u + lambda*q,
//                     A'*u - beta*v, b - A*x, fused square-sums)
//   multidot_kernel   K4a  h = V^T w   (hybrid_ab_gmres_rtp.m:21 for all j at once)
//   lincomb_kernel    K4b/K5/K7/K8  w -= V h (+||w||^2), x = Q y (+||x-x_true||^2),
//                     r = b - W y (+||r||^2)
//   reduce_kernel     second stage of every reduction
//   vector kernels    K5 scale, K10 LSQR/LSMR recurrences
#include "common.cuh"

namespace {

constexpr int kBlock = 256;

// ---- streaming loads: read-only path, do not allocate in L1 ----------------
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
__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum (fixed tree); result valid in thread 0.
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double s_red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect s_red reuse
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? s_red[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

// ---------------------------------------------------------------------------
// CSR SpMV, TPR threads cooperate on one row
// ---------------------------------------------------------------------------
template <int TPR>
__global__ void __launch_bounds__(kBlock)
spmv_csr_kernel(int64_t rows, const int64_t* __restrict__ rowptr, const int* __restrict__ colind,
                const double* __restrict__ vals, const double* __restrict__ x,
                double* __restrict__ y, double alpha, const double* __restrict__ z1, double g1,
                const double* __restrict__ z2, double g2, const double* __restrict__ ref,
                double* __restrict__ stat) {
    constexpr int RPB = kBlock / TPR;
    const int lane = threadIdx.x % TPR;
    const int64_t row = (int64_t)blockIdx.x * RPB + threadIdx.x / TPR;
    double sq = 0.0;
    const bool valid = row < rows;
    // invalid rows run zero trips but still take part in the shuffles below
    const int64_t s = valid ? rowptr[row] : 0;
    const int64_t e = valid ? rowptr[row + 1] : 0;
    // Software-pipelined batches of U*TPR entries: every (col,val) load of a batch — the
    // ragged tail included, by predication — is issued together, and the next batch's
    // stream loads are in flight while this batch's dependent x gathers resolve, so a row
    // exposes one DRAM latency plus one L2 latency per batch instead of two serial
    // latencies per trip and one per tail element (ncu: long_scoreboard dominated).
    constexpr int U = 4;
    double a[U];
    int c[U];
    double v[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = 0.0;
    int64_t i = s + lane;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        ok[u] = i + u * TPR < e;
        c[u] = ok[u] ? ld_stream(colind + i + u * TPR) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ok[u] ? ld_stream(vals + i + u * TPR) : 0.0;
    while (i - lane < e) {  // warp-uniform per row group
        double xg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) xg[u] = ok[u] ? __ldg(x + c[u]) : 0.0;
        const int64_t in = i + U * TPR;
        int cn[U];
        double vn[U];
        bool okn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            okn[u] = in + u * TPR < e;
            cn[u] = okn[u] ? ld_stream(colind + in + u * TPR) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) vn[u] = okn[u] ? ld_stream(vals + in + u * TPR) : 0.0;
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
    for (int o = TPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o, TPR);
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
        const double t = block_sum(sq);
        if (threadIdx.x == 0) stat[blockIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------
// multidot: partials[j*nslabs + slab] = sum_{r in slab} V[r,j] * w[r]
// One CTA owns a row slab (w slab staged in shared memory once), warps own
// columns; each column slab is a contiguous 8*R-byte stream.
// ---------------------------------------------------------------------------
constexpr int kDotWarps = 8;

__global__ void __launch_bounds__(kDotWarps * 32)
multidot_kernel(const double* __restrict__ V, int64_t ld, int64_t n, int k,
                const double* __restrict__ w, double* __restrict__ partials, int nslabs, int R) {
    extern __shared__ double sw[];
    const int slab = blockIdx.x;
    const int64_t r0 = (int64_t)slab * R;
    const int len = (int)min((int64_t)R, n - r0);
    for (int i = threadIdx.x; i < R; i += blockDim.x) sw[i] = i < len ? w[r0 + i] : 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool full = (len == R);
    for (int j = warp; j < k; j += kDotWarps) {
        const double* col = V + (int64_t)j * ld + r0;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if (full) {
            // R is a multiple of 256: four independent 16-byte loads per lane per trip
            for (int i = 2 * lane; i < R; i += 256) {
                const double2 v0 = ld_stream2(col + i);
                const double2 v1 = ld_stream2(col + i + 64);
                const double2 v2 = ld_stream2(col + i + 128);
                const double2 v3 = ld_stream2(col + i + 192);
                a0 = fma(v0.x, sw[i], a0);
                a0 = fma(v0.y, sw[i + 1], a0);
                a1 = fma(v1.x, sw[i + 64], a1);
                a1 = fma(v1.y, sw[i + 65], a1);
                a2 = fma(v2.x, sw[i + 128], a2);
                a2 = fma(v2.y, sw[i + 129], a2);
                a3 = fma(v3.x, sw[i + 192], a3);
                a3 = fma(v3.y, sw[i + 193], a3);
            }
        } else {
            for (int i = lane; i < len; i += 32) a0 = fma(ld_stream(col + i), sw[i], a0);
        }
        double sum = (a0 + a1) + (a2 + a3);
        sum = warp_sum(sum);
        if (lane == 0) partials[(int64_t)j * nslabs + slab] = sum;
    }
}

// ---------------------------------------------------------------------------
// second-stage reduction: one block per output j
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
reduce_kernel(const double* __restrict__ partials, int np, double* __restrict__ out,
              int accumulate, double* __restrict__ out2, int do_sqrt) {
    const int j = blockIdx.x;
    const double* p = partials + (int64_t)j * np;
    // eight independent partial sums per thread: the loads of a thread are in flight together (one block
    // sums up to tens of thousands of per-CTA partials of an SpMV epilogue; a single dependent chain made
    // this kernel ~15 us at 512^2).  Fixed order: bit-identical reruns.
    double a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = 0.0;
    int i = threadIdx.x;
    const int stride = blockDim.x;
    for (; i + 7 * stride < np; i += 8 * stride) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] += p[i + u * stride];
    }
    for (; i < np; i += stride) a[0] += p[i];
    double v = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    v = block_sum(v);
    if (threadIdx.x == 0) {
        if (out2) out2[j] = v;
        if (do_sqrt) out[j] = sqrt(v);
        else out[j] = accumulate ? out[j] + v : v;
    }
}

// ---------------------------------------------------------------------------
// lincomb: out[r] = z[r] + s * sum_j V[r,j] c[j]   (sequential in j)
// Threads own two consecutive rows (one 16-byte load per column), eight
// independent column loads in flight per thread.
// ---------------------------------------------------------------------------
template <bool PUSH>
__global__ void __launch_bounds__(kBlock)
lincomb_kernel(const double* __restrict__ V, int64_t ld, int64_t n, int k,
               const double* __restrict__ c, double s, const double* __restrict__ z,
               double* __restrict__ out, const double* __restrict__ ref,
               double* __restrict__ stat, hg_out_list extra) {
    extern __shared__ double sc[];
    for (int j = threadIdx.x; j < k; j += blockDim.x) sc[j] = s * c[j];
    __syncthreads();
    const int64_t blk = blockIdx.x;
    const int64_t r = (blk * kBlock + threadIdx.x) * 2;
    double sq = 0.0;
    if (r + 1 < n) {
        double ax = 0.0, ay = 0.0;
        if (z) {
            const double2 zz = *reinterpret_cast<const double2*>(z + r);
            ax = zz.x;
            ay = zz.y;
        }
        const double* p = V + r;
        int j = 0;
        for (; j + 8 <= k; j += 8) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = ld_stream2(p + (int64_t)(j + u) * ld);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                ax = fma(sc[j + u], v[u].x, ax);
                ay = fma(sc[j + u], v[u].y, ay);
            }
        }
        for (; j < k; ++j) {
            const double2 v = ld_stream2(p + (int64_t)j * ld);
            ax = fma(sc[j], v.x, ax);
            ay = fma(sc[j], v.y, ay);
        }
        if (out) *reinterpret_cast<double2*>(out + r) = make_double2(ax, ay);
        if (PUSH) {  // the same rows into further (peer-GPU) destinations
#pragma unroll 4
            for (int d = 0; d < extra.n; ++d) *reinterpret_cast<double2*>(extra.p[d] + r) = make_double2(ax, ay);
        }
        if (stat) {
            double dx = ax, dy = ay;
            if (ref) {
                const double2 rr = *reinterpret_cast<const double2*>(ref + r);
                dx -= rr.x;
                dy -= rr.y;
            }
            sq = dx * dx + dy * dy;
        }
    } else if (r < n) {  // odd tail row
        double ax = z ? z[r] : 0.0;
        for (int j = 0; j < k; ++j) ax = fma(sc[j], ld_stream(V + (int64_t)j * ld + r), ax);
        if (out) out[r] = ax;
        if (PUSH)
            for (int d = 0; d < extra.n; ++d) extra.p[d][r] = ax;
        if (stat) {
            const double dx = ref ? ax - ref[r] : ax;
            sq = dx * dx;
        }
    }
    if (stat) {
        const double t = block_sum(sq);
        if (threadIdx.x == 0) stat[blk] = t;
    }
}


// ---------------------------------------------------------------------------
// Iterate + both histories of one hybrid iteration in ONE launch (hybrid_ba_gmres_rtp.m:30-33,
// hybrid_ab_gmres_rtp.m:33-36 with A*x taken from the cached columns T = A*Q_k):
//   job 0 (blocks [0, g0)):      x = Q y           and  sum (x - x_true)^2
//   job 1 (blocks [g0, g0+g1)):  r = b - T y (not stored)  and  sum r^2
// then the LAST block to finish (ticket counter) adds the per-block sums of each job in block order and
// writes out[0] = ||x - x_true||, out[1] = ||r||: no second-stage kernels, fixed summation order.
// ---------------------------------------------------------------------------
struct hg_iter_job {
    const double* V;    // n x k column-major
    const double* c;    // k coefficients (device)
    int k;
    int64_t ld, n;
    double s;           // scale of the combination (+1: x = V y, -1: r = z - V y)
    const double* z;    // added vector or nullptr
    double* out;        // result or nullptr
    const double* ref;  // subtracted before squaring, or nullptr
    int grid;           // blocks of this job
};

__global__ void __launch_bounds__(kBlock)
iterate_kernel(hg_iter_job j0, hg_iter_job j1, double* __restrict__ stat,
               unsigned int* __restrict__ ticket, double* __restrict__ out2) {
    extern __shared__ double sc[];
    __shared__ bool s_last;
    const bool second = (int)blockIdx.x >= j0.grid;
    const hg_iter_job& J = second ? j1 : j0;
    const int64_t blk = second ? (int64_t)blockIdx.x - j0.grid : blockIdx.x;
    const int k = J.k;
    for (int j = threadIdx.x; j < k; j += blockDim.x) sc[j] = J.s * J.c[j];
    __syncthreads();
    const int64_t r = (blk * kBlock + threadIdx.x) * 2;
    const int64_t n = J.n, ld = J.ld;
    double sq = 0.0;
    if (r + 1 < n) {
        double ax = 0.0, ay = 0.0;
        if (J.z) {
            const double2 zz = *reinterpret_cast<const double2*>(J.z + r);
            ax = zz.x;
            ay = zz.y;
        }
        const double* p = J.V + r;
        int j = 0;
        for (; j + 8 <= k; j += 8) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = ld_stream2(p + (int64_t)(j + u) * ld);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                ax = fma(sc[j + u], v[u].x, ax);
                ay = fma(sc[j + u], v[u].y, ay);
            }
        }
        for (; j < k; ++j) {
            const double2 v = ld_stream2(p + (int64_t)j * ld);
            ax = fma(sc[j], v.x, ax);
            ay = fma(sc[j], v.y, ay);
        }
        if (J.out) *reinterpret_cast<double2*>(J.out + r) = make_double2(ax, ay);
        double dx = ax, dy = ay;
        if (J.ref) {
            const double2 rr = *reinterpret_cast<const double2*>(J.ref + r);
            dx -= rr.x;
            dy -= rr.y;
        }
        sq = dx * dx + dy * dy;
    } else if (r < n) {
        double ax = J.z ? J.z[r] : 0.0;
        for (int j = 0; j < k; ++j) ax = fma(sc[j], ld_stream(J.V + (int64_t)j * ld + r), ax);
        if (J.out) J.out[r] = ax;
        const double dx = J.ref ? ax - J.ref[r] : ax;
        sq = dx * dx;
    }
    const double t = block_sum(sq);
    if (threadIdx.x == 0) {
        stat[blockIdx.x] = t;
        __threadfence();
        const unsigned int done = atomicAdd(ticket, 1u);
        s_last = done == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {  // every block's partial is visible (fence before its ticket); sum in block order
        __threadfence();
        for (int job = 0; job < 2; ++job) {
            const int lo = job ? j0.grid : 0, cnt = job ? j1.grid : j0.grid;
            double v = 0.0;
            for (int i = threadIdx.x; i < cnt; i += blockDim.x) v += __ldcg(stat + lo + i);
            v = block_sum(v);
            if (threadIdx.x == 0) out2[job] = sqrt(v);
            __syncthreads();
        }
        if (threadIdx.x == 0) *ticket = 0u;  // ready for the next launch on this stream
    }
}

// ---------------------------------------------------------------------------
// vector kernels (grid-stride free: one element pair per thread)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
scale_div_kernel(double* __restrict__ v, int64_t n, const double* __restrict__ d_div) {
    const double d = *d_div;
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < n) v[i] = v[i] / d;
}

__global__ void __launch_bounds__(kBlock)
axpby_kernel(int64_t n, double a, const double* __restrict__ x, double b,
             const double* __restrict__ y, double* __restrict__ out,
             const double* __restrict__ ref, double* __restrict__ stat) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    double sq = 0.0;
    if (i < n) {
        double o = a * x[i];
        if (y) o += b * y[i];
        if (out) out[i] = o;
        if (stat) {
            const double d = ref ? o - ref[i] : o;
            sq = d * d;
        }
    }
    if (stat) {
        const double t = block_sum(sq);
        if (threadIdx.x == 0) stat[blockIdx.x] = t;
    }
}

// True residual of LSQR on [A; sqrt(lambda) I] without a product with A: the Golub-Kahan relation gives
// A v_k = alpha_k u_k + beta_{k+1} u_{k+1} (top block), so with d_k = A w_k and r_k = b - A x_k
//   d_k = A v_k - cprev d_{k-1}   (w_k = v_k - (theta_k / rho_{k-1}) w_{k-1}, hybrid_lsqr_solver.m:40)
//   r_k = r_{k-1} - step d_k      (x_k = x_{k-1} + (phi_k / rho_k) w_k,       hybrid_lsqr_solver.m:39)
// stat: partial sums of r_k^2 (hybrid_lsqr_solver.m:43 takes norm(b - A*x))
// LSMR (d2 != nullptr) updates x along hbar_k = h_k - c0 hbar_{k-1} with h_k in the role of w_k (lsmr_solver.m:61-67):
//   d_k = A h_k as above,  d2_k = A hbar_k = d_k - c0 d2_{k-1},  r_k = r_{k-1} - step d2_k
__global__ void __launch_bounds__(kBlock)
gkb_resid_kernel(int64_t n, const double* __restrict__ u_old, double a, const double* __restrict__ u_new, double b,
                 double* __restrict__ d, double cprev, double* __restrict__ d2, double c0, int first,
                 double* __restrict__ r, double step, double* __restrict__ stat) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    double sq = 0.0;
    if (i < n) {
        double dn = a * u_old[i] + b * u_new[i];
        if (!first) dn -= cprev * d[i];
        d[i] = dn;
        if (d2) {
            if (!first) dn -= c0 * d2[i];
            d2[i] = dn;
        }
        const double rn = r[i] - step * dn;
        r[i] = rn;
        sq = rn * rn;
    }
    const double t = block_sum(sq);
    if (threadIdx.x == 0) stat[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kBlock)
lsqr_update_kernel(int64_t n, double* __restrict__ x, double* __restrict__ w,
                   const double* __restrict__ v, double c1, double c2,
                   const double* __restrict__ ref, double* __restrict__ stat) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    double sq = 0.0;
    if (i < n) {
        const double wi = w[i];
        const double xi = x[i] + c1 * wi;  // hybrid_lsqr_solver.m:39
        x[i] = xi;
        w[i] = v[i] - c2 * wi;  // hybrid_lsqr_solver.m:40
        if (stat) {
            const double d = ref ? xi - ref[i] : xi;
            sq = d * d;
        }
    }
    if (stat) {
        const double t = block_sum(sq);
        if (threadIdx.x == 0) stat[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kBlock)
lsmr_update_kernel(int64_t n, double* __restrict__ x, double* __restrict__ h,
                   double* __restrict__ hbar, const double* __restrict__ v, int first, double c0,
                   double c1, double c2, const double* __restrict__ ref,
                   double* __restrict__ stat) {
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    double sq = 0.0;
    if (i < n) {
        const double hi = h[i];
        const double hb = first ? hi : hi - c0 * hbar[i];  // lsmr_solver.m:61-65
        hbar[i] = hb;
        const double xi = x[i] + c1 * hb;  // :66
        x[i] = xi;
        h[i] = v[i] - c2 * hi;  // :67
        if (stat) {
            const double d = ref ? xi - ref[i] : xi;
            sq = d * d;
        }
    }
    if (stat) {
        const double t = block_sum(sq);
        if (threadIdx.x == 0) stat[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kBlock)
sumsq_kernel(const double* __restrict__ x, int64_t n, int per_block, double* __restrict__ stat) {
    const int64_t base = (int64_t)blockIdx.x * per_block;
    const int64_t end = min(n, base + per_block);
    double sq = 0.0;
    for (int64_t i = base + threadIdx.x; i < end; i += kBlock) {
        const double v = x[i];
        sq = fma(v, v, sq);
    }
    const double t = block_sum(sq);
    if (threadIdx.x == 0) stat[blockIdx.x] = t;
}

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace

// ===========================================================================
// host wrappers
// ===========================================================================

static double spmv_bytes(const hg_matrix* m, const hg_spmv_epilogue& ep, bool store) {
    // (8+iw)*nnz + pw*(r+1) + 8c + 8r (+8r per epilogue vector read); iw = 4 bytes per
    // column index, or 2 + 4/group when the matrix runs with 16-bit offsets (group = 128 or 32 entries)
    double idx = 4.0 * (double)m->nnz;
    if (m->sell_state > 0 && m->sell_col8) idx = 1.0 * (double)m->nnz + 16.0 * (double)(m->sell_entries / 128);
    else if (m->sell_state > 0 && m->sell_col16) idx = 2.0 * (double)m->nnz + 4.0 * (double)(m->sell_entries / 128);
    else if (m->sell_state <= 0 && m->grp_state > 0 && m->grp_d16) idx = 2.0 * (double)m->nnz + 128.0 * (double)m->grp_groups;  // one 32-lane checkpoint read per warp
    else if (m->sell_state <= 0 && m->csr16_state > 0) idx = 2.0 * (double)m->nnz + 4.0 * (double)m->csr_groups + 8.0 * (double)m->rows;
    double b = 8.0 * (double)m->nnz + idx + 8.0 * (double)(m->rows + 1) + 8.0 * (double)m->cols;
    if (store) b += 8.0 * (double)m->rows;
    if (ep.z1) b += 8.0 * (double)m->rows;
    if (ep.z2) b += 8.0 * (double)m->rows;
    if (ep.ref) b += 8.0 * (double)m->rows;
    return b;
}

// matrix-stream bytes of one product with m in the form it runs with (values + indices + pointers)
double hg_spmv_stream_bytes(const hg_matrix* m) {
    hg_spmv_epilogue ep;
    return spmv_bytes(m, ep, false) - 8.0 * (double)m->cols;
}

int hg_k_spmv(hg_ctx* ctx, const hg_matrix* m, const double* x, double* y,
              const hg_spmv_epilogue& ep, int* nparts) {
    if (nparts) *nparts = 0;
    if (m->rows == 0) return HG_OK;
    if (hg_spmv_stream_eligible(m)) return hg_k_spmv_stream(ctx, m, x, y, ep, nparts);
    if ((hg_spmv_mode() == 0 || hg_spmv_mode() == 3) && hg_sell_ready(ctx, m))
        return hg_k_spmv_sell(ctx, m, x, y, ep, spmv_bytes(m, ep, y != nullptr), nparts);
    if (m->tpr == 32 && hg_spmv_mode() == 0 && hg_group_ready(ctx, m))
        return hg_k_spmv_group(ctx, m, x, y, ep, spmv_bytes(m, ep, y != nullptr), nparts);
    if (m->tpr == 32 && hg_spmv_mode() == 0 && hg_idx16_csr_enabled() && hg_csr16_ready(ctx, m))
        return hg_k_spmv_csr16(ctx, m, x, y, ep, spmv_bytes(m, ep, y != nullptr), nparts);
    const int tpr = m->tpr;
    const int rpb = kBlock / tpr;
    const int64_t grid = cdiv(m->rows, rpb);
    HG_REQUIRE(grid < (int64_t)2147483647, "spmv: too many rows for one launch");
    if (nparts && ep.stat) *nparts = (int)grid;
    hg_launch_scope scope(ctx, HG_K_SPMV, spmv_bytes(m, ep, y != nullptr));
#define HG_SPMV_CASE(T)                                                                       \
    case T:                                                                                   \
        spmv_csr_kernel<T><<<(unsigned)grid, kBlock, 0, ctx->stream>>>(                       \
            m->rows, m->rowptr, m->colind, m->vals, x, y, ep.alpha, ep.z1, ep.g1, ep.z2,      \
            ep.g2, ep.ref, ep.stat);                                                          \
        break;
    switch (tpr) {
        HG_SPMV_CASE(2)
        HG_SPMV_CASE(4)
        HG_SPMV_CASE(8)
        HG_SPMV_CASE(16)
        HG_SPMV_CASE(32)
        default:
            hg_set_error("spmv: bad threads-per-row %d", tpr);
            return HG_ERR_INVALID;
    }
#undef HG_SPMV_CASE
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_multidot_slab_rows(const hg_ctx* ctx, int64_t n) {
    // aim for ~4 CTAs per SM; slab a multiple of 256 rows in [256, 4096]
    int64_t target = cdiv(n, (int64_t)ctx->sm_count * 4);
    int64_t R = cdiv(target, 256) * 256;
    if (R < 256) R = 256;
    if (R > 4096) R = 4096;
    return (int)R;
}

int hg_multidot_nslabs(const hg_ctx* ctx, int64_t n) {
    return (int)cdiv(n, hg_multidot_slab_rows(ctx, n));
}

int hg_k_multidot(hg_ctx* ctx, const double* V, int64_t ld, int64_t n, int k, const double* w,
                  double* partials, int* nslabs) {
    const int R = hg_multidot_slab_rows(ctx, n);
    const int ns = (int)cdiv(n, R);
    if (nslabs) *nslabs = ns;
    if (k <= 0 || n <= 0) return HG_OK;
    hg_launch_scope scope(ctx, HG_K_MULTIDOT, 8.0 * (double)n * (double)(k + 1));
    multidot_kernel<<<ns, kDotWarps * 32, R * sizeof(double), ctx->stream>>>(V, ld, n, k, w,
                                                                              partials, ns, R);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_reduce(hg_ctx* ctx, const double* partials, int np, int k, double* out, bool accumulate,
                double* out2, bool do_sqrt) {
    if (k <= 0) return HG_OK;
    hg_launch_scope scope(ctx, HG_K_REDUCE, 8.0 * (double)np * (double)k);
    reduce_kernel<<<k, kBlock, 0, ctx->stream>>>(partials, np, out, accumulate ? 1 : 0, out2,
                                                  do_sqrt ? 1 : 0);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_lincomb_push(hg_ctx* ctx, const double* V, int64_t ld, int64_t n, int k, const double* c,
                      double s, const double* z, double* out, const double* ref, double* stat,
                      int* nparts, const hg_out_list* extra) {
    const int64_t grid = cdiv(cdiv(n, 2), kBlock);
    if (nparts) *nparts = stat ? (int)grid : 0;
    if (n <= 0) return HG_OK;
    double bytes = 8.0 * (double)n * (double)k;
    if (z) bytes += 8.0 * (double)n;
    if (out) bytes += 8.0 * (double)n;
    if (ref) bytes += 8.0 * (double)n;
    if (extra) bytes += 8.0 * (double)n * extra->n;
    hg_launch_scope scope(ctx, HG_K_LINCOMB, bytes);
    const size_t smem = (size_t)(k > 0 ? k : 1) * sizeof(double);
    if (extra && extra->n > 0)
        lincomb_kernel<true><<<(unsigned)grid, kBlock, smem, ctx->stream>>>(V, ld, n, k, c, s, z, out, ref, stat, *extra);
    else
        lincomb_kernel<false><<<(unsigned)grid, kBlock, smem, ctx->stream>>>(V, ld, n, k, c, s, z, out, ref, stat,
                                                                            hg_out_list());
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_lincomb(hg_ctx* ctx, const double* V, int64_t ld, int64_t n, int k, const double* c,
                 double s, const double* z, double* out, const double* ref, double* stat,
                 int* nparts) {
    return hg_k_lincomb_push(ctx, V, ld, n, k, c, s, z, out, ref, stat, nparts, nullptr);
}

// x = Q y (+ ||x - x_true||) and ||b - T y|| in one launch; out2[0] = error norm, out2[1] = residual norm
// (device doubles); stat needs cdiv(n, 512) + cdiv(m, 512) doubles; ticket is a zero-initialised device word.
int hg_k_iterate(hg_ctx* ctx, const double* Q, int64_t ldq, int64_t n, const double* T, int64_t ldt, int64_t m,
                 int k, const double* y, const double* b, double* x, const double* x_true, double* stat,
                 unsigned int* ticket, double* out2) {
    return hg_k_iterate2(ctx, Q, ldq, n, k, y, x, x_true, T, ldt, m, k, y, b, stat, ticket, out2);
}

// general form: job 0  x = V0 c0 (n0 rows, k0 columns), error vs x_true;  job 1  ||b - V1 c1|| (n1 rows, k1 columns)
int hg_k_iterate2(hg_ctx* ctx, const double* V0, int64_t ld0, int64_t n0, int k0, const double* c0, double* x,
                  const double* x_true, const double* V1, int64_t ld1, int64_t n1, int k1, const double* c1,
                  const double* b, double* stat, unsigned int* ticket, double* out2) {
    hg_iter_job j0{V0, c0, k0, ld0, n0, 1.0, nullptr, x, x_true, (int)cdiv(n0, 2 * kBlock)};
    hg_iter_job j1{V1, c1, k1, ld1, n1, -1.0, b, nullptr, nullptr, (int)cdiv(n1, 2 * kBlock)};
    const int grid = j0.grid + j1.grid;
    if (grid == 0) return HG_OK;
    hg_launch_scope scope(ctx, HG_K_LINCOMB, 8.0 * ((double)k0 * (double)n0 + (double)k1 * (double)n1) +
                                                 24.0 * (double)n0 + 8.0 * (double)n1);
    const int kmax = k0 > k1 ? k0 : k1;
    const size_t smem = (size_t)(kmax > 0 ? kmax : 1) * sizeof(double);
    iterate_kernel<<<(unsigned)grid, kBlock, smem, ctx->stream>>>(j0, j1, stat, ticket, out2);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_scale_div(hg_ctx* ctx, double* v, int64_t n, const double* d_div) {
    if (n <= 0) return HG_OK;
    hg_launch_scope scope(ctx, HG_K_VECTOR, 16.0 * (double)n);
    scale_div_kernel<<<(unsigned)cdiv(n, kBlock), kBlock, 0, ctx->stream>>>(v, n, d_div);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_axpby(hg_ctx* ctx, int64_t n, double a, const double* x, double b, const double* y,
               double* out, const double* ref, double* stat, int* nparts) {
    const int64_t grid = cdiv(n, kBlock);
    if (nparts) *nparts = stat ? (int)grid : 0;
    if (n <= 0) return HG_OK;
    double bytes = 8.0 * (double)n * (1 + (y ? 1 : 0) + (out ? 1 : 0) + (ref ? 1 : 0));
    hg_launch_scope scope(ctx, HG_K_VECTOR, bytes);
    axpby_kernel<<<(unsigned)grid, kBlock, 0, ctx->stream>>>(n, a, x, b, y, out, ref, stat);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_gkb_resid(hg_ctx* ctx, int64_t n, const double* u_old, double a, const double* u_new, double b, double* d,
                   double cprev, double* d2, double c0, bool first, double* r, double step, double* stat,
                   int* nparts) {
    const int64_t grid = cdiv(n, kBlock);
    if (nparts) *nparts = (int)grid;
    if (n <= 0) return HG_OK;
    hg_launch_scope scope(ctx, HG_K_VECTOR, 8.0 * (double)n * (d2 ? 8 : 6));
    gkb_resid_kernel<<<(unsigned)grid, kBlock, 0, ctx->stream>>>(n, u_old, a, u_new, b, d, cprev, d2, c0, first ? 1 : 0,
                                                                r, step, stat);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_lsqr_update(hg_ctx* ctx, int64_t n, double* x, double* w, const double* v, double c1,
                     double c2, const double* ref, double* stat, int* nparts) {
    const int64_t grid = cdiv(n, kBlock);
    if (nparts) *nparts = stat ? (int)grid : 0;
    if (n <= 0) return HG_OK;
    hg_launch_scope scope(ctx, HG_K_VECTOR, 8.0 * (double)n * (5 + (ref ? 1 : 0)));
    lsqr_update_kernel<<<(unsigned)grid, kBlock, 0, ctx->stream>>>(n, x, w, v, c1, c2, ref, stat);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_lsmr_update(hg_ctx* ctx, int64_t n, double* x, double* h, double* hbar, const double* v,
                     int first, double c0, double c1, double c2, const double* ref, double* stat,
                     int* nparts) {
    const int64_t grid = cdiv(n, kBlock);
    if (nparts) *nparts = stat ? (int)grid : 0;
    if (n <= 0) return HG_OK;
    hg_launch_scope scope(ctx, HG_K_VECTOR, 8.0 * (double)n * (7 + (ref ? 1 : 0)));
    lsmr_update_kernel<<<(unsigned)grid, kBlock, 0, ctx->stream>>>(n, x, h, hbar, v, first, c0, c1,
                                                                    c2, ref, stat);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_sumsq(hg_ctx* ctx, const double* x, int64_t n, double* stat, int* nparts) {
    const int per_block = 4096;
    const int64_t grid = n > 0 ? cdiv(n, per_block) : 1;
    if (nparts) *nparts = (int)grid;
    hg_launch_scope scope(ctx, HG_K_VECTOR, 8.0 * (double)n);
    sumsq_kernel<<<(unsigned)grid, kBlock, 0, ctx->stream>>>(x, n, per_block, stat);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_reduce_to_scalar(hg_ctx* ctx, int np, int slot, bool do_sqrt) {
    return hg_k_reduce(ctx, ctx->d_partials, np, 1, ctx->d_scalars + slot, false, nullptr, do_sqrt);
}

int hg_norm2_sync(hg_ctx* ctx, const double* x, int64_t n, double* out) {
    int np = 0;
    HG_TRY(hg_ensure_partials(ctx, (size_t)cdiv(n > 0 ? n : 1, 4096) + 1));
    HG_TRY(hg_k_sumsq(ctx, x, n, ctx->d_partials, &np));
    HG_TRY(hg_reduce_to_scalar(ctx, np, 0, false));
    HG_CUDA(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, sizeof(double), cudaMemcpyDeviceToHost,
                            ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = ctx->h_scalars[0];
    return HG_OK;
}
