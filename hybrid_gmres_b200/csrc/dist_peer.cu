// NVLink peer-memory transport of the sharded Arnoldi step: the collective
// that follows a kernel is done BY a kernel, over memory of the other GPUs mapped into this
// process (cudaIpc), instead of by a separate NCCL call.
//
//   B^p u_p  -> reduce-scatter -> +shift*q -> V'w     one kernel: every CTA of the first CGS2
//                                                     multi-dot PULLS its row slab of all P
//                                                     partial vectors over NVLink (summed in rank
//                                                     order), adds shift*q, keeps the slab in
//                                                     shared memory for the dots
//   partial dots -> all-reduce(k)                     one kernel: block j finishes sum j, stores it
//                                                     into every rank's inbox as two 8-byte
//                                                     {half, epoch} words (NCCL-LL style: the data
//                                                     carries its own flag, one NVLink one-way
//                                                     latency, no fence round trip), polls its own
//                                                     inbox row and sums in rank order
//                                                     (bit-identical on every rank)
//   w - Q h    -> all-gather                          the second CGS2 update PUSHES its (not yet
//                                                     normalised) rows into every rank's replicated
//                                                     copy; the norm all-reduce that follows is
//                                                     also the barrier for those stores; each rank
//                                                     then scales its full copy locally
// Synchronisation is by monotone epochs (st.release.sys / ld.acquire.sys flag words, or the epoch
// inside the LL words); every rank runs the same kernel sequence on its stream, so a wait can only
// be for a kernel a peer has already queued.  Waits are bounded: a time-out sets an error word
// the host reports.
#include <algorithm>

#include "dist_internal.h"

int hg_multidot_slab_rows(const hg_ctx* ctx, int64_t n);

namespace {

constexpr int kBlock = 256;
constexpr long long kSpinLimit = 40000000000LL;  // ~20 s of SM clocks

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void wait_flag(const unsigned long long* f, unsigned long long epoch,
                                          unsigned long long* err) {
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < epoch) {
        if (clock64() - t0 > kSpinLimit) {
            *err = 1ull;  // mapped host word
            __threadfence_system();
            return;
        }
        __nanosleep(32);
    }
}

// signal every peer, wait for every peer (lane p talks to rank p)
__device__ __forceinline__ void exchange_flags(const hg_peer_tbl& t, size_t row_off,
                                               unsigned long long epoch, unsigned long long* err) {
    const int p = threadIdx.x;
    if (p < t.P) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(t.base[p] + row_off) + t.rank, epoch);
        wait_flag(reinterpret_cast<const unsigned long long*>(t.base[t.rank] + row_off) + p, epoch, err);
    }
}

__global__ void peer_barrier_kernel(hg_peer_tbl t, size_t row_off, unsigned long long epoch,
                                    unsigned long long* err) {
    exchange_flags(t, row_off, epoch, err);
}

// "my previous kernel is complete": no wait here, the consumers wait (wait_all)
__global__ void peer_signal_kernel(hg_peer_tbl t, size_t row_off, unsigned long long epoch) {
    const int p = threadIdx.x;
    if (p < t.P) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(t.base[p] + row_off) + t.rank, epoch);
    }
}

// CTA-wide: until every rank has signalled `epoch` on flag row `row_off`
__device__ __forceinline__ void wait_all(const hg_peer_tbl& t, size_t row_off, unsigned long long epoch,
                                         unsigned long long* err, int skip = -1) {
    if (threadIdx.x < t.P && (int)threadIdx.x != skip)  // skip: this rank's own flag (stream order covers it)
        wait_flag(reinterpret_cast<const unsigned long long*>(t.base[t.rank] + row_off) + threadIdx.x, epoch, err);
    __syncthreads();
}

// ---------------------------------------------------------------------------
// reduce-scatter (pull) + shift + multi-dot.  Same slab / warp-per-column structure as
// multidot_kernel (kernels.cu); the w slab is produced here instead of read.
// ---------------------------------------------------------------------------
constexpr int kDotWarps = 8;

__global__ void __launch_bounds__(kDotWarps * 32)
pull_multidot_kernel(hg_peer_tbl t, size_t ypart_off, size_t flag_off, unsigned long long epoch, int signal,
                     unsigned long long* err, int64_t row0, const double* __restrict__ q_slice,
                     double shift, double* __restrict__ w_out, const double* __restrict__ V, int64_t ld,
                     int64_t n, int k, double* __restrict__ partials, int nslabs, int R) {
    extern __shared__ double sw[];
    // signal: this rank's B^p u_p (the previous kernel of the stream) is complete — CTA 0 tells every peer, no
    // separate launch; then every CTA waits until every rank has said so
    if (signal && blockIdx.x == 0 && threadIdx.x < t.P) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(t.base[threadIdx.x] + flag_off) + t.rank, epoch);
    }
    wait_all(t, flag_off, epoch, err, signal ? t.rank : -1);
    const int slab = blockIdx.x;
    const int64_t r0 = (int64_t)slab * R;
    const int len = (int)min((int64_t)R, n - r0);
    for (int i = 2 * threadIdx.x; i < R; i += 2 * blockDim.x) {  // R and n are even (multiples of 32)
        double2 s = make_double2(0.0, 0.0);
        if (i < len) {
            const int64_t g = row0 + r0 + i;
#pragma unroll 4
            for (int p = 0; p < t.P; ++p) {  // fixed rank order: every rank forms the same bits
                const double2 v = __ldcg(reinterpret_cast<const double2*>(
                    reinterpret_cast<const double*>(t.base[p] + ypart_off) + g));
                s.x += v.x;
                s.y += v.y;
            }
            if (shift != 0.0) {
                const double2 qv = *reinterpret_cast<const double2*>(q_slice + r0 + i);
                s.x += shift * qv.x;
                s.y += shift * qv.y;
            }
            *reinterpret_cast<double2*>(w_out + r0 + i) = s;
        }
        sw[i] = s.x;
        sw[i + 1] = s.y;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool full = (len == R);
    for (int j = warp; j < k; j += kDotWarps) {
        const double* col = V + (int64_t)j * ld + r0;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if (full) {
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
// second-stage reduction + one-shot all-reduce, LL protocol: a double travels as two 8-byte
// words {lo32 | epoch32}, {hi32 | epoch32}; 8-byte stores are single-copy atomic, so a word whose
// epoch matches carries valid data and no fence / separate flag is needed.  Inbox entry
// (slot, j, src) is 16 bytes at ((slot*kpad + j)*P + src).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kBlock)
reduce_allreduce_kernel(const double* __restrict__ partials, int np, int k, hg_peer_tbl t,
                        size_t inbox_off, unsigned int epoch, double* __restrict__ out,
                        double* __restrict__ acc, int accumulate, int do_sqrt, int barrier,
                        const double* __restrict__ partials2, int np2, unsigned long long* err) {
    __shared__ double s_red[32];
    __shared__ double s_part[HG_MAX_PEERS];
    const int j = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the last column may come from a second partials array (two statistics in one exchange)
    const bool second = partials2 != nullptr && j == k - 1;
    const double* p = second ? partials2 : partials + (int64_t)j * np;
    const int cnt = second ? np2 : np;
    double v = 0.0;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) v += p[i];
    v = warp_sum(v);
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (kBlock >> 5) ? s_red[lane] : 0.0;
        v = warp_sum(v);  // every lane of warp 0 holds this rank's sum j
        if (lane < t.P) {
            // barrier: earlier kernels' stores to peer memory (the pushed basis rows) are ordered before the
            // words that announce them.  A plain all-reduce needs no fence: a word carries its own epoch.
            if (barrier) __threadfence_system();
            const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(t.base[lane] + inbox_off) +
                                      ((size_t)j * t.P + t.rank) * 2;
            st_relaxed_sys(dst, (bits << 32) | epoch);
            st_relaxed_sys(dst + 1, (bits & 0xffffffff00000000ull) | epoch);
            // poll my own inbox for rank `lane`'s contribution
            const unsigned long long* src = reinterpret_cast<const unsigned long long*>(t.base[t.rank] + inbox_off) +
                                            ((size_t)j * t.P + lane) * 2;
            const long long t0 = clock64();
            unsigned long long w0, w1;
            bool ok = true;
            for (;;) {
                w0 = ld_relaxed_sys(src);
                w1 = ld_relaxed_sys(src + 1);
                if ((unsigned int)w0 == epoch && (unsigned int)w1 == epoch) break;
                if (clock64() - t0 > kSpinLimit) {
                    *err = 1ull;
                    __threadfence_system();
                    ok = false;
                    break;
                }
            }
            // acquire side of the exchange: this all-reduce is also the barrier that makes the rows peers
            // pushed into our replicated vector visible to the kernels that follow, so order every later
            // read after the observed epochs (the sender fences before its store)
            if (barrier) __threadfence_system();
            s_part[lane] = ok ? __longlong_as_double((long long)((w0 >> 32) | (w1 & 0xffffffff00000000ull))) : 0.0;
        }
        __syncwarp();
        if (lane == 0) {
            double s = 0.0;
            for (int r = 0; r < t.P; ++r) s += s_part[r];  // rank order: identical bits on every rank
            if (do_sqrt) s = sqrt(s);
            out[j] = s;
            if (acc) acc[j] = accumulate ? acc[j] + s : s;
        }
    }
}

// v[i] /= *d_div over a local vector (the replicated q after its rows arrived, and the basis slice)
__global__ void __launch_bounds__(kBlock)
scale2_kernel(double* __restrict__ a, int64_t na, double* __restrict__ b, int64_t nb,
              const double* __restrict__ d_div) {
    const int64_t i = ((int64_t)blockIdx.x * kBlock + threadIdx.x) * 2;
    const double d = *d_div;
    if (i < na) {
        double2 x = *reinterpret_cast<const double2*>(a + i);
        x.x = x.x / d;  // division, as `v / H(k+1,k)` in the reference (hybrid_ab_gmres_rtp.m:26)
        x.y = x.y / d;
        *reinterpret_cast<double2*>(a + i) = x;
    } else if (i - na < nb) {
        double2 x = *reinterpret_cast<const double2*>(b + (i - na));
        x.x = x.x / d;
        x.y = x.y / d;
        *reinterpret_cast<double2*>(b + (i - na)) = x;
    }
}

hg_peer_layout make_layout(int64_t n_pad, int kmax, int P) {
    hg_peer_layout l;
    l.kpad = (int)align_up((size_t)kmax + 2, 32);
    l.n_pad = n_pad;
    l.flags = 0;
    l.inbox = align_up((size_t)kFlagRows * HG_MAX_PEERS * 8, 256);
    l.ypart = align_up(l.inbox + (size_t)kInboxSlots * P * l.kpad * 16, 256);
    l.qfull = l.ypart + (size_t)n_pad * 8;
    l.total = l.qfull + 2 * (size_t)n_pad * 8;
    return l;
}

static int g_dist_transport = -1;

}  // namespace

int hg_dist_transport_wanted() {
    if (g_dist_transport < 0) {
        const char* e = getenv("HG_DIST");
        g_dist_transport = 0;
        if (e && strcmp(e, "nccl") == 0) g_dist_transport = 1;
        if (e && strcmp(e, "peer") == 0) g_dist_transport = 2;
    }
    return g_dist_transport;
}
void hg_dist_transport_set(int v) { g_dist_transport = v; }

static void close_peers(hg_comm* c) {
    for (int p = 0; p < c->nranks && p < HG_MAX_PEERS; ++p) {
        if (p != c->rank && c->tbl.base[p]) cudaIpcCloseMemHandle(c->tbl.base[p]);
        c->tbl.base[p] = nullptr;
    }
}

// Collective.  (Re)allocates the symmetric workspace when it is too small and maps every peer's.
static bool ensure_ws(hg_comm* c, size_t bytes) {
    hg_ctx* ctx = c->ctx;
    cudaStream_t st = ctx->stream;
    const int P = c->nranks;
    if (c->ws && c->ws_bytes >= bytes) return true;
    // nobody may still be using (or have mapped) the old workspace
    if (cudaStreamSynchronize(st) != cudaSuccess) return false;
    if (P > 1 && hg_nccl_barrier(c, st) != HG_OK) return false;
    if (c->ws) {
        close_peers(c);
        if (P > 1 && hg_nccl_barrier(c, st) != HG_OK) return false;  // every rank unmapped mine
        cudaFree(c->ws);
        c->ws = nullptr;
        c->ws_bytes = 0;
    }
    double ok = 1.0;
    const size_t want = align_up(bytes + bytes / 4, 1 << 20);  // head-room: fewer collective re-allocations
    if (cudaMalloc(&c->ws, want) != cudaSuccess) {
        cudaGetLastError();
        c->ws = nullptr;
        ok = 0.0;
    }
    if (ok != 0.0 && cudaMemsetAsync(c->ws, 0, want, st) != cudaSuccess) ok = 0.0;
    if (!c->d_counter) {
        if (cudaMalloc(&c->d_counter, 64) != cudaSuccess || cudaMemsetAsync(c->d_counter, 0, 64, st) != cudaSuccess)
            ok = 0.0;
    }
    if (!c->h_err) {
        if (cudaHostAlloc(&c->h_err, 64, cudaHostAllocMapped) != cudaSuccess) ok = 0.0;
        else {
            c->h_err[0] = 0;
            if (cudaHostGetDevicePointer(&c->d_err, c->h_err, 0) != cudaSuccess) ok = 0.0;
        }
    }
    for (int i = 0; i < kFlagRows; ++i) c->bar_seq[i] = 0;
    c->tbl.P = P;
    c->tbl.rank = c->rank;
    for (int p = 0; p < HG_MAX_PEERS; ++p) c->tbl.base[p] = nullptr;
    c->tbl.base[c->rank] = c->ws;
    if (P > 1) {
        // exchange the IPC handles through NCCL (64 bytes per rank), then agree on the outcome
        char* d_h = nullptr;
        std::vector<cudaIpcMemHandle_t> handles((size_t)P);
        cudaIpcMemHandle_t mine;
        memset(&mine, 0, sizeof(mine));
        if (ok != 0.0 && cudaIpcGetMemHandle(&mine, c->ws) != cudaSuccess) {
            cudaGetLastError();
            ok = 0.0;
        }
        bool staged = cudaMalloc(&d_h, sizeof(mine) * (size_t)(P + 1)) == cudaSuccess;
        if (!staged) {
            cudaGetLastError();
            return false;  // cannot even talk to the peers: a fatal allocation failure
        }
        cudaMemcpyAsync(d_h + sizeof(mine) * (size_t)P, &mine, sizeof(mine), cudaMemcpyHostToDevice, st);
        int rc = hg_nccl_allgather_bytes(c, d_h + sizeof(mine) * (size_t)P, d_h, sizeof(mine), st);
        if (rc == HG_OK &&
            cudaMemcpyAsync(handles.data(), d_h, sizeof(mine) * (size_t)P, cudaMemcpyDeviceToHost, st) != cudaSuccess)
            rc = HG_ERR_CUDA;
        if (rc == HG_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = HG_ERR_CUDA;
        if (rc != HG_OK) ok = 0.0;
        if (ok != 0.0) {
            for (int p = 0; p < P; ++p) {
                if (p == c->rank) continue;
                void* ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, handles[(size_t)p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    snprintf(c->why, sizeof(c->why), "NCCL collectives (cudaIpcOpenMemHandle of rank %d failed: %s)", p,
                             cudaGetErrorString(cudaGetLastError()));
                    ok = 0.0;
                    break;
                }
                c->tbl.base[p] = static_cast<char*>(ptr);
            }
        }
        // all ranks take the same decision
        double* d_ok = reinterpret_cast<double*>(d_h);
        cudaMemcpyAsync(d_ok, &ok, 8, cudaMemcpyHostToDevice, st);
        double all_ok = 0.0;
        if (hg_nccl_allreduce_min(c, d_ok, st) == HG_OK &&
            cudaMemcpyAsync(&all_ok, d_ok, 8, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
            cudaStreamSynchronize(st) == cudaSuccess) {
            ok = all_ok;
        } else {
            ok = 0.0;
        }
        cudaFree(d_h);
    } else if (cudaStreamSynchronize(st) != cudaSuccess) {
        ok = 0.0;
    }
    if (ok == 0.0) {  // the same verdict on every rank (all-reduced above)
        close_peers(c);
        if (P > 1) hg_nccl_barrier(c, st);  // nobody frees memory a peer still has mapped
        if (c->ws) cudaFree(c->ws);
        c->ws = nullptr;
        c->ws_bytes = 0;
        cudaGetLastError();
        return false;
    }
    c->ws_bytes = want;
    return true;
}

bool hg_peer_acquire(hg_comm* c, int64_t n_pad, int kmax, const void* owner) {
    const int want = hg_dist_transport_wanted();
    c->transport = 0;
    if (want == 1) {
        snprintf(c->why, sizeof(c->why), "NCCL collectives (dist_transport = nccl)");
        return false;
    }
    if (c->nranks > HG_MAX_PEERS || !c->ctx) {
        snprintf(c->why, sizeof(c->why), "NCCL collectives (more than %d ranks)", HG_MAX_PEERS);
        return false;
    }
    if (c->ws_owner && c->ws_owner != owner) {  // one sharded Arnoldi at a time owns the workspace
        snprintf(c->why, sizeof(c->why), "NCCL collectives (peer workspace in use by another sharded Arnoldi)");
        return false;
    }
    const hg_peer_layout lay = make_layout(n_pad, kmax, c->nranks);
    if (!ensure_ws(c, lay.total)) {
        if (strncmp(c->why, "NCCL collectives (cudaIpc", 25) != 0)
            snprintf(c->why, sizeof(c->why), "NCCL collectives (peer workspace could not be set up on every rank)");
        return false;
    }
    if (lay.kpad != c->lay.kpad || lay.n_pad != c->lay.n_pad) {
        // the regions move: words of the old vectors must not be mistaken for LL entries.  Nobody
        // may write into my inbox before it is cleared, hence the barrier (creation-time cost only).
        bool okz = cudaMemsetAsync(c->ws + lay.inbox, 0, lay.ypart - lay.inbox, c->ctx->stream) == cudaSuccess;
        if (okz) okz = c->nranks > 1 ? hg_nccl_barrier(c, c->ctx->stream) == HG_OK
                                     : cudaStreamSynchronize(c->ctx->stream) == cudaSuccess;
        if (!okz) {
            snprintf(c->why, sizeof(c->why), "NCCL collectives (peer workspace could not be cleared)");
            return false;
        }
    }
    c->lay = lay;
    c->ws_owner = owner;
    c->transport = 1;
    snprintf(c->why, sizeof(c->why), "NVLink peer memory (cudaIpc), %d ranks", c->nranks);
    return true;
}

void hg_peer_release(hg_comm* c, const void* owner) {
    if (c->ws_owner == owner) c->ws_owner = nullptr;
}

void hg_peer_destroy(hg_comm* c) {
    if (c->ws) {
        cudaStreamSynchronize(c->ctx->stream);
        if (c->nranks > 1) hg_nccl_barrier(c, c->ctx->stream);
        close_peers(c);
        if (c->nranks > 1) hg_nccl_barrier(c, c->ctx->stream);
        cudaFree(c->ws);
        c->ws = nullptr;
    }
    if (c->d_counter) cudaFree(c->d_counter);
    if (c->h_err) cudaFreeHost(c->h_err);
    c->d_counter = nullptr;
    c->h_err = nullptr;
}

int hg_peer_check(hg_comm* c) {
    if (c->h_err && c->h_err[0] != 0) {
        c->h_err[0] = 0;
        hg_set_error("peer transport: a flag wait timed out (a rank left the kernel sequence)");
        return HG_ERR_STATE;
    }
    return HG_OK;
}

int hg_k_peer_barrier(hg_comm* c, int row) {
    hg_ctx* ctx = c->ctx;
    const unsigned long long epoch = ++c->bar_seq[row];
    hg_launch_scope scope(ctx, HG_K_COMM, 16.0 * c->nranks);
    peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(c->tbl, c->lay.flags + (size_t)row * HG_MAX_PEERS * 8, epoch,
                                                    c->d_err);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_peer_signal(hg_comm* c, int row) {
    hg_ctx* ctx = c->ctx;
    const unsigned long long epoch = ++c->bar_seq[row];
    hg_launch_scope scope(ctx, HG_K_COMM, 8.0 * c->nranks);
    peer_signal_kernel<<<1, 32, 0, ctx->stream>>>(c->tbl, c->lay.flags + (size_t)row * HG_MAX_PEERS * 8, epoch);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_pull_multidot(hg_comm* c, int64_t row0, const double* q_slice, double shift, double* w_out,
                       const double* V, int64_t ld, int64_t n_p, int k, double* partials, int* nslabs, bool signal) {
    hg_ctx* ctx = c->ctx;
    if (signal) ++c->bar_seq[HG_FLAG_Y];
    const int R = hg_multidot_slab_rows(ctx, n_p);
    const int ns = (int)cdiv(n_p, R);
    if (nslabs) *nslabs = ns;
    if (n_p <= 0) {  // nothing to pull here, but the peers still wait for this rank's signal
        if (signal) {
            --c->bar_seq[HG_FLAG_Y];
            return hg_k_peer_signal(c, HG_FLAG_Y);
        }
        return HG_OK;
    }
    // waits for the epoch of the latest HG_FLAG_Y signal (sent here when `signal`, else by hg_k_peer_signal)
    // local reads: V_k and q slice; remote/local pulls: P slices; write w
    hg_launch_scope scope(ctx, HG_K_MULTIDOT, 8.0 * (double)n_p * (double)(k + c->nranks + 2));
    pull_multidot_kernel<<<ns, kDotWarps * 32, R * sizeof(double), ctx->stream>>>(
        c->tbl, c->lay.ypart, c->lay.flags + (size_t)HG_FLAG_Y * HG_MAX_PEERS * 8, c->bar_seq[HG_FLAG_Y],
        signal ? 1 : 0, c->d_err,
        row0, q_slice, shift, w_out, V, ld, n_p, k, partials, ns, R);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_reduce_allreduce(hg_comm* c, const double* partials, int np, int k, double* out, double* acc,
                          bool accumulate, bool do_sqrt, bool barrier, const double* partials2, int np2) {
    if (k <= 0) return HG_OK;
    hg_ctx* ctx = c->ctx;
    HG_REQUIRE(k <= c->lay.kpad, "peer all-reduce: %d coefficients exceed the inbox (%d)", k, c->lay.kpad);
    const unsigned long long seq = c->bar_seq[HG_FLAG_AR]++;
    const size_t inbox_off = c->lay.inbox + (size_t)(seq % kInboxSlots) * c->nranks * c->lay.kpad * 16;
    const unsigned int epoch = (unsigned int)(seq % 0xfffffffeull) + 1u;  // never 0 (= empty inbox)
    hg_launch_scope scope(ctx, HG_K_REDUCE, 8.0 * (double)np * (double)k + 32.0 * k * c->nranks);
    reduce_allreduce_kernel<<<k, kBlock, 0, ctx->stream>>>(partials, np, k, c->tbl, inbox_off, epoch, out, acc,
                                                            accumulate ? 1 : 0, do_sqrt ? 1 : 0, barrier ? 1 : 0,
                                                            partials2, np2, c->d_err);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

int hg_k_scale2(hg_comm* c, double* a, int64_t na, double* b, int64_t nb, const double* d_div) {
    if (na + nb <= 0) return HG_OK;
    hg_ctx* ctx = c->ctx;
    hg_launch_scope scope(ctx, HG_K_VECTOR, 16.0 * (double)(na + nb));
    scale2_kernel<<<(unsigned)cdiv(na + nb, 2 * kBlock), kBlock, 0, ctx->stream>>>(a, na, b, nb, d_div);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

void hg_peer_push_list(hg_comm* c, int buf, int64_t row0, hg_out_list* out) {
    out->n = c->nranks;
    for (int p = 0; p < c->nranks; ++p)
        out->p[p] = reinterpret_cast<double*>(c->tbl.base[p] + c->lay.qfull) + (size_t)buf * c->lay.n_pad + row0;
}
