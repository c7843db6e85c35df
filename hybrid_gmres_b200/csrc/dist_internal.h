// Internal declarations of the multi-GPU layer shared by dist.cu and dist_peer.cu.
#pragma once

#include "common.cuh"

typedef struct ncclComm* ncclComm_t;

constexpr int HG_MAX_PEERS = 16;

// Base addresses of every rank's symmetric workspace as mapped into THIS process
// (cudaIpcOpenMemHandle; base[rank] is the local allocation).  Passed to kernels by value.
struct hg_peer_tbl {
    char* base[HG_MAX_PEERS];
    int P;
    int rank;
};

// Offsets (identical on every rank) inside the symmetric workspace.
struct hg_peer_layout {
    size_t flags = 0;   // uint64 [kFlagRows][HG_MAX_PEERS]: row f, column p = epoch signalled by rank p
    size_t inbox = 0;   // 16-byte LL entries [kInboxSlots][kpad][P]: small all-reduce contributions
    size_t ypart = 0;   // double [n_pad]: this rank's partial B^p u_p (peers pull their slices)
    size_t qfull = 0;   // double [2][n_pad]: replicated basis vector (peers push their slices)
    size_t total = 0;
    int kpad = 0;
    int64_t n_pad = 0;
};

constexpr int kFlagRows = 4;
constexpr int kInboxSlots = 4;
enum { HG_FLAG_Y = 0, HG_FLAG_Q = 1, HG_FLAG_AR = 2 };

struct hg_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    hg_ctx* ctx = nullptr;
    // NVLink peer-memory transport (dist_peer.cu)
    int transport = 0;  // 0: NCCL collectives, 1: peer memory
    char* ws = nullptr;
    size_t ws_bytes = 0;
    hg_peer_tbl tbl = {};
    hg_peer_layout lay;
    uint64_t bar_seq[kFlagRows] = {0, 0, 0, 0};  // epochs of the flag rows (HG_FLAG_AR counts all-reduces)
    unsigned int* d_counter = nullptr;            // spare device word
    unsigned long long* h_err = nullptr;          // pinned + mapped: kernels report a barrier time-out here
    unsigned long long* d_err = nullptr;
    const void* ws_owner = nullptr;               // the sharded Arnoldi currently using the workspace
    char why[200] = "NCCL collectives";
};

// dist.cu: NCCL helpers used by the workspace exchange
int hg_nccl_allgather_bytes(hg_comm* c, const void* d_send, void* d_recv, size_t bytes_per_rank, cudaStream_t st);
int hg_nccl_allreduce_min(hg_comm* c, double* d_buf, cudaStream_t st);
int hg_nccl_barrier(hg_comm* c, cudaStream_t st);

// dist_peer.cu
int hg_dist_transport_wanted();  // 0 auto, 1 NCCL only, 2 peer memory required
// Collective: make a symmetric workspace for vectors of n_pad doubles / coefficient vectors of
// kmax+2 doubles and hand it to `owner`.  Returns true when the peer transport is usable.
bool hg_peer_acquire(hg_comm* c, int64_t n_pad, int kmax, const void* owner);
void hg_peer_release(hg_comm* c, const void* owner);
void hg_peer_destroy(hg_comm* c);
int hg_peer_check(hg_comm* c);  // HG_ERR_STATE if a kernel reported a barrier time-out

inline double* hg_peer_ypart(hg_comm* c) { return reinterpret_cast<double*>(c->ws + c->lay.ypart); }
inline double* hg_peer_qfull(hg_comm* c, int buf) {
    return reinterpret_cast<double*>(c->ws + c->lay.qfull) + (size_t)buf * c->lay.n_pad;
}

// all ranks: barrier on flag row `row` (signal every peer, wait for every peer)
int hg_k_peer_barrier(hg_comm* c, int row);
// signal every peer on flag row `row` (no wait): "everything queued before this is complete"
int hg_k_peer_signal(hg_comm* c, int row);
// signal: CTA 0 first announces "this rank's partial product is complete" to every peer (a new HG_FLAG_Y epoch; no
// separate hg_k_peer_signal launch).  Waits (in every CTA) for the latest HG_FLAG_Y signal of every rank, then
// w[r] = sum_p ypart_p[row0 + r] (+ shift * q_slice[r]) for this rank's slice, stored to w_out,
// fused with partials[j*nslabs + slab] = sum_slab V[:,j] .* w  (k may be 0)
int hg_k_pull_multidot(hg_comm* c, int64_t row0, const double* q_slice, double shift, double* w_out,
                       const double* V, int64_t ld, int64_t n_p, int k, double* partials, int* nslabs,
                       bool signal = false);
// out[j] = sum over ranks (rank order) of sum_i partials[j*np + i], j < k; identical bits on every
// rank.  acc (optional): acc[j] = accumulate ? acc[j] + out[j] : out[j].  do_sqrt: out[j] = sqrt(.)
// barrier: stores of earlier kernels to peer memory are visible to a rank once it has this result (system-scope
// fences on both sides; a plain all-reduce needs none, every word carries its epoch).  partials2/np2 (optional): the
// last column (j = k-1) is summed from this second array instead — two statistics in one exchange.
int hg_k_reduce_allreduce(hg_comm* c, const double* partials, int np, int k, double* out, double* acc,
                          bool accumulate, bool do_sqrt, bool barrier = false, const double* partials2 = nullptr,
                          int np2 = 0);
// a[0..na) /= *d_div and b[0..nb) /= *d_div (na, nb even)
int hg_k_scale2(hg_comm* c, double* a, int64_t na, double* b, int64_t nb, const double* d_div);
// destinations of this rank's rows [row0, row0+n_p) in every rank's replicated vector `buf`
void hg_peer_push_list(hg_comm* c, int buf, int64_t row0, hg_out_list* out);

// cgs2_step.cu: the whole sharded CGS2 step (pull reduce-scatter, three in-kernel all-reduces, push
// all-gather, normalisation) in one persistent cooperative kernel
int hg_k_cgs2_step_peer(hg_comm* c, const double* V, int64_t ld, int64_t n_p, int k, double* w0, double* w1,
                        double* qnext, double* Hcol, double* hcur, double* partials, int64_t row0,
                        const double* q_slice, double shift, int buf);
