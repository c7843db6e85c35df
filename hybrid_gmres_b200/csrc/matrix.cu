// Device-resident CSR matrices: upload from CSR / MATLAB CSC / dense, and a
// deterministic device transposition (B and A' are always explicit CSR — the
// hot path never does an atomics-based transposed product).
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kBlock = 256;
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

__global__ void widen_ptr_kernel(const int32_t* __restrict__ in, int64_t* __restrict__ out,
                                 int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

__global__ void narrow_idx_kernel(const int64_t* __restrict__ in, int32_t* __restrict__ out,
                                  int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (in[i] < 0 || in[i] > 2147483647LL) ? -1 : (int32_t)in[i];  // -1: caught by validate_csr
}

// Structural validation of caller-supplied CSR arrays: bit 0 rowptr[0] != 0 or rowptr[rows] != nnz,
// bit 1 a decreasing rowptr, bit 2 a column index outside [0, cols).  One pass, 4 bytes per entry.
__global__ void validate_csr_kernel(int64_t rows, int64_t cols, int64_t nnz, const int64_t* __restrict__ rowptr,
                                    const int32_t* __restrict__ colind, unsigned int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int b = 0;
    if (i == 0 && (rowptr[0] != 0 || rowptr[rows] != nnz)) b |= 1u;
    if (i < rows && rowptr[i] > rowptr[i + 1]) b |= 2u;
    if (i < nnz) {
        const int32_t c = colind[i];
        if (c < 0 || (int64_t)c >= cols) b |= 4u;
    }
    if (b) atomicOr(bad, b);
}

__global__ void dense_to_csr_kernel(int64_t rows, int64_t cols, const double* __restrict__ a,
                                    int64_t lda, int64_t* __restrict__ rowptr,
                                    int32_t* __restrict__ colind, double* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= rows) rowptr[i] = i * cols;
    if (i < rows * cols) {
        const int64_t r = i / cols, c = i % cols;
        colind[i] = (int32_t)c;
        vals[i] = a[r + c * lda];
    }
}

// ---- transposition ----------------------------------------------------------
__global__ void count_cols_kernel(const int32_t* __restrict__ colind, int64_t nnz,
                                  unsigned int* __restrict__ cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) atomicAdd(&cnt[colind[i]], 1u);  // integer atomics: result is deterministic
}

// one warp per input row scatters its entries; slot order within an output row
// is arbitrary here and canonicalised by sort_rows_kernel
__global__ void scatter_kernel(int64_t rows, const int64_t* __restrict__ rowptr,
                               const int32_t* __restrict__ colind, const double* __restrict__ vals,
                               const int64_t* __restrict__ tptr, unsigned int* __restrict__ cursor,
                               int32_t* __restrict__ tcol, double* __restrict__ tval) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    for (int64_t i = s + lane; i < e; i += 32) {
        const int c = colind[i];
        const unsigned int slot = atomicAdd(&cursor[c], 1u);
        const int64_t dst = tptr[c] + slot;
        tcol[dst] = (int32_t)row;
        tval[dst] = vals[i];
    }
}

__device__ __forceinline__ bool entry_less(int ka, double va, int kb, double vb) {
    if (ka != kb) return ka < kb;
    return __double_as_longlong(va) < __double_as_longlong(vb);  // total order => deterministic
}

// one CTA per output row: bitonic sort of (col, val) by col in shared memory
__global__ void __launch_bounds__(kBlock)
sort_rows_kernel(int64_t rows, const int64_t* __restrict__ tptr, int32_t* __restrict__ tcol,
                 double* __restrict__ tval, int cap) {
    extern __shared__ unsigned char smem_raw[];
    double* sv = reinterpret_cast<double*>(smem_raw);
    int* sk = reinterpret_cast<int*>(sv + cap);
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const int64_t s = tptr[row];
        const int len = (int)(tptr[row + 1] - s);
        if (len <= 1) continue;
        if (len > cap) continue;  // handled by sort_long_rows_kernel
        int P = 1;
        while (P < len) P <<= 1;
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            sk[i] = i < len ? tcol[s + i] : 0x7fffffff;
            sv[i] = i < len ? tval[s + i] : 0.0;
        }
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < P; i += blockDim.x) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const bool up = (i & k) == 0;
                        const int ka = sk[i], kb = sk[ixj];
                        const double va = sv[i], vb = sv[ixj];
                        const bool swap = up ? entry_less(kb, vb, ka, va) : entry_less(ka, va, kb, vb);
                        if (swap) {
                            sk[i] = kb;
                            sk[ixj] = ka;
                            sv[i] = vb;
                            sv[ixj] = va;
                        }
                    }
                }
                __syncthreads();
            }
        }
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
            tcol[s + i] = sk[i];
            tval[s + i] = sv[i];
        }
        __syncthreads();
    }
}

// rows longer than the shared-memory capacity: rank sort through scratch
__global__ void __launch_bounds__(kBlock)
sort_long_rows_kernel(int64_t rows, const int64_t* __restrict__ tptr, int32_t* __restrict__ tcol,
                      double* __restrict__ tval, int cap, int32_t* __restrict__ scratch_k,
                      double* __restrict__ scratch_v) {
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const int64_t s = tptr[row];
        const int64_t len = tptr[row + 1] - s;
        if (len <= cap) continue;
        for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
            const int ki = tcol[s + i];
            const double vi = tval[s + i];
            int64_t rank = 0;
            for (int64_t t = 0; t < len; ++t) {
                const int kt = tcol[s + t];
                const double vt = tval[s + t];
                if (entry_less(kt, vt, ki, vi) || (t < i && kt == ki &&
                    __double_as_longlong(vt) == __double_as_longlong(vi))) ++rank;
            }
            scratch_k[s + rank] = ki;
            scratch_v[s + rank] = vi;
        }
        __syncthreads();
        for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
            tcol[s + i] = scratch_k[s + i];
            tval[s + i] = scratch_v[s + i];
        }
        __syncthreads();
    }
}

__global__ void max_row_len_kernel(int64_t rows, const int64_t* __restrict__ ptr,
                                   unsigned long long* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) atomicMax(out, (unsigned long long)(ptr[i + 1] - ptr[i]));
}

// ---- permutation ------------------------------------------------------------
__global__ void perm_row_len_kernel(int64_t rows, const int64_t* __restrict__ rowptr,
                                    const int32_t* __restrict__ rowperm, int32_t* __restrict__ len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) {
        const int64_t src = rowperm ? rowperm[i] : i;
        len[i] = (int32_t)(rowptr[src + 1] - rowptr[src]);
    }
}

// one warp per output row: out row i = in row rowperm[i], column c relabelled to colinv[c]
__global__ void perm_copy_kernel(int64_t rows, const int64_t* __restrict__ rowptr,
                                 const int32_t* __restrict__ colind, const double* __restrict__ vals,
                                 const int32_t* __restrict__ rowperm, const int32_t* __restrict__ colinv,
                                 const int64_t* __restrict__ optr, int32_t* __restrict__ ocol,
                                 double* __restrict__ oval) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int64_t src = rowperm ? rowperm[row] : row;
    const int64_t s = rowptr[src], e = rowptr[src + 1], d = optr[row];
    for (int64_t i = s + lane; i < e; i += 32) {
        const int c = colind[i];
        ocol[d + (i - s)] = colinv ? colinv[c] : c;
        oval[d + (i - s)] = vals[i];
    }
}

}  // namespace

int hg_matrix_alloc(hg_ctx* ctx, int64_t rows, int64_t cols, int64_t nnz, hg_matrix** out) {
    *out = nullptr;
    HG_REQUIRE(rows >= 0 && cols >= 0 && nnz >= 0, "matrix: negative dimension");
    HG_REQUIRE(cols <= 2147483647LL, "matrix: more than 2^31-1 columns is not supported");
    hg_matrix* m = new (std::nothrow) hg_matrix();
    if (!m) {
        hg_set_error("matrix: out of host memory");
        return HG_ERR_NOMEM;
    }
    m->ctx = ctx;
    m->rows = rows;
    m->cols = cols;
    m->nnz = nnz;
    cudaError_t e = hg_dmalloc(ctx, &m->rowptr, (size_t)(rows + 1) * sizeof(int64_t));
    if (e == cudaSuccess) e = hg_dmalloc(ctx, &m->colind, (size_t)(nnz + kNnzPad) * sizeof(int32_t));
    if (e == cudaSuccess) e = hg_dmalloc(ctx, &m->vals, (size_t)(nnz + kNnzPad) * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(m->colind + nnz, 0, kNnzPad * sizeof(int32_t), ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->vals + nnz, 0, kNnzPad * sizeof(double), ctx->stream);
    if (e != cudaSuccess) {
        hg_set_error("matrix: device allocation of %lld nnz failed: %s", (long long)nnz,
                     cudaGetErrorString(e));
        hg_matrix_destroy(m);
        return HG_ERR_NOMEM;
    }
    hg_matrix_pick_tpr(m);
    *out = m;
    return HG_OK;
}

std::mutex& hg_matrix_form_mutex() {
    static std::mutex mu;
    return mu;
}

void hg_matrix_pick_tpr(hg_matrix* m) {
    const double mean = m->rows > 0 ? (double)m->nnz / (double)m->rows : 0.0;
    int tpr = 2;
    if (mean >= 48) tpr = 32;
    else if (mean >= 24) tpr = 16;
    else if (mean >= 12) tpr = 8;
    else if (mean >= 6) tpr = 4;
    static const int forced = [] {
        const char* e = getenv("HG_TPR");  // experiments: threads per row of the CSR kernel (2..32)
        const int v = e ? atoi(e) : 0;
        return (v == 2 || v == 4 || v == 8 || v == 16 || v == 32) ? v : 0;
    }();
    if (forced) tpr = forced;
    m->tpr = tpr;
}

extern "C" int hg_matrix_destroy(hg_matrix* m) {
    if (!m) return HG_OK;
    hg_dfree(m->rowptr);
    hg_dfree(m->colind);
    hg_dfree(m->vals);
    if (m->unit_row) cudaFree(m->unit_row);
    hg_dfree(m->sell_ptr);
    hg_dfree(m->sell_col);
    hg_dfree(m->sell_val);
    hg_idx16_free(m);
    hg_group_free(m);
    delete m;
    return HG_OK;
}

extern "C" int hg_matrix_info(const hg_matrix* m, int64_t* rows, int64_t* cols, int64_t* nnz) {
    HG_REQUIRE(m, "hg_matrix_info: NULL matrix");
    if (rows) *rows = m->rows;
    if (cols) *cols = m->cols;
    if (nnz) *nnz = m->nnz;
    return HG_OK;
}

static int upload_ptr(hg_ctx* ctx, const void* ptr, int bits, int64_t count, int64_t* d_out) {
    if (bits == 64) {
        HG_CUDA(cudaMemcpyAsync(d_out, ptr, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
        return HG_OK;
    }
    int32_t* tmp = nullptr;
    HG_CUDA(hg_dmalloc(ctx, &tmp, (size_t)count * 4));
    cudaError_t e = cudaMemcpyAsync(tmp, ptr, (size_t)count * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        hg_launch_scope scope(ctx, HG_K_SETUP, 12.0 * (double)count);
        widen_ptr_kernel<<<(unsigned)cdiv(count, kBlock), kBlock, 0, ctx->stream>>>(tmp, d_out, count);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    hg_dfree(tmp);  // also on the error paths
    if (e != cudaSuccess) {
        hg_set_error("matrix upload: %s", cudaGetErrorString(e));
        return HG_ERR_CUDA;
    }
    return HG_OK;
}

// A malformed input must come back as HG_ERR_INVALID, not as out-of-bounds reads in every SpMV (or
// out-of-bounds atomics in the transposition).  Runs on the uploaded arrays; synchronises.
static int validate_csr(hg_ctx* ctx, const hg_matrix* m, const char* who) {
    unsigned int* d_bad = reinterpret_cast<unsigned int*>(ctx->d_scalars + 48);
    unsigned int h_bad = 0;
    HG_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned int), ctx->stream));
    {
        const int64_t work = std::max<int64_t>(std::max(m->rows, m->nnz), 1);
        hg_launch_scope scope(ctx, HG_K_SETUP, 4.0 * (double)m->nnz + 8.0 * (double)m->rows);
        validate_csr_kernel<<<(unsigned)cdiv(work, kBlock), kBlock, 0, ctx->stream>>>(m->rows, m->cols, m->nnz,
                                                                                     m->rowptr, m->colind, d_bad);
    }
    HG_CUDA(cudaGetLastError());
    HG_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_bad) {
        hg_set_error("%s: malformed sparse input:%s%s%s", who,
                     (h_bad & 1u) ? " pointer array does not start at 0 / end at nnz;" : "",
                     (h_bad & 2u) ? " pointer array is not non-decreasing;" : "",
                     (h_bad & 4u) ? " an index is outside the matrix;" : "");
        return HG_ERR_INVALID;
    }
    return HG_OK;
}

extern "C" int hg_matrix_from_csr(hg_ctx* ctx, int64_t rows, int64_t cols, int64_t nnz,
                                  const void* rowptr, int ptr_bits, const int32_t* colind,
                                  const double* vals, hg_matrix** out) {
    HG_REQUIRE(ctx && out, "hg_matrix_from_csr: NULL argument");
    HG_REQUIRE(ptr_bits == 32 || ptr_bits == 64, "hg_matrix_from_csr: ptr_bits must be 32 or 64");
    HG_REQUIRE(rowptr && (nnz == 0 || (colind && vals)), "hg_matrix_from_csr: NULL array");
    HG_CUDA(cudaSetDevice(ctx->device));
    hg_matrix* m = nullptr;
    HG_TRY(hg_matrix_alloc(ctx, rows, cols, nnz, &m));
    int st = upload_ptr(ctx, rowptr, ptr_bits, rows + 1, m->rowptr);
    if (st == HG_OK && nnz > 0) {
        cudaError_t e = cudaMemcpyAsync(m->colind, colind, (size_t)nnz * 4, cudaMemcpyHostToDevice,
                                        ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(m->vals, vals, (size_t)nnz * 8, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            hg_set_error("hg_matrix_from_csr: upload failed: %s", cudaGetErrorString(e));
            st = HG_ERR_CUDA;
        }
    }
    if (st == HG_OK) st = validate_csr(ctx, m, "hg_matrix_from_csr");
    if (st != HG_OK) {
        hg_matrix_destroy(m);
        return st;
    }
    *out = m;
    return HG_OK;
}

extern "C" int hg_matrix_from_csc(hg_ctx* ctx, int64_t rows, int64_t cols, int64_t nnz,
                                  const void* jc, const void* ir, int idx_bits, const double* pr,
                                  hg_matrix** out) {
    HG_REQUIRE(ctx && out, "hg_matrix_from_csc: NULL argument");
    HG_REQUIRE(idx_bits == 32 || idx_bits == 64, "hg_matrix_from_csc: idx_bits must be 32 or 64");
    HG_REQUIRE(jc && (nnz == 0 || (ir && pr)), "hg_matrix_from_csc: NULL array");
    HG_REQUIRE(rows <= 2147483647LL, "hg_matrix_from_csc: more than 2^31-1 rows is not supported");
    HG_CUDA(cudaSetDevice(ctx->device));
    // CSC(A) arrays are CSR(A^T): upload them as a cols x rows matrix T, then transpose.
    hg_matrix* t = nullptr;
    if (idx_bits == 32) {
        HG_TRY(hg_matrix_from_csr(ctx, cols, rows, nnz, jc, 32, (const int32_t*)ir, pr, &t));
    } else {
        HG_TRY(hg_matrix_alloc(ctx, cols, rows, nnz, &t));
        int st = upload_ptr(ctx, jc, 64, cols + 1, t->rowptr);
        // narrow the 64-bit row indices in bounded chunks
        const int64_t chunk = (int64_t)1 << 24;
        int64_t* tmp = nullptr;
        if (st == HG_OK && nnz > 0) {
            if (hg_dmalloc(ctx, &tmp, (size_t)std::min(chunk, nnz) * 8) != cudaSuccess) {
                hg_set_error("hg_matrix_from_csc: staging allocation failed");
                st = HG_ERR_NOMEM;
            }
        }
        for (int64_t off = 0; st == HG_OK && off < nnz; off += chunk) {
            const int64_t cnt = std::min(chunk, nnz - off);
            cudaError_t e = cudaMemcpyAsync(tmp, (const int64_t*)ir + off, (size_t)cnt * 8,
                                            cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) {
                hg_launch_scope scope(ctx, HG_K_SETUP, 12.0 * (double)cnt);
                narrow_idx_kernel<<<(unsigned)cdiv(cnt, kBlock), kBlock, 0, ctx->stream>>>(
                    tmp, t->colind + off, cnt);
                e = cudaGetLastError();
            }
            if (e != cudaSuccess) {
                hg_set_error("hg_matrix_from_csc: upload failed: %s", cudaGetErrorString(e));
                st = HG_ERR_CUDA;
            }
        }
        if (st == HG_OK && nnz > 0) {
            cudaError_t e = cudaMemcpyAsync(t->vals, pr, (size_t)nnz * 8, cudaMemcpyHostToDevice,
                                            ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) {
                hg_set_error("hg_matrix_from_csc: upload failed: %s", cudaGetErrorString(e));
                st = HG_ERR_CUDA;
            }
        }
        if (tmp) hg_dfree(tmp);
        if (st == HG_OK) st = validate_csr(ctx, t, "hg_matrix_from_csc");
        if (st != HG_OK) {
            hg_matrix_destroy(t);
            return st;
        }
        hg_matrix_pick_tpr(t);
    }
    int st = hg_transpose_device(ctx, t, out);
    hg_matrix_destroy(t);
    return st;
}

extern "C" int hg_matrix_from_dense(hg_ctx* ctx, int64_t rows, int64_t cols, const double* a,
                                    int64_t lda, hg_matrix** out) {
    HG_REQUIRE(ctx && out && a, "hg_matrix_from_dense: NULL argument");
    HG_REQUIRE(lda >= rows, "hg_matrix_from_dense: lda < rows");
    HG_CUDA(cudaSetDevice(ctx->device));
    hg_matrix* m = nullptr;
    HG_TRY(hg_matrix_alloc(ctx, rows, cols, rows * cols, &m));
    double* d_a = nullptr;
    const size_t bytes = (size_t)std::max<int64_t>(lda * cols, 1) * 8;
    if (hg_dmalloc(ctx, &d_a, bytes) != cudaSuccess) {
        hg_matrix_destroy(m);
        hg_set_error("hg_matrix_from_dense: staging allocation failed");
        return HG_ERR_NOMEM;
    }
    cudaError_t e = cudaMemcpyAsync(d_a, a, (size_t)(lda * cols) * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        hg_launch_scope scope(ctx, HG_K_SETUP, 20.0 * (double)(rows * cols));
        const int64_t work = std::max(rows * cols, rows + 1);
        dense_to_csr_kernel<<<(unsigned)cdiv(work, kBlock), kBlock, 0, ctx->stream>>>(
            rows, cols, d_a, lda, m->rowptr, m->colind, m->vals);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    hg_dfree(d_a);
    if (e != cudaSuccess) {
        hg_set_error("hg_matrix_from_dense: %s", cudaGetErrorString(e));
        hg_matrix_destroy(m);
        return HG_ERR_CUDA;
    }
    *out = m;
    return HG_OK;
}

// Canonicalise a CSR matrix in place: (col, val) of every row sorted by col (ties by the bit
// pattern of val => deterministic).  Rows up to 16384 entries sort in shared memory.
static int hg_sort_rows_device(hg_ctx* ctx, hg_matrix* t) {
    const int64_t tr = t->rows;
    if (tr == 0 || t->nnz == 0) return HG_OK;
    int st = HG_OK;
    unsigned long long* d_max = nullptr;
    int32_t* sk = nullptr;
    double* sv = nullptr;
#define TR_CUDA(call)                                                                    \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess && st == HG_OK) {                                          \
            hg_set_error("sort rows: %s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__); \
            st = HG_ERR_CUDA;                                                            \
        }                                                                                \
    } while (0)
    TR_CUDA(hg_dmalloc(ctx, &d_max, sizeof(unsigned long long)));
    if (st == HG_OK) TR_CUDA(cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), ctx->stream));
    if (st == HG_OK) {
        {
            hg_launch_scope scope(ctx, HG_K_SETUP, 8.0 * (double)tr);
            max_row_len_kernel<<<(unsigned)cdiv(tr, kBlock), kBlock, 0, ctx->stream>>>(tr, t->rowptr, d_max);
            TR_CUDA(cudaGetLastError());
        }
        unsigned long long h_max = 0;
        TR_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost, ctx->stream));
        TR_CUDA(cudaStreamSynchronize(ctx->stream));
        if (st == HG_OK) {
            const int cap_max = 16384;  // 12 B per entry -> 192 KB of shared memory
            int cap = 256;
            while (cap < (int)std::min<unsigned long long>(h_max, cap_max)) cap <<= 1;
            const size_t smem = (size_t)cap * 12;
            if (smem > 48 * 1024)
                TR_CUDA(cudaFuncSetAttribute(sort_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int64_t grid = std::min<int64_t>(tr, (int64_t)ctx->sm_count * 64);
            if (st == HG_OK) {
                hg_launch_scope scope(ctx, HG_K_SETUP, 24.0 * (double)t->nnz);
                sort_rows_kernel<<<(unsigned)grid, kBlock, smem, ctx->stream>>>(tr, t->rowptr, t->colind, t->vals, cap);
                TR_CUDA(cudaGetLastError());
            }
            if (st == HG_OK && h_max > (unsigned long long)cap) {
                TR_CUDA(hg_dmalloc(ctx, &sk, (size_t)t->nnz * 4));
                TR_CUDA(hg_dmalloc(ctx, &sv, (size_t)t->nnz * 8));
                if (st == HG_OK) {
                    hg_launch_scope scope(ctx, HG_K_SETUP, 24.0 * (double)t->nnz);
                    sort_long_rows_kernel<<<(unsigned)grid, kBlock, 0, ctx->stream>>>(
                        tr, t->rowptr, t->colind, t->vals, cap, sk, sv);
                    TR_CUDA(cudaGetLastError());
                }
            }
        }
    }
    if (st == HG_OK) TR_CUDA(cudaStreamSynchronize(ctx->stream));
#undef TR_CUDA
    if (d_max) hg_dfree(d_max);
    if (sk) hg_dfree(sk);
    if (sv) hg_dfree(sv);
    return st;
}

int hg_transpose_device(hg_ctx* ctx, const hg_matrix* m, hg_matrix** out) {
    *out = nullptr;
    HG_REQUIRE(m->rows <= 2147483647LL, "transpose: more than 2^31-1 rows is not supported");
    hg_matrix* t = nullptr;
    HG_TRY(hg_matrix_alloc(ctx, m->cols, m->rows, m->nnz, &t));
    const int64_t tr = t->rows;
    unsigned int* cnt = nullptr;
    int st = HG_OK;
    std::vector<unsigned int> h_cnt;
    std::vector<int64_t> h_ptr;
#define TR_CUDA(call)                                                                    \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess && st == HG_OK) {                                          \
            hg_set_error("transpose: %s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__); \
            st = HG_ERR_CUDA;                                                            \
        }                                                                                \
    } while (0)
    TR_CUDA(hg_dmalloc(ctx, &cnt, (size_t)(tr + 1) * sizeof(unsigned int)));
    if (st == HG_OK) TR_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(tr + 1) * sizeof(unsigned int), ctx->stream));
    if (st == HG_OK && m->nnz > 0) {
        hg_launch_scope scope(ctx, HG_K_SETUP, 4.0 * (double)m->nnz);
        count_cols_kernel<<<(unsigned)cdiv(m->nnz, kBlock), kBlock, 0, ctx->stream>>>(m->colind, m->nnz, cnt);
        TR_CUDA(cudaGetLastError());
    }
    if (st == HG_OK) {
        // exclusive scan on the host (setup path; tr+1 integers)
        h_cnt.resize((size_t)tr + 1);
        h_ptr.resize((size_t)tr + 1);
        TR_CUDA(cudaMemcpyAsync(h_cnt.data(), cnt, (size_t)(tr + 1) * sizeof(unsigned int),
                                cudaMemcpyDeviceToHost, ctx->stream));
        TR_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (st == HG_OK) {
        int64_t acc = 0;
        for (int64_t i = 0; i < tr; ++i) {
            h_ptr[(size_t)i] = acc;
            acc += h_cnt[(size_t)i];
        }
        h_ptr[(size_t)tr] = acc;
        TR_CUDA(cudaMemcpyAsync(t->rowptr, h_ptr.data(), (size_t)(tr + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        TR_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(tr + 1) * sizeof(unsigned int), ctx->stream));
    }
    if (st == HG_OK && m->nnz > 0) {
        {
            hg_launch_scope scope(ctx, HG_K_SETUP, 24.0 * (double)m->nnz);
            scatter_kernel<<<(unsigned)cdiv(m->rows * 32, kBlock), kBlock, 0, ctx->stream>>>(
                m->rows, m->rowptr, m->colind, m->vals, t->rowptr, cnt, t->colind, t->vals);
            TR_CUDA(cudaGetLastError());
        }
        if (st == HG_OK) st = hg_sort_rows_device(ctx, t);
    }
    if (st == HG_OK) TR_CUDA(cudaStreamSynchronize(ctx->stream));
#undef TR_CUDA
    if (cnt) hg_dfree(cnt);
    if (st != HG_OK) {
        hg_matrix_destroy(t);
        return st;
    }
    hg_matrix_pick_tpr(t);
    *out = t;
    return HG_OK;
}

extern "C" int hg_matrix_transpose(hg_ctx* ctx, const hg_matrix* m, hg_matrix** out) {
    HG_REQUIRE(ctx && m && out, "hg_matrix_transpose: NULL argument");
    HG_CUDA(cudaSetDevice(ctx->device));
    return hg_transpose_device(ctx, m, out);
}

static bool invert_perm(const int32_t* perm, int64_t n, std::vector<int32_t>& inv) {
    inv.assign((size_t)n, -1);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t p = perm[i];
        if (p < 0 || p >= n || inv[(size_t)p] != -1) return false;
        inv[(size_t)p] = (int32_t)i;
    }
    return true;
}

extern "C" int hg_matrix_permute(hg_ctx* ctx, const hg_matrix* m, const int32_t* rowperm,
                                 const int32_t* colperm, int flags, hg_matrix** out) {
    HG_REQUIRE(ctx && m && out, "hg_matrix_permute: NULL argument");
    HG_REQUIRE(m->rows <= 2147483647LL, "hg_matrix_permute: more than 2^31-1 rows is not supported");
    HG_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    std::vector<int32_t> inv;
    if (rowperm) HG_REQUIRE(invert_perm(rowperm, m->rows, inv), "hg_matrix_permute: rowperm is not a permutation of 0..rows-1");
    if (colperm) HG_REQUIRE(invert_perm(colperm, m->cols, inv), "hg_matrix_permute: colperm is not a permutation of 0..cols-1");
    hg_matrix* t = nullptr;
    HG_TRY(hg_matrix_alloc(ctx, m->rows, m->cols, m->nnz, &t));
    int32_t *d_rp = nullptr, *d_ci = nullptr, *d_len = nullptr;
    int st = HG_OK;
    std::vector<int32_t> h_len((size_t)m->rows);
    std::vector<int64_t> h_ptr((size_t)m->rows + 1);
#define PM_CUDA(call)                                                                    \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess && st == HG_OK) {                                          \
            hg_set_error("permute: %s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__); \
            st = HG_ERR_CUDA;                                                            \
        }                                                                                \
    } while (0)
    if (rowperm) {
        PM_CUDA(hg_dmalloc(ctx, &d_rp, (size_t)std::max<int64_t>(m->rows, 1) * 4));
        if (st == HG_OK) PM_CUDA(cudaMemcpyAsync(d_rp, rowperm, (size_t)m->rows * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (colperm) {
        PM_CUDA(hg_dmalloc(ctx, &d_ci, (size_t)std::max<int64_t>(m->cols, 1) * 4));
        if (st == HG_OK) PM_CUDA(cudaMemcpyAsync(d_ci, inv.data(), (size_t)m->cols * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    PM_CUDA(hg_dmalloc(ctx, &d_len, (size_t)std::max<int64_t>(m->rows, 1) * 4));
    if (st == HG_OK && m->rows > 0) {
        {
            hg_launch_scope scope(ctx, HG_K_SETUP, 16.0 * (double)m->rows);
            perm_row_len_kernel<<<(unsigned)cdiv(m->rows, kBlock), kBlock, 0, ctx->stream>>>(m->rows, m->rowptr, d_rp, d_len);
            PM_CUDA(cudaGetLastError());
        }
        PM_CUDA(cudaMemcpyAsync(h_len.data(), d_len, (size_t)m->rows * 4, cudaMemcpyDeviceToHost, ctx->stream));
        PM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (st == HG_OK) {
        int64_t acc = 0;
        for (int64_t i = 0; i < m->rows; ++i) {
            h_ptr[(size_t)i] = acc;
            acc += h_len[(size_t)i];
        }
        h_ptr[(size_t)m->rows] = acc;
        PM_CUDA(cudaMemcpyAsync(t->rowptr, h_ptr.data(), (size_t)(m->rows + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (st == HG_OK && m->nnz > 0) {
        hg_launch_scope scope(ctx, HG_K_SETUP, 24.0 * (double)m->nnz);
        perm_copy_kernel<<<(unsigned)cdiv(m->rows * 32, kBlock), kBlock, 0, ctx->stream>>>(
            m->rows, m->rowptr, m->colind, m->vals, d_rp, d_ci, t->rowptr, t->colind, t->vals);
        PM_CUDA(cudaGetLastError());
    }
    // relabelled columns: restore the canonical order unless the caller only runs products with it
    if (st == HG_OK && colperm && !(flags & HG_PERMUTE_KEEP_ENTRY_ORDER)) st = hg_sort_rows_device(ctx, t);
    if (st == HG_OK) PM_CUDA(cudaStreamSynchronize(ctx->stream));
#undef PM_CUDA
    if (d_rp) hg_dfree(d_rp);
    if (d_ci) hg_dfree(d_ci);
    if (d_len) hg_dfree(d_len);
    if (st != HG_OK) {
        hg_matrix_destroy(t);
        return st;
    }
    hg_matrix_pick_tpr(t);
    *out = t;
    return HG_OK;
}

extern "C" int hg_matrix_spmv_form(hg_ctx* ctx, const hg_matrix* m, int* form) {
    HG_REQUIRE(ctx && m && form, "hg_matrix_spmv_form: NULL argument");
    HG_CUDA(cudaSetDevice(ctx->device));
    if (hg_spmv_stream_eligible(m)) *form = 2;
    else if (m->rows > 0 && (hg_spmv_mode() == 0 || hg_spmv_mode() == 3) && hg_sell_ready(ctx, m))
        *form = m->sell_col8 ? 1 | 32 : (m->sell_col16 ? 1 | 16 : 1);
    else if (m->rows > 0 && m->tpr == 32 && hg_spmv_mode() == 0 && hg_group_ready(ctx, m))
        *form = m->grp_d16 ? 3 | 16 : 3;
    else if (m->rows > 0 && m->tpr == 32 && hg_spmv_mode() == 0 && hg_idx16_csr_enabled() && hg_csr16_ready(ctx, m))
        *form = 0 | 16;
    else *form = 0;
    return HG_OK;
}

extern "C" int hg_matrix_download_csr(hg_ctx* ctx, const hg_matrix* m, int64_t* rowptr,
                                      int32_t* colind, double* vals) {
    HG_REQUIRE(ctx && m, "hg_matrix_download_csr: NULL argument");
    HG_CUDA(cudaSetDevice(ctx->device));
    if (rowptr)
        HG_CUDA(cudaMemcpyAsync(rowptr, m->rowptr, (size_t)(m->rows + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (colind && m->nnz)
        HG_CUDA(cudaMemcpyAsync(colind, m->colind, (size_t)m->nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (vals && m->nnz)
        HG_CUDA(cudaMemcpyAsync(vals, m->vals, (size_t)m->nnz * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    return HG_OK;
}
