"""Host-side partition logic of the multi-GPU path.

Rows of ``A`` (sinogram index, view-major) are split into ``P`` contiguous blocks with
balanced nonzero counts; ``B`` is split into the matching column blocks.  Krylov vectors
are split into equal row slices of ``n_p = roundup32(ceil(n/P))`` entries (NCCL's
reduce-scatter / all-gather need equal counts; the tail is zero padded).  Pure index
arithmetic: no floating-point work happens here.
"""
from __future__ import annotations

import numpy as np


def balanced_row_blocks(indptr, P: int):
    """Boundaries ``r[0..P]`` of ``P`` contiguous row blocks with ~equal nnz:
    ``r[p]`` is the first row whose start offset is >= ``p/P`` of the nonzeros."""
    indptr = np.asarray(indptr, dtype=np.int64)
    rows = indptr.shape[0] - 1
    nnz = int(indptr[-1])
    bounds = [0]
    for p in range(1, P):
        target = (nnz * p) // P
        r = int(np.searchsorted(indptr, target, side="left"))
        r = min(max(r, bounds[-1]), rows)
        bounds.append(r)
    bounds.append(rows)
    return np.array(bounds, dtype=np.int64)


def uniform_row_blocks(rows: int, P: int):
    """Equal-count contiguous blocks (CT: every view carries about the same nnz)."""
    return np.array([(rows * p) // P for p in range(P + 1)], dtype=np.int64)


def slice_len(n: int, P: int) -> int:
    """Per-rank slice length of an n-vector: ceil(n/P) rounded up to 32 doubles."""
    per = -(-n // P)
    return -(-per // 32) * 32


def shard_host_matrices(A, B, P: int, rank: int, bounds=None):
    """``(A_p, B^p, (r_lo, r_hi))`` for SciPy matrices: rows ``[r_lo,r_hi)`` of ``A`` and the
    matching columns of ``B`` (data preparation on the host; indices only)."""
    A = A.tocsr()
    if bounds is None:
        bounds = balanced_row_blocks(A.indptr, P)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    A_p = A[lo:hi, :].tocsr()
    B_p = B.tocsc()[:, lo:hi].tocsr()
    return A_p, B_p, (lo, hi)
