"""ctypes binding of ``libhgmres.so`` — the C ABI declared in ``include/hgmres.h``.

This is the Python stand-in for the MEX gateway: it only
unpacks array pointers and sizes.  There is no fallback of any kind: if the
shared library is missing the import raises, and if no B200 is present
``hg_ctx_create`` fails with ``HG_ERR_CUDA``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhgmres.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int64_p = C.POINTER(C.c_int64)
c_void_pp = C.POINTER(C.c_void_p)


class HgExtras(C.Structure):
    _fields_ = [("H", c_double_p), ("beta", c_double_p), ("X_hist", c_double_p), ("aux", c_double_p)]


class HgSolverOpts(C.Structure):
    _fields_ = [("residual_mode", C.c_int), ("error_mode", C.c_int), ("reserved", C.c_int * 6)]


class HgError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libhgmres status {status}: {message}")
        self.status = status


# name -> (restype, argtypes); every symbol include/hgmres.h declares
_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
PROTOTYPES = {
    "hg_last_error": (C.c_char_p, []),
    "hg_version": (_i, []),
    "hg_set_option": (_i, [C.c_char_p, _i]),
    "hg_ctx_create": (_i, [_i, _vp, c_void_pp]),
    "hg_ctx_destroy": (_i, [_vp]),
    "hg_ctx_sync": (_i, [_vp]),
    "hg_ctx_trim": (_i, [_vp]),
    "hg_ctx_launch_count": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "hg_ctx_timing_enable": (_i, [_vp, _i]),
    "hg_ctx_timing_get": (_i, [_vp, _i, c_double_p, C.POINTER(C.c_uint64), c_double_p]),
    "hg_ctx_timing_reset": (_i, [_vp]),
    "hg_host_register": (_i, [_vp, C.c_size_t]),
    "hg_host_unregister": (_i, [_vp]),
    "hg_matrix_from_csr": (_i, [_vp, _i64, _i64, _i64, _vp, _i, _vp, _vp, c_void_pp]),
    "hg_matrix_from_csc": (_i, [_vp, _i64, _i64, _i64, _vp, _vp, _i, _vp, c_void_pp]),
    "hg_matrix_from_dense": (_i, [_vp, _i64, _i64, _vp, _i64, c_void_pp]),
    "hg_matrix_transpose": (_i, [_vp, _vp, c_void_pp]),
    "hg_matrix_permute": (_i, [_vp, _vp, _vp, _vp, _i, c_void_pp]),
    "hg_matrix_spmv_form": (_i, [_vp, _vp, c_int_p]),
    "hg_matrix_info": (_i, [_vp, c_int64_p, c_int64_p, c_int64_p]),
    "hg_matrix_download_csr": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "hg_matrix_destroy": (_i, [_vp]),
    "hg_ct_projector": (_i, [_vp, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp, c_void_pp]),
    "hg_ct_backprojector": (_i, [_vp, _i, _i, _i, _i, _d, _vp, _vp, c_void_pp]),
    "hg_ct_projector_rows": (_i, [_vp, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _i64, _i64, c_void_pp]),
    "hg_ct_backprojector_cols": (_i, [_vp, _i, _i, _i, _i, _d, _vp, _vp, _i64, _i64, c_void_pp]),
    "hg_comm_unique_id": (_i, [_vp]),
    "hg_comm_init": (_i, [_vp, _i, _i, _vp, c_void_pp]),
    "hg_comm_destroy": (_i, [_vp]),
    "hg_comm_transport": (_i, [_vp, c_int_p, C.c_char_p, _i]),
    "hg_darnoldi_create": (_i, [_vp, _vp, _vp, _vp, _i, c_void_pp]),
    "hg_darnoldi_destroy": (_i, [_vp]),
    "hg_darnoldi_set_rhs": (_i, [_vp, _vp]),
    "hg_darnoldi_reset": (_i, [_vp, _d]),
    "hg_darnoldi_steps": (_i, [_vp, _i]),
    "hg_darnoldi_get": (_i, [_vp, _vp, _i, c_double_p, c_int_p]),
    "hg_darnoldi_get_q": (_i, [_vp, _i, _vp, c_int64_p, c_int64_p]),
    "hg_darnoldi_step_bytes": (_i, [_vp, _i, c_double_p]),
    "hg_dist_gkb_solver": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _i, _d, _vp, _vp, _vp, _vp, c_int_p,
                                C.POINTER(HgExtras)]),
    "hg_dist_gcv_prepare": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, c_void_pp]),
    "hg_dist_hybrid_rtp": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _i, _d, _vp, _vp, _vp, c_int_p, c_int_p,
                                C.POINTER(HgExtras)]),
    "hg_spmv": (_i, [_vp, _vp, _vp, _vp]),
    "hg_multidot": (_i, [_vp, _i64, _i, _vp, _i64, _vp, _vp]),
    "hg_lincomb": (_i, [_vp, _i64, _i, _vp, _i64, _vp, _d, _vp, _vp, c_double_p]),
    "hg_cgs_mid": (_i, [_vp, _i64, _i, _vp, _i64, _vp, _vp, _i, _vp, _vp]),
    "hg_cgs2_step": (_i, [_vp, _i64, _i, _vp, _i64, _vp, _vp, _vp]),
    "hg_arnoldi_create": (_i, [_vp, _vp, _vp, _i, _i, c_void_pp]),
    "hg_arnoldi_destroy": (_i, [_vp]),
    "hg_arnoldi_set_rhs": (_i, [_vp, _vp]),
    "hg_arnoldi_reset": (_i, [_vp, _d]),
    "hg_arnoldi_steps": (_i, [_vp, _i]),
    "hg_arnoldi_get": (_i, [_vp, _vp, _i, c_double_p, c_int_p]),
    "hg_arnoldi_get_q": (_i, [_vp, _i, _vp]),
    "hg_arnoldi_step_bytes": (_i, [_vp, _i, c_double_p]),
    "hg_hybrid_ab_gmres_rtp": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _d, _vp, _vp, _vp, c_int_p, c_int_p,
                                    C.POINTER(HgSolverOpts), C.POINTER(HgExtras)]),
    "hg_hybrid_ba_gmres_rtp": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _d, _vp, _vp, _vp, c_int_p, c_int_p,
                                    C.POINTER(HgSolverOpts), C.POINTER(HgExtras)]),
    "hg_last_solve_stats": (_i, [_vp, _i]),
    "hg_gmres_ptr": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _d, _i, _d, _vp, _vp, _vp, c_int_p, c_int_p,
                          C.POINTER(HgExtras)]),
    "hg_gmres_ptr_gcv": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _d, _i, _vp, _i, _vp, _vp, _vp, _vp, c_int_p, c_int_p,
                              C.POINTER(HgExtras)]),
    "hg_gcv_prepare": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, c_void_pp]),
    "hg_gcv_eval": (_i, [_vp, _d, c_double_p]),
    "hg_gcv_get": (_i, [_vp, _vp, c_double_p]),
    "hg_gcv_fminbnd": (_i, [_vp, _d, _d, _d, c_double_p, c_double_p, c_int_p, _vp, _i]),
    "hg_gcv_destroy": (_i, [_vp]),
    "hg_gcv_surface": (_i, [_vp, _vp, _i, _vp, _vp]),
    "hg_gcv_from_H": (_i, [_vp, _i, _i, _d, _d, c_void_pp]),
    "hg_host_hessenberg_ls": (_i, [_vp, _i, _i, _d, _vp]),
    "hg_host_solve_square": (_i, [_i, _vp, _i, _vp, _vp]),
    "hg_host_singular_values": (_i, [_i, _vp, _i, _vp]),
    "hg_hybrid_lsqr_solver": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _d, _vp, _vp, _vp, c_int_p,
                                   C.POINTER(HgExtras)]),
    "hg_hybrid_lsmr_solver": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _d, _vp, _vp, _vp, c_int_p,
                                   C.POINTER(HgExtras)]),
    "hg_lsqr_solver": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _vp, _vp, _vp, c_int_p, C.POINTER(HgExtras)]),
    "hg_lsmr_solver": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _vp, _vp, _vp, _vp, c_int_p,
                            C.POINTER(HgExtras)]),
}

_lib = None


def load() -> C.CDLL:
    """Load the library (once).  Raises if it has not been built — the product
    path must fail loudly rather than fall back (task rule ③)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C hybrid_gmres_b200/csrc`. hybrid_gmres_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().hg_last_error()
        raise HgError(status, msg.decode("utf-8", "replace") if msg else "")
