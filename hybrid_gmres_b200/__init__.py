"""hybrid_gmres_b200 — B200-native (sm_100a) Arnoldi / Golub-Kahan hot path of
luisayang-malaxiangguo/Hybrid-GMRES behind the reference's own solver signatures.

All compute is in ``libhgmres.so`` (hand-written CUDA, C ABI in
``include/hgmres.h``); this package is the host-side mirror of the reference's
MATLAB interface.  Importing it without the built library raises.
"""
from . import _lib

_lib.load()  # fail loudly if the CUDA library has not been built

from .api import (  # noqa: E402,F401
    ABgmres_hybrid_bounds,
    ABgmres_nonhybrid_bounds,
    BAgmres_hybrid_bounds,
    BAgmres_nonhybrid_bounds,
    Arnoldi,
    Context,
    DeviceMatrix,
    GcvProblem,
    KERNEL_CLASSES,
    clear_matrix_cache,
    default_context,
    fminbnd_gcv,
    gcv_function,
    gcv_prepare,
    hybrid_ab_gmres_rtp,
    hybrid_ba_gmres_rtp,
    hybrid_gmres_gcv,
    hybrid_lsmr_solver,
    hybrid_lsqr_solver,
    lsmr_solver,
    lsqr_solver,
    matrix_cache_info,
    set_option,
)
from .ct import ct_backprojector, ct_projector, ray_tables  # noqa: E402,F401
