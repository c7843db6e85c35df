"""CPU oracle for the Hybrid-GMRES hot path — TEST INFRASTRUCTURE ONLY.

This package is a line-literal NumPy/SciPy FP64 restatement of the reference's
MATLAB solvers (``/root/reference/*.m``; file:line cited per function).  It is
the checker, never the product:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import it;
* nothing under ``hybrid_gmres_b200/`` imports it, and the product fails loudly
  when the CUDA library is missing.

PARITY PIN (round 2): the reference ships no tests, golden vectors or recorded outputs
 and neither MATLAB nor Octave exists in the build container.  The pin is
therefore built from the reference's own source text: ``oracle/mlab.py`` is an interpreter for the
MATLAB subset the reference uses; ``tests/golden/make_reference_golden.py`` EXECUTES the untouched
``/root/reference/*.m`` files with it (12 of the 22 files: the seven §8a solvers, ``gcv_function``,
the four ``*_bounds`` PTR solvers, ``generate_test_problem``, the ``fminbnd`` call of
``analyze_regularization.m:35-46`` and ``compute_gcv_surface`` of ``plot_gcv_surface.m``) and commits
the outputs as ``tests/golden/ref_*.npz``.  ``tests/test_reference_golden.py`` holds this hand
restatement to those outputs (same stopping iterations and history lengths, H / iterates /
histories <= 1e-10 on the CT cases, the `==0` breakdown epilogue exactly) and re-executes the
reference on every CPU run in the build container to prove the fixtures' provenance.
What remains unpinned, and is stated as such: MATLAB's BUILT-INS (``*``, ``norm``, ``mldivide``, ``svd``,
``eig``, ``fminbnd``) are NumPy/SciPy/LAPACK here, chosen after MATLAB's documented algorithms, so
rounding-level differences against MathWorks' kernels are possible; the un-vendored third-party
generators (``shaw``/``heat``/``deriv2``, ``PRtomo_mismatched``) are restated from their published
definitions; MATLAB's ``rng(0)`` ``randn`` stream is not reproducible.  ``oracle/replay.m`` lets a
MATLAB/Octave user replay the committed fixtures through the real interpreter; not exercised.
"""

from .generators import generate_test_problem, deriv2, shaw, heat  # noqa: F401
from .solvers import (  # noqa: F401
    hybrid_ab_gmres_rtp,
    hybrid_ba_gmres_rtp,
    hybrid_lsqr_solver,
    hybrid_lsmr_solver,
    lsqr_solver,
    lsmr_solver,
    gcv_function,
    arnoldi,
)
from .fminbnd import fminbnd  # noqa: F401
from .ptr import (  # noqa: F401,E402
    ABgmres_hybrid_bounds,
    ABgmres_nonhybrid_bounds,
    BAgmres_hybrid_bounds,
    BAgmres_nonhybrid_bounds,
    hybrid_gmres_gcv,
)
