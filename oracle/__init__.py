"""CPU oracle for the Hybrid-GMRES hot path — TEST INFRASTRUCTURE ONLY.

This package is a line-literal NumPy/SciPy FP64 restatement of the reference's
MATLAB solvers (``/root/reference/*.m``; file:line cited per function).  It is
the checker, never the product:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import it;
* nothing under ``hybrid_gmres_b200/`` imports it, and the product fails loudly
  when the CUDA library is missing.

PARITY UNPINNED.  The reference ships no tests, golden vectors, ``.mat``
fixtures or recorded outputs, and neither MATLAB nor
Octave exists in the build container, so the reference cannot be executed
here.  The oracle is therefore pinned only to (i) a literal reading of the
``.m`` sources, (ii) closed-form properties of the un-vendored generators
(``deriv2``/``shaw``/``heat``), and (iii) the relations asserted in the
reference's figure titles (``run_equivalence_plots.m:33,44,55,66``,
``run_ptr_rtp_comparison.m:29,39``), all checked in ``tests/test_oracle.py``.
``oracle/replay.m`` lets a MATLAB/Octave user replay the committed fixtures
through the untouched reference; it has not been exercised.
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
)
