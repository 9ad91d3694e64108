"""Global switches of the solve, same three names as the reference (options.py:1-3).

Unlike the reference wrappers, which bind the values at import time
(petsc_solver_wrapper.py:9), the cuda backend reads them at call time, so
``options.IGNORE_DIVERGED = True`` takes effect for the next ``solve``.
"""

PLOT_JACOBIAN = False        # debug only; the cuda backend dumps the sparsity pattern instead of plotting
SHOW_SOLVER_DETAILS = True   # log function-evaluation count and convergence reason per solve
IGNORE_DIVERGED = False      # do not raise SolverDivergedException on a diverged Newton solve
