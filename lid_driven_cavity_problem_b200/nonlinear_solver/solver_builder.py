"""Solver name -> bound ``solve(graph)`` (reference: nonlinear_solver/solver_builder.py:1-15).

``'cuda'`` is the device-resident Newton-Krylov solver of this package; ``'scipy'`` is the
small-case dense path.  ``'petsc'`` is the reference's own solver: it is not part of this
backend (PETSc is an external dependency of the reference) and is reported as such."""
CUDA_SOLVER = 'cuda'
SCIPY_SOLVER = 'scipy'
PETSC_SOLVER = 'petsc'


def build(solver_name, residual_f, **solver_options):
    if solver_name == CUDA_SOLVER:
        from lid_driven_cavity_problem_b200.nonlinear_solver import cuda_solver_wrapper
        wrapper = cuda_solver_wrapper.CudaSolverWrapper(residual_f, **solver_options)
    elif solver_name == SCIPY_SOLVER:
        from lid_driven_cavity_problem_b200.nonlinear_solver import scipy_solver_wrapper
        wrapper = scipy_solver_wrapper.ScipySolverWrapper(residual_f)
    elif solver_name == PETSC_SOLVER:
        raise RuntimeError("the 'petsc' solver lives in the reference package "
                           "(lid_driven_cavity_problem.nonlinear_solver.petsc_solver_wrapper); "
                           "this backend provides 'cuda' and 'scipy'")
    else:
        assert False, "Unknown solver named %s" % (solver_name,)
    solve = wrapper.solve
    return solve
