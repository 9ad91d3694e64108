class SolverDivergedException(RuntimeError):
    """Raised by a solver wrapper when the Newton solve of one pseudo-time step diverged
    (reference: nonlinear_solver/exceptions.py:1-2; caught by time_stepper.run_simulation)."""
