"""B200-native ``cuda`` backend for the coupled lid-driven-cavity solve.

Mirrors the plug points of tarcisiofischer/lid_driven_cavity_problem
(``residual_function(X, graph)``, ``solver_builder.build(name, residual_f)``,
``time_stepper.run_simulation``, ``options``) on top of hand-written sm_100a CUDA
reached through the C-ABI declared in ``include/ldc_b200.h``.
"""
__version__ = "0.1.0"
