"""Pseudo-transient driver with the semantics of the reference's ``run_simulation``
(time_stepper.py:9-56): call ``solver(graph)`` once per step; a diverged step
halves dt and is retried; steady mode stops when max|U - U_old| / bc < 1e-6 and otherwise grows
dt by 1.2; fixed-time mode clamps the last step onto ``final_time``.

Differences, all on the safe side: PETSc is not imported; ``solver=None`` builds the cuda
solver (the reference's default constructs its PETSc wrapper without the required argument,
time_stepper.py:10-12); with ``adaptative_dt=False`` a diverging step raises instead of being
retried forever; when the solver returns a device-backed graph the steady-state norm comes from
the device reduction done during the solve, so no field crosses PCIe between steps.
"""
import logging

import numpy as np

from lid_driven_cavity_problem_b200.nonlinear_solver.exceptions import SolverDivergedException

logger = logging.getLogger(__name__)


def _steady_norm(graph):
    fn = getattr(graph, 'steady_state_norm', None)
    if fn is not None:
        return fn()
    return np.linalg.norm(graph.ns_x_mesh.phi - graph.ns_x_mesh.phi_old, ord=np.inf) / graph.bc


def run_simulation(graph, final_time, solver=None, minimum_dt=1e-6, adaptative_dt=True, max_steps=None,
                   on_step=None):
    if solver is None:
        from lid_driven_cavity_problem_b200.nonlinear_solver import solver_builder
        solver = solver_builder.build(solver_builder.CUDA_SOLVER, None)

    t = 0.0
    steps = 0
    while final_time is None or t < final_time:
        if final_time is not None and t + graph.dt > final_time:
            graph.dt = final_time - t
            logger.info("graph.dt would override final_time. new graph.dt=%s", graph.dt)

        logger.info("time: %s/%s", t, final_time)
        try:
            new_graph = solver(graph)
        except SolverDivergedException:
            logger.info('Simulation diverged.')
            if not adaptative_dt:
                raise RuntimeError("Newton solve diverged and adaptative_dt is off.")
            graph.dt /= 2.0
            logger.info("Will try with dt=%s", graph.dt)
            if graph.dt <= minimum_dt:
                raise RuntimeError("Timestep has reached a too low value. Giving up.")
            continue

        graph = new_graph
        t += graph.dt
        steps += 1
        if on_step is not None:
            on_step(graph, t)

        if final_time is None:
            norm = _steady_norm(graph)
            if norm < 1e-6:
                logger.info("Simulation reached Steady State at t=%s (norm=%s)", t, norm)
                break
            logger.info("norm=%s", norm)
            if adaptative_dt:
                graph.dt *= 1.2
        elif abs(final_time - t) <= 1e-5:
            logger.info("Simulation converged.")
            if adaptative_dt:
                graph.dt *= 2.0
                logger.info("Will update dt=%s", graph.dt)
        else:
            logger.info("Simulation converged. Finished at t=%s", t)
        if max_steps is not None and steps >= max_steps:
            break

    return graph
