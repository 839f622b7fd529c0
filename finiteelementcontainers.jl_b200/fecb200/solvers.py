"""Callers of the hot path (src/Solvers.jl, src/integrators/QuasiStaticIntegrator.jl), run
device-resident inside libfecb200."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FECError, check, lib
from .assemblers import update_bc_values, update_time


class IterativeLinearSolver:
    """IterativeLinearSolver(asm, :cg) (src/Solvers.jl:92-153).  Krylov.jl CG defaults."""

    def __init__(self, assembler, solver_sym="cg", matrix_free=None):
        if str(solver_sym).lower() not in ("cg", "cgsolver"):
            raise ValueError("only CG is implemented on the device")
        self.assembler = assembler
        self.matrix_free = assembler.matrix_free if matrix_free is None else bool(matrix_free)
        self.cg_iterations = 0

    def solve(self, b, x=None, atol=-1.0, rtol=-1.0, itmax=0):
        asm = self.assembler
        x = np.empty(asm.sizes()[2]) if x is None else x
        its, rn = C.c_int64(), C.c_double()
        check(lib.fecb200_cg_solve(asm._require(), _lib.ptr(b), _lib.ptr(x), atol, rtol, itmax,
                                   int(self.matrix_free), C.byref(its), C.byref(rn)))
        self.cg_iterations = its.value
        return x, its.value, rn.value


class DirectLinearSolver:
    """DirectLinearSolver(asm) + solve!(solver, Uu, p) (src/Solvers.jl:37-86): residual (+ source + Neumann loads) and
    stiffness assembled on the device, the sparse direct solve of K x = R on the host -- exactly where the reference does it ("currently doesn't
    work on GPU", Solvers.jl:81; incl. the Robin vector / matrix terms, Solvers.jl:73-76).  Uu is updated in place."""

    def __init__(self, assembler):
        if assembler.matrix_free:
            raise FECError("DirectLinearSolver needs an assembled matrix (matrix_free=false)")
        self.assembler = assembler
        self.matrix_free = False
        self.dUu = None

    def solve(self, Uu, p):
        import scipy.sparse.linalg as spla
        from . import assemblers as A
        from .physics import residual, stiffness
        asm = self.assembler
        A.assemble_vector(asm, residual, Uu, p)
        A.assemble_vector_source(asm, Uu, p)
        A.assemble_vector_neumann_bc(asm, Uu, p)
        A.assemble_stiffness(asm, stiffness, Uu, p)
        if getattr(p, "robin_bcs", None) is not None and len(p.robin_bcs):     # Solvers.jl:73-76
            A.assemble_vector_robin_bc(asm, Uu, p)
            A.assemble_matrix_robin_bc(asm, Uu, p)
        R = A._residual_accessor(asm)
        K = A._stiffness_accessor(asm)
        self.dUu = -spla.spsolve(K.tocsc(), R)
        Uu += self.dUu
        return Uu


class NewtonSolver:
    """NewtonSolver(linear_solver) (src/Solvers.jl:178-220): <= 10 iterations, tolerances 1e-12."""

    def __init__(self, linear_solver, max_iters=10, tol=1e-12):
        self.linear_solver = linear_solver
        self.max_iters, self.tol = max_iters, tol
        self.iterations = 0
        self.cg_iterations = 0
        self.residual_norm = 0.0

    def solve(self, Uu, p):
        asm = self.linear_solver.assembler
        if isinstance(self.linear_solver, DirectLinearSolver):
            # solve!(::NewtonSolver) (src/Solvers.jl:193-220) around the host-side direct solve
            from . import assemblers as A
            r0 = 0.0
            for n in range(1, self.max_iters + 1):
                self.linear_solver.solve(Uu, p)
                n_du = float(np.linalg.norm(self.linear_solver.dUu))
                n_r = float(np.linalg.norm(A._residual_accessor(asm)))
                if n == 1:
                    r0 = n_r
                rel = n_r / r0 if r0 > 0.0 else n_r
                self.iterations, self.residual_norm = n, n_r
                if n_du < self.tol or n_r < self.tol or rel < self.tol:
                    break
            self.cg_iterations = 0
            return Uu
        nit, cgit, rn = C.c_int32(), C.c_int64(), C.c_double()
        check(lib.fecb200_newton_solve(asm._require(), _lib.ptr(Uu), self.max_iters, self.tol,
                                       int(self.linear_solver.matrix_free), C.byref(nit), C.byref(cgit), C.byref(rn)))
        self.iterations, self.cg_iterations, self.residual_norm = nit.value, cgit.value, rn.value
        return Uu


class QuasiStaticIntegrator:
    """QuasiStaticIntegrator(solver) + evolve! (src/integrators/QuasiStaticIntegrator.jl:16-34)"""

    def __init__(self, solver):
        self.solver = solver
        self.solution = None
        self.failed = False

    def evolve(self, p, commit_state=True):
        """evolve!(integrator, p) (QuasiStaticIntegrator.jl:16-34).  After the solve the reference pushes the converged
        solution into p.field (`_update_for_assembly!`, :21-24); so does this.  The reference never advances
        state_old; with `commit_state` (default, needed by any stateful law) the state is re-evaluated AT the converged
        solution by one residual pass and then committed (state_old <- state_new); pass False for the reference's
        behaviour."""
        from .assemblers import assemble_vector, update_field
        from .physics import residual
        asm = self.solver.linear_solver.assembler
        if self.solution is None:
            self.solution = np.zeros(asm.sizes()[2])
        update_time(p)
        update_bc_values(p)
        self.solver.solve(self.solution, p)
        stateful = any(ph.NS > 0 for ph in p.physics)
        if commit_state and stateful:
            assemble_vector(asm, residual, self.solution, p)   # field + state_new at the converged solution
            check(lib.fecb200_state_swap(asm._require()))
        else:
            update_field(p, self.solution)
        # the reference flags a step whose last Newton increment AND residual are both above tolerance (:27-31);
        # the device loop only leaves early when one of its three tests passed, so "ran out of iterations" is that case
        self.failed = bool(self.solver.iterations >= self.solver.max_iters and not self.solver.residual_norm < self.solver.tol)
        return self.solution
