"""Physics interface (src/Physics.jl:1-18) for the CUDA-side physics.

In the reference a physics is a Julia struct plus closures `residual/stiffness/...(physics, interps,
x_el, t, dt, u_el, u_el_old, state_old_q, state_new_q, props_el)`.  Closures cannot cross the C ABI,
so here a physics is a tag (`physics_id`) + property vector, and the element-level functions are the
tokens below, mapped BY IDENTITY to library kinds (an arbitrary user closure is rejected: there is no
CPU fallback).
"""
from __future__ import annotations

import numpy as np

from . import _lib


class AbstractPhysics:
    """AbstractPhysics{NF, NP, NS}"""
    NF = 0
    NP = 0
    NS = 0
    physics_id = 0

    def create_properties(self):
        return np.zeros(0)

    def create_initial_state(self):
        return np.zeros(self.NS)


class Poisson(AbstractPhysics):
    """Poisson(func) (test/poisson/TestPoissonCommon.jl:4-6), AbstractPhysics{1,0,0}.
    `func(X, t)`: X is (npts, ND); returns (npts,).  It is evaluated on the host at the quadrature
    points (as the reference does for Sources, src/bcs/Sources.jl:55-66) and uploaded."""
    NF, NP, NS = 1, 0, 0
    physics_id = _lib.PHYS_POISSON

    def __init__(self, func=None):
        self.func = func


class ThreeDimensional:
    ND = 3


class PlaneStrain:
    ND = 2


class _MechanicsBase(AbstractPhysics):
    def __init__(self, formulation=None):
        self.formulation = formulation if formulation is not None else ThreeDimensional()
        self.NF = self.formulation.ND


class Mechanics(_MechanicsBase):
    """Linear-elastic Mechanics(formulation) (test/mechanics/TestMechanicsCommon.jl:3-12),
    props = (rho, K, G) = (1e3, 10e9, 1e9)."""
    NP, NS = 3, 0
    physics_id = _lib.PHYS_LINEAR_ELASTIC

    def create_properties(self):
        return np.array([1e3, 10.0e9, 1.0e9])


class NeoHookean(_MechanicsBase):
    """Neo-Hookean Mechanics of test/mechanics/TestMechanicsLargeDeformation.jl:17-27,
    props = (rho, K, G) = (1e3, 10e6, 1e6).  variant='as_written' keeps the script's volumetric
    term verbatim."""
    NP, NS = 3, 0

    def __init__(self, formulation=None, variant="standard"):
        super().__init__(formulation)
        assert variant in ("standard", "as_written")
        self.physics_id = _lib.PHYS_NEOHOOKEAN if variant == "standard" else _lib.PHYS_NEOHOOKEAN_AS_WRITTEN

    def create_properties(self):
        return np.array([1e3, 10.0e6, 1.0e6])


class NonSymmetricTestPhysics(_MechanicsBase):
    """TEST law (not in the reference): linear elasticity + beta * delta_ij T_kl, a tangent without major symmetry, so
    that the transposed COO labelling of the reference's pattern is visible in a parity test.
    props = (rho, K, G, beta)."""
    NP, NS = 4, 0
    physics_id = _lib.PHYS_TEST_NONSYMMETRIC

    def create_properties(self):
        return np.array([1e3, 10.0e6, 1.0e6, 3.0e6])


class J2Plasticity(_MechanicsBase):
    """Stateful mechanics: AbstractPhysics{3,5,7} (hooks: test/mechanics_with_state/
    TestMechanicsWithState.jl:15-67); props = (rho, K, G, sigma_y, H), 7 states per qp."""
    NP, NS = 5, 7
    physics_id = _lib.PHYS_J2_PLASTICITY

    def create_properties(self):
        return np.array([1e3, 10.0e9, 1.0e9, 2.0e8, 1.0e8])


class _ElementFunction:
    """Token for one of the reference's element-level generic functions.  Calling it with an
    assembler dispatches to the accessor of the same name (Julia has both methods on one
    generic function: `residual(physics, interps, ...)` and `residual(asm)`)."""

    def __init__(self, name, kind, accessor=None, inplace=False, action=False):
        self.name, self.kind, self._accessor, self.inplace, self.action = name, kind, accessor, inplace, action

    def __call__(self, asm, *args):
        if self._accessor is None:
            raise TypeError(f"{self.name} has no assembler accessor")
        from . import assemblers
        return getattr(assemblers, self._accessor)(asm, *args)

    def __repr__(self):
        return f"<fecb200 element function {self.name}>"


residual = _ElementFunction("residual", _lib.RESIDUAL, "_residual_accessor")
residual_b = _ElementFunction("residual!", _lib.RESIDUAL, inplace=True)
stiffness = _ElementFunction("stiffness", _lib.STIFFNESS, "_stiffness_accessor")
stiffness_b = _ElementFunction("stiffness!", _lib.STIFFNESS, inplace=True)
mass = _ElementFunction("mass", _lib.MASS, "_mass_accessor")
mass_b = _ElementFunction("mass!", _lib.MASS, inplace=True)
stiffness_action = _ElementFunction("stiffness_action", _lib.STIFFNESS, action=True)
stiffness_action_b = _ElementFunction("stiffness_action!", _lib.STIFFNESS, inplace=True, action=True)
mass_action = _ElementFunction("mass_action", _lib.MASS, action=True)
mass_action_b = _ElementFunction("mass_action!", _lib.MASS, inplace=True, action=True)
lumped_mass = _ElementFunction("lumped_mass", _lib.LUMPED_MASS, "_vector_values_accessor")
energy = _ElementFunction("energy", _lib.ENERGY)


def kind_of(func, allowed):
    if not isinstance(func, _ElementFunction):
        raise TypeError("fecb200 assembles only the shipped element functions (residual, stiffness, mass, "
                        "stiffness_action, ...); arbitrary closures cannot run on the device and there is no CPU fallback")
    if func.kind not in allowed:
        raise ValueError(f"{func.name} is not valid here")
    return func.kind
