"""Python restatement of the reference's kernel test fixtures
(unit_tests/kernels/UnitTestKernelUtils.h:217-551): one Hex8 element (`generated:1x1x1`),
analytic fields.  Shared by the oracle tests (CPU) and the CUDA parity tests (GPU)."""
import numpy as np

from nalu_b200 import fields, mesh


class LowMachFixture:
    """LowMachKernelHex8Mesh / MomentumKernelHex8Mesh (UnitTestKernelUtils.h:263-352).
    Multi-state fields are declared with 2 states and only StateNP1 is filled, so
    the N state is zero-initialised."""

    def __init__(self, perturb=False, nx=1, ny=1, nz=1):
        self.mesh = mesh.generated_box(nx, ny, nz)
        if perturb:
            mesh.perturb_coords_lcg(self.mesh, 0.125)
        c = self.mesh.coords
        n = self.mesh.n_nodes
        self.coords = c
        self.velocity = fields.velocity(c)
        self.pressure = fields.pressure(c)
        self.dpdx = fields.dpdx(c)
        self.density = np.ones(n)
        self.velocityN = np.zeros((n, 3))
        self.densityN = np.zeros(n)
        self.viscosity = np.full(n, 0.1)
        self.mdot = None

    def calc_mdot(self, oracle):
        self.mdot = oracle.calc_mass_flow_rate_scs(self.mesh.elem_nodes, self.coords, self.density, self.velocity)
        return self.mdot


class MixtureFractionFixture:
    """MixtureFractionKernelHex8Mesh (UnitTestKernelUtils.h:486-551), incl. the swapped
    (amf_, znot_) call at :526 -> znot=2, amf=1."""

    def __init__(self, perturb=True):
        self.mesh = mesh.generated_box(1, 1, 1)
        if perturb:
            mesh.perturb_coords_lcg(self.mesh, 0.125)
        c = self.mesh.coords
        self.coords = c
        znot_, amf_ = 1.0, 2.0
        self.mixFraction = fields.mixture_fraction(c, amf_, znot_)
        self.velocity = fields.velocity(c)
        self.density = fields.inverse_property_from_mixture_fraction(self.mixFraction, 0.163, 1.18)
        self.viscosity = fields.property_from_mixture_fraction(self.mixFraction, 1.967e-5, 1.85e-5)
        self.mdot = None

    def calc_mdot(self, oracle):
        self.mdot = oracle.calc_mass_flow_rate_scs(self.mesh.elem_nodes, self.coords, self.density, self.velocity)
        return self.mdot
