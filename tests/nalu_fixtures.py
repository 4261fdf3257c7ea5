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


class SSTFixture(LowMachFixture):
    """SSTKernelHex8Mesh (unit_tests/kernels/UnitTestKernelUtils.h:362-405): the low-Mach fields, then
    tvisc = 0.3 and the trigonometric density / tke / sdr / f-one-blending fields of
    TrigFieldFunction (unit_tests/kernels/UnitTestKernelUtils.C:93-220, a = 0.3)."""

    def __init__(self, perturb=False):
        super().__init__(perturb=perturb)
        c = self.coords
        a = 0.3 * np.pi
        x, y, z = c[:, 0], c[:, 1], c[:, 2]
        n = self.mesh.n_nodes
        self.tvisc = np.full(n, 0.3)
        self.density = np.cos(a * x) * np.cos(a * y) * np.cos(a * z)
        self.tke = 2.0 + np.cos(a * x) * np.sin(a * y) * np.cos(a * z)
        self.sdr = 2.0 + np.cos(a * x) * np.sin(a * y) * np.sin(a * z)
        self.fOneBlend = np.sin(a * x) * np.cos(a * y) * np.sin(a * z)

    # SolutionOptions::initialize_turbulence_constants with turbulenceModel_ == LAMINAR (the tests never set it):
    # src/SolutionOptions.C:742-770
    TURB = dict(betaStar=0.09, tkeProdLimitRatio=500.0, sigmaWTwo=0.856, betaOne=0.075, betaTwo=0.0828,
                gammaOne=5.0 / 9.0, gammaTwo=0.44)
