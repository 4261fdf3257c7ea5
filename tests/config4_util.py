"""BASELINE config 4 (512^3 ABL-style box, SST turbulence): coordinates, nodal state and the kernel lists of the tke / sdr
equations, shared by the `-m gpu` tests and tools/bench_config4.py.  TEST / BENCH INFRASTRUCTURE (imports the oracle).

Kernel lists, in the order the equation systems register them on ONE AssembleElemSolverAlgorithm
(src/TurbKineticEnergyEquationSystem.C:447-494, src/SpecificDissipationRateEquationSystem.C:334-361):
  tke: [ScalarMassElemKernel(lumped), ScalarUpwAdvDiffElemKernel, TurbKineticEnergySSTSrcElemKernel]
  sdr: [ScalarMassElemKernel(lumped), ScalarUpwAdvDiffElemKernel, SpecificDissipationRateSSTSrcElemKernel]
Fields: the SST trig fixtures of unit_tests/kernels/UnitTestKernelUtils.C (tke / sdr / tvisc / fOneBlend test functions),
offset so that tke, sdr > 0 everywhere, evaluated at coordinates scaled to O(1) wavelengths per `wave` length units."""
import numpy as np

import oracle

BDF2 = dict(dt=0.05, gamma1=1.5, gamma2=-2.0, gamma3=0.5)
TURB = dict(betaStar=0.09, tkeProdLimitRatio=10.0, sigmaWTwo=0.856, betaOne=0.075, betaTwo=0.0828, gammaOne=5.0 / 9.0,
            gammaTwo=0.44)
TURB_ABI = dict(beta_star=0.09, tke_prod_limit_ratio=10.0, sigma_w_two=0.856, beta_one=0.075, beta_two=0.0828,
                gamma_one=5.0 / 9.0, gamma_two=0.44)
UPW = dict(alpha=0.0, alphaUpw=1.0, hoUpwind=1.0, hybridFactor=1.0)
UPW_ABI = dict(alpha=0.0, alpha_upw=1.0, ho_upwind=1.0, hybrid_factor=1.0)


def abl_coords(lattice, h_xy=0.4, h0=0.02, growth=1.25, k_uniform=12.0, h_top=0.3):
    """lattice (index-space, possibly perturbed) coordinates -> ABL-style physical coordinates: uniform spacing h_xy in x, y;
    cell heights growing geometrically from the wall (first cell h0, ratio `growth`) up to layer k_uniform, uniform above.
    A continuous map, so a perturbation of the lattice coordinate scales with the local cell height."""
    c = np.empty_like(lattice)
    c[:, 0] = h_xy * lattice[:, 0]
    c[:, 1] = h_xy * lattice[:, 1]
    k = lattice[:, 2]
    c[:, 2] = h0 * (growth ** np.minimum(k, k_uniform) - 1.0) / (growth - 1.0) + h_top * np.maximum(k - k_uniform, 0.0)
    return c


def state(c, eq, wave=1.0):
    """nodal fields of the tke (`eq`="tke") or sdr equation at physical coordinates c"""
    from nalu_b200 import fields
    a = 0.3 * np.pi / wave
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    vel = fields.velocity(c / wave) + np.stack([np.log1p(20 * z / wave), 0 * z, 0.02 * np.sin(a * x)], axis=1)
    rho = 1.0 + 0.1 * np.cos(a * x) * np.cos(a * y)
    tke = 2.0 + np.cos(a * x) * np.sin(a * y) * np.cos(a * z)
    sdr = 2.0 + np.cos(a * x) * np.sin(a * y) * np.sin(a * z)
    tvisc = 0.3 + 0.1 * np.sin(a * z)
    f1 = 0.5 + 0.4 * np.sin(a * x) * np.cos(a * y)
    q = tke if eq == "tke" else sdr
    gjq = np.stack([np.cos(a * c[:, j] + 0.1 * j) * (0.2 + 0.05 * j) / wave for j in range(3)], axis=1)
    visc_eff = 1.8e-5 + tvisc * (0.85 if eq == "tke" else 0.5)
    return dict(coordinates=c, velocity=vel, density=rho, density_n=0.98 * rho, density_nm1=0.95 * rho, tke=tke, sdr=sdr,
                tvisc=tvisc, f_one_blend=f1, scalarq=q, scalarq_n=0.97 * q, scalarq_nm1=0.93 * q, gjq=gjq, viscosity=visc_eff)


def oracle_list(m, f, eq, src_lumped=True, limiter=True):
    """the oracle's descriptors of the equation's kernel list on mesh m with nodal state f (coordinates may be a shifted
    copy: f["coordinates"]); returns (descs, keepalive)"""
    keep = []
    c, q, rho = f["coordinates"], f["scalarq"], f["density"]
    mdot = f.get("mdot")
    if mdot is None:
        mdot = oracle.calc_mass_flow_rate_scs(m.elem_nodes, c, rho, f["velocity"])
    src = "tke_sst_src" if eq == "tke" else "sdr_sst_src"
    descs = [
        oracle.make_desc("scalar_mass", keep, coordinates=c, scalarQ=q, scalarQN=f["scalarq_n"], scalarQNm1=f["scalarq_nm1"],
                         density=rho, densityN=f["density_n"], densityNm1=f["density_nm1"], lumpedMass=1, **BDF2),
        oracle.make_desc("scalar_upw_adv_diff", keep, coordinates=c, velocity=f["velocity"], Gjq=f["gjq"], scalarQ=q, density=rho,
                         viscosity=f["viscosity"], massFlowRate=mdot, useLimiter=int(limiter), **UPW),
        oracle.make_desc(src, keep, coordinates=c, velocity=f["velocity"], density=rho, tke=f["tke"], sdr=f["sdr"],
                         tvisc=f["tvisc"], fOneBlend=f["f_one_blend"], lumpedMass=int(src_lumped), **TURB)]
    return descs, keep  # the descriptors point into numpy arrays held by `keep`


def product_list(ls, d, eq, mdot, src_lumped=True, limiter=True):
    """the product's descriptors of the same list; d: {name: device tensor} with the equation's dof field shared between the
    kernels (d[eq] is d["scalarq"]: one field, one pointer, as in Nalu) for the fused tiled launch"""
    src = "tke_sst_src" if eq == "tke" else "sdr_sst_src"
    return [ls.make_desc("scalar_mass", fields=d, options=("lumped_mass",), **BDF2),
            ls.make_desc("scalar_upw_adv_diff", fields=d, options=("use_limiter",) if limiter else (), mass_flow_rate=mdot,
                         upw=UPW_ABI),
            ls.make_desc(src, fields=d, options=("lumped_mass",) if src_lumped else (), turb=TURB_ABI)]
