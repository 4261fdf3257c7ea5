"""Shared helpers for the CUDA parity tests: one set of numpy inputs feeds both the oracle
(CPU restatement of the reference) and the product (CUDA through the C ABI)."""
import numpy as np

import oracle
from nalu_b200 import fields, mesh


class State:
    """Nodal / element fields of one rank mesh (numpy, float64)."""

    def __init__(self, m, seed=0, smooth=True):
        c = m.coords
        n = m.n_nodes
        rng = np.random.default_rng(seed)
        self.mesh = m
        self.coordinates = c
        self.velocity = fields.velocity(c) + 0.05 * np.sin(0.7 * c[:, [2, 0, 1]])
        self.velocity_n = 0.9 * self.velocity + 0.01
        self.velocity_nm1 = 0.8 * self.velocity - 0.02
        self.pressure = fields.pressure(c)
        self.dpdx = fields.dpdx(c) + 0.01
        self.density = fields.smooth_density(c)
        self.density_n = 0.98 * self.density
        self.density_nm1 = 0.95 * self.density
        self.viscosity = fields.smooth_viscosity(c)
        self.scalarq = fields.mixture_fraction(c, 1.0, 0.37)
        self.scalarq_n = 0.97 * self.scalarq
        self.scalarq_nm1 = 0.93 * self.scalarq
        self.dual_volume = np.full(n, 1.0) + 0.1 * np.cos(c[:, 0])
        # projected nodal gradients of the upwind kernels (any smooth field does: they are inputs)
        a = 0.3 * np.pi
        self.gjq = np.stack([np.cos(a * c[:, j] + 0.1 * j) * (0.2 + 0.05 * j) for j in range(3)], axis=1)
        self.gju = np.stack([np.sin(a * c[:, j % 3] + 0.2 * j) * (0.3 + 0.02 * j) for j in range(9)], axis=1)
        self.mdot = oracle.calc_mass_flow_rate_scs(m.elem_nodes, c, self.density, self.velocity)
        del rng

    def names(self):
        return ["coordinates", "velocity", "velocity_n", "velocity_nm1", "viscosity", "density", "density_n",
                "density_nm1", "pressure", "dpdx", "scalarq", "scalarq_n", "scalarq_nm1", "dual_volume", "gjq", "gju"]

    def to_device(self, device="cuda"):
        import torch
        d = {k: torch.as_tensor(np.ascontiguousarray(getattr(self, k)), device=device) for k in self.names()}
        d["mdot"] = torch.as_tensor(np.ascontiguousarray(self.mdot), device=device)
        return d


# kernel spec: (kind, options tuple, scalars dict)
_ORC_FIELD = dict(coordinates="coordinates", velocity="velocity", velocity_n="velocityN", velocity_nm1="velocityNm1",
                  viscosity="viscosity", density="density", density_n="densityN", density_nm1="densityNm1",
                  pressure="pressure", dpdx="Gpdx", scalarq="scalarQ", scalarq_n="scalarQN",
                  scalarq_nm1="scalarQNm1", gjq="Gjq", gju="Gju")
_ORC_OPT = dict(shifted_grad_op="shiftedGradOp", skew_symmetric="skewSymmetric", lumped_mass="lumpedMass",
                shift_mdot="shiftMdot", shift_poisson="shiftPoisson", reduced_sens="reducedSensitivities",
                use_limiter="useLimiter", peclet_tanh="pecletTanh")
_ORC_UPW = dict(alpha="alpha", alpha_upw="alphaUpw", ho_upwind="hoUpwind", hybrid_factor="hybridFactor",
                tanh_trans="tanhTrans", tanh_width="tanhWidth")
_ORC_SCALAR = dict(include_divu="includeDivU", dens_fac="densFac", om_dens_fac="om_densFac", dt="dt",
                   gamma1="gamma1", gamma2="gamma2", gamma3="gamma3")


def oracle_desc(st, kind, options=(), keep=None, fixed_block=None, **scalars):
    keep = keep if keep is not None else []
    kw = {o: getattr(st, p) for p, o in _ORC_FIELD.items()}
    kw["massFlowRate"] = st.mdot
    for o in options:
        kw[_ORC_OPT[o]] = 1
    for k, v in scalars.items():
        if k == "upw":
            for kk, vv in v.items():
                kw[_ORC_UPW[kk]] = vv
        else:
            kw[_ORC_SCALAR[k]] = v
    if fixed_block is not None:
        kw["fixedBlock"] = fixed_block
    d = oracle.make_desc(kind, keep, **kw)
    d._keep = keep
    return d


def product_desc(ls, dev, kind, options=(), fixed_block=None, **scalars):
    f = {k: v for k, v in dev.items() if k != "mdot"}
    return ls.make_desc(kind, fields=f, options=options, mass_flow_rate=dev["mdot"], fixed_block=fixed_block,
                        **scalars)


def rel_err(a, b, floor_scale):
    """max |a-b| / max(|a|,|b|, floor) elementwise; floor = eps-level of the row/matrix scale so that
    exact zeros produced by cancellation do not blow the ratio up."""
    a = np.asarray(a)
    b = np.asarray(b)
    denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor_scale)
    return float(np.max(np.abs(a - b) / denom)) if a.size else 0.0


def rel_err_rows(a, b, rowptr, floor_frac):
    """max over entries of |a-b| / max(|a|, |b|, floor_frac * max|row of b|): the per-row floor of SURVEY 8(d) -- an
    entry that cancels to (near) zero is compared against the scale of ITS OWN row, not of the whole matrix."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.size == 0:
        return 0.0
    rp = np.asarray(rowptr, dtype=np.int64)
    lens = np.diff(rp)
    rowmax = np.zeros(lens.size)
    nz = lens > 0
    rowmax[nz] = np.maximum.reduceat(np.abs(b), rp[:-1][nz])
    floor = np.repeat(floor_frac * rowmax, lens)
    denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), np.maximum(floor, 1e-300))
    return float(np.max(np.abs(a - b) / denom))


def perturbed_box(nx, ny, nz, amp=0.12):
    m = mesh.generated_box(nx, ny, nz)
    mesh.perturb_coords_hash(m, amp)
    return m
