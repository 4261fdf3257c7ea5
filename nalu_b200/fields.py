"""Analytic nodal test fields (vectorised) used by the tests and the synthetic benchmark.

They follow the reference's unit-test field functions so that the fixtures reproduce the
reference's golden vectors: unit_tests/kernels/UnitTestKernelUtils.C:23-274 (steady Taylor
vortex `TrigFieldFunction`, a=0.3) and :487-536 (mixture-fraction properties).
"""
import numpy as np

PI = np.arccos(-1.0)
A = 0.3


def velocity(coords):
    x, y = coords[:, 0], coords[:, 1]
    u = np.zeros_like(coords)
    u[:, 0] = -np.cos(A * PI * x) * np.sin(A * PI * y)
    u[:, 1] = +np.sin(A * PI * x) * np.cos(A * PI * y)
    return u


def pressure(coords):
    x, y = coords[:, 0], coords[:, 1]
    return -1.0 / 4.0 * (np.cos(2.0 * A * PI * x) + np.cos(2.0 * A * PI * y))


def dpdx(coords):
    x, y = coords[:, 0], coords[:, 1]
    g = np.zeros_like(coords)
    g[:, 0] = 0.5 * A * PI * np.sin(2.0 * A * PI * x)
    g[:, 1] = 0.5 * A * PI * np.sin(2.0 * A * PI * y)
    return g


def dudx(coords):
    """TrigFieldFunction::dudx (unit_tests/kernels/UnitTestKernelUtils.C:38-63): [n][3*i+j] = du_i/dx_j of `velocity`"""
    x, y = coords[:, 0], coords[:, 1]
    a_pi = A * PI
    cosx, sinx, cosy, siny = np.cos(a_pi * x), np.sin(a_pi * x), np.cos(a_pi * y), np.sin(a_pi * y)
    g = np.zeros((coords.shape[0], 9))
    g[:, 0] = a_pi * sinx * siny
    g[:, 1] = -a_pi * cosx * cosy
    g[:, 3] = a_pi * cosx * cosy
    g[:, 4] = -a_pi * sinx * siny
    return g


def density_trig(coords, rhonot=1.0):
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    return rhonot * (np.cos(A * PI * x) * np.cos(A * PI * y) * np.cos(A * PI * z))


def mixture_fraction(coords, znot, amf):
    """mixture_fraction_test_function(bulk, coords, mixFrac, znot, amf)
    (unit_tests/kernels/UnitTestKernelUtils.C:519-536).  NOTE the fixture calls it with
    (amf_, znot_) swapped (UnitTestKernelUtils.h:526), i.e. znot=2, amf=1."""
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    return znot * np.cos(amf * PI * x) * np.cos(amf * PI * y) * np.cos(amf * PI * z)


def property_from_mixture_fraction(z, primary, secondary):
    return primary * z + secondary * (1.0 - z)


def inverse_property_from_mixture_fraction(z, primary, secondary):
    return 1.0 / (z / primary + (1.0 - z) / secondary)


def smooth_viscosity(coords, mu0=0.1):
    """Strictly positive, smoothly varying viscosity for the synthetic boxes."""
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    return mu0 * (1.0 + 0.25 * np.sin(0.11 * x) * np.cos(0.07 * y) * np.cos(0.05 * z))


def smooth_density(coords, rho0=1.0):
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    return rho0 * (1.0 + 0.2 * np.cos(0.09 * x) * np.sin(0.06 * y) * np.cos(0.04 * z))
