"""Surface defects -- ModuleDefects.Zernike of the reference (ART/ModuleDefects.py:149-181).

Per ray the polynomials are evaluated inside the CUDA trace kernel (csrc/art_device.cuh
zernike_eval) from a shared-memory table.  The host methods below evaluate ONE point (chief-ray
alignment, user inspection) with the same factorisation the kernel uses:
    Z_(n,m) = R_n^l(rho) * cos(l theta)  for 2m >= n,   R_n^l(rho) * sin(l theta)  for 2m < n,   l = |2m - n|
(unnormalised; ART/recursive_zernike_generator.py's index convention: (1,0) = y, (1,1) = x).
The gridded defects of the reference (Fourrier, MeasuredMap) are not part of this path.
"""
from __future__ import annotations

import math

import numpy as np


def _radial(n, l, rho2):
    """R_n^l(rho) / rho^l as a polynomial in rho^2 and its derivative with respect to rho^2."""
    k = (n - l) // 2
    val = der = 0.0
    for s in range(k + 1):
        c = (-1) ** s * math.factorial(n - s) / (math.factorial(s) * math.factorial((n + l) // 2 - s)
                                                 * math.factorial((n - l) // 2 - s))
        p = k - s  # power of rho^2
        val += c * rho2**p
        if p > 0:
            der += c * p * rho2 ** (p - 1)
    return val, der


class Defect:
    pass


class Zernike(Defect):
    """Zernike(Support, coefficients): coefficients = {(n, m): c in mm}, 0 <= m <= n; normalised
    radius R = Support._CircumCirc()."""

    def __init__(self, Support, coefficients):
        self.coefficients = coefficients
        self.max_order = int(np.max([k[0] for k in coefficients]))
        self.support = Support
        self.R = Support._CircumCirc()

    def _eval(self, Point):
        x, y = float(Point[0]) / self.R, float(Point[1]) / self.R
        s = x * x + y * y
        w = complex(x, y)
        Z = dX = dY = 0.0
        for (n, m), c in self.coefficients.items():
            l = abs(2 * m - n)
            q, dq = _radial(n, l, s)
            ang = w**l
            dang = l * w ** (l - 1) if l > 0 else 0j  # d/dx (x+iy)^l; d/dy is i times this
            if 2 * m >= n:
                a, ax, ay = ang.real, dang.real, -dang.imag
            else:
                a, ax, ay = ang.imag, dang.imag, dang.real
            Z += c * q * a
            dX += c * (dq * 2 * x * a + q * ax)
            dY += c * (dq * 2 * y * a + q * ay)
        return Z, dX / self.R, dY / self.R

    def get_offset(self, Point):
        """Height of the defect at Point (relative to the optic centre), mm."""
        return self._eval(Point)[0]

    def get_normal(self, Point):
        """Unnormalised defect normal (-dZ/dx, -dZ/dy, 1) at Point."""
        _, dx, dy = self._eval(Point)
        return np.array([-dx, -dy, 1.0])

    def RMS(self):
        return np.sqrt(np.sum([c**2 for c in self.coefficients.values()]))

    def PV(self):
        pass

    def _lower(self):
        """(radius, n[], m[], c[]) for ArtZernikeDesc."""
        keys = list(self.coefficients)
        return (float(self.R), [int(k[0]) for k in keys], [int(k[1]) for k in keys],
                [float(self.coefficients[k]) for k in keys])
