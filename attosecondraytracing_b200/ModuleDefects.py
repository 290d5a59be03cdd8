"""Surface defects -- ModuleDefects.Zernike of the reference (ART/ModuleDefects.py:149-181).

Per ray the polynomials are evaluated inside the CUDA trace kernel (csrc/art_device.cuh
zernike_eval) from a shared-memory table.  The host methods below evaluate ONE point (chief-ray
alignment, user inspection) with the same factorisation the kernel uses:
    Z_(n,m) = R_n^l(rho) * cos(l theta)  for 2m >= n,   R_n^l(rho) * sin(l theta)  for 2m < n,   l = |2m - n|
(unnormalised; ART/recursive_zernike_generator.py's index convention: (1,0) = y, (1,1) = x).
The gridded defects of the reference (MeasuredMap, Fourrier: a height map and two slope maps on a
regular grid, ART/ModuleDefects.py:34-146) are built on the host exactly as the reference builds them
and evaluated per ray on the device by bilinear interpolation (csrc/art_optics.cuh gridmap_eval).
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


class GridDefect(Defect):
    """A gridded defect: arrays `h`, `dx`, `dy` of shape (len(X), len(Y)) -- what the reference hands to
    its RegularGridInterpolators -- on X = linspace(x0, x1), Y = linspace(y0, y1)."""

    def _set_grid(self, h, dx, dy, x0, x1, y0, y1):
        self._h = np.ascontiguousarray(h, dtype=np.float64)
        self._dx = np.ascontiguousarray(dx, dtype=np.float64)
        self._dy = np.ascontiguousarray(dy, dtype=np.float64)
        self._extent = (float(x0), float(x1), float(y0), float(y1))
        self._device_cache = {}

    def _interp(self, V, Point):
        x0, x1, y0, y1 = self._extent
        nx, ny = V.shape
        fx = (float(Point[0]) - x0) / (x1 - x0) * (nx - 1)
        fy = (float(Point[1]) - y0) / (y1 - y0) * (ny - 1)
        ix = min(max(int(math.floor(fx)), 0), nx - 2)
        iy = min(max(int(math.floor(fy)), 0), ny - 2)
        tx, ty = fx - ix, fy - iy
        return ((1 - tx) * (1 - ty) * V[ix, iy] + tx * (1 - ty) * V[ix + 1, iy] + (1 - tx) * ty * V[ix, iy + 1]
                + tx * ty * V[ix + 1, iy + 1])

    def get_offset(self, Point):
        return self._interp(self._h, Point)

    def get_normal(self, Point):
        """(dX, dY, 1)/norm as the reference returns it (NOT negated, ART/ModuleDefects.py:52-58,119-129)."""
        dX, dY = self._interp(self._dx, Point), self._interp(self._dy, Point)
        norm = math.sqrt(dX * dX + dY * dY + 1)
        dX, dY = dX / norm, dY / norm
        return np.array([dX, dY, math.sqrt(1 - dX**2 - dY**2)])

    def RMS(self):
        return self.rms

    def PV(self):
        pass

    def _lower(self, device=None):
        """(nx, ny, x0, x1, y0, y1, h, dx, dy, keepalive): array pointers on `device` (torch CUDA tensors,
        cached) or in host memory when device is None (the CPU check of the device code)."""
        nx, ny = self._h.shape
        if device is None:
            arrs = (self._h, self._dx, self._dy)
            ptrs = [a.ctypes.data for a in arrs]
        else:
            import torch
            key = str(device)
            if key not in self._device_cache:
                self._device_cache[key] = tuple(torch.from_numpy(a).to(device) for a in (self._h, self._dx, self._dy))
            arrs = self._device_cache[key]
            ptrs = [a.data_ptr() for a in arrs]
        return (nx, ny) + self._extent + tuple(ptrs) + (arrs,)


class MeasuredMap(GridDefect):
    """MeasuredMap(Support, Map): a measured height map (mm), ART/ModuleDefects.py:34-61.

    The reference's semantics, kept literally: the map spans X in [-rect_x, rect_x], Y in [-rect_y, rect_y] with
    rect = Support._CircumRect(); slopes are `np.gradient(Map, rect_x / n0, rect_y / n1)` (the reference writes
    `np.gradient(Map, rect / Map.shape)`, which numpy rejects; one spacing per axis is the evident intent -- the
    compatibility patch of oracle/make_ref.py); the interpolators pair the grid (X[n0], Y[n1]) with the
    TRANSPOSED arrays (:45-47), so the value at (X[i], Y[j]) is Map[j, i] -- which, as in the reference, only fits
    square maps.  Pinned by tests/golden/par_measured_*.npz (oracle/gen_golden_gridmap.py)."""

    def __init__(self, Support, Map):
        self.deformation = np.asarray(Map, dtype=np.float64)
        if self.deformation.ndim != 2 or self.deformation.shape[0] != self.deformation.shape[1]:
            raise ValueError("MeasuredMap needs a square 2-D map (the reference pairs an (n0, n1) grid with the "
                             "transposed map, ART/ModuleDefects.py:45-47)")
        self.Support = Support
        rect = Support._CircumRect()
        spacing = rect / self.deformation.shape
        self.DerivX, self.DerivY = np.gradient(self.deformation, spacing[0], spacing[1])
        self.rms = np.std(self.deformation)
        self._set_grid(np.transpose(self.deformation), np.transpose(self.DerivX), np.transpose(self.DerivY),
                       -rect[0], rect[0], -rect[1], rect[1])


class Fourrier(GridDefect):
    """Fourrier(Support, RMS, slope=-2, smallest=0.1, biggest=None): a random rough surface with a
    power-law spectrum between the wavelengths `smallest` and `biggest` (mm), scaled to the given RMS
    (ART/ModuleDefects.py:69-117).  The phases come from numpy's global RNG as in the reference; `seed`
    (an addition) seeds it first so that a surface can be reproduced."""

    def __init__(self, Support, RMS, slope=-2, smallest=0.1, biggest=None, seed=None):
        rect = Support._CircumRect()
        if biggest is None:
            biggest = np.max(rect)
        k_max, k_min = 2 / smallest, 2 / biggest
        ResX = int(round(k_max * rect[0] / 2)) + 1
        ResY = int(round(k_max * rect[1]))
        kXX, kYY = np.meshgrid(np.linspace(0, k_max, num=ResX, dtype="float32", endpoint=False),
                               np.linspace(-k_max, k_max, num=ResY, dtype="float32", endpoint=False), sparse=True)
        band = np.ma.masked_outside(np.sqrt(kXX**2 + kYY**2), k_min, k_max)
        if seed is not None:
            np.random.seed(seed)
        spectrum = band**slope * np.exp(1j * np.random.uniform(0, 2 * np.pi, size=band.shape).astype("float32"))
        spectrum = spectrum.data * (1 - spectrum.mask)
        deformation = np.fft.irfft2(np.fft.ifftshift(spectrum, axes=0))
        factor = RMS / np.std(deformation)
        deformation *= factor
        DerivX = np.fft.irfft2(np.fft.ifftshift(spectrum * 1j * kXX * factor, axes=0)) * np.pi / 2
        kY = np.concatenate((kYY[kYY.shape[0] // 2:], kYY[:kYY.shape[0] // 2]))
        DerivY = np.fft.irfft2(np.fft.ifftshift(spectrum * 1j * factor, axes=0) * kY) * np.pi / 2
        self.DerivX, self.DerivY = DerivX, DerivY
        self.deformation = deformation
        self.rms = np.std(deformation)
        self.support = Support
        self._set_grid(np.transpose(deformation), np.transpose(DerivX), np.transpose(DerivY),
                       -rect[0] / 2, rect[0] / 2, -rect[1] / 2, rect[1] / 2)


class RawGridMap(GridDefect):
    """A gridded defect given directly by its interpolation arrays (test fixtures, external maps)."""

    def __init__(self, h, dx, dy, x0, x1, y0, y1):
        self.rms = float(np.std(h))
        self._set_grid(h, dx, dy, x0, x1, y0, y1)


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
