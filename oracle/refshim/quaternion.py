"""
TEST INFRASTRUCTURE ONLY -- stand-in for the `numpy-quaternion` package (moble/quaternion),
which the reference imports at ART/ModuleGeometry.py:13 and uses at :321-329 but which is
not installed in this image (no network).

Only what ModuleGeometry.RotationAroundAxis needs is provided:
  quaternion(x, y, z) / quaternion(w, x, y, z), Hamilton product, exp (reached through
  np.exp on an object scalar), conjugate (through np.conjugate), and the `.imag` 3-vector.

This file is written from the published quaternion algebra, not from the moble sources.
It is only ever placed on sys.path by oracle/refshim/load_reference.py.
"""
import math

import numpy as np


class quaternion:
    __slots__ = ("w", "x", "y", "z")

    def __init__(self, *a):
        if len(a) == 3:
            self.w = 0.0
            self.x, self.y, self.z = (float(v) for v in a)
        elif len(a) == 4:
            self.w, self.x, self.y, self.z = (float(v) for v in a)
        else:
            raise TypeError("quaternion takes 3 or 4 components")

    def __mul__(self, o):
        a, b = self, o
        return quaternion(
            a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
            a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
            a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w,
        )

    def exp(self):
        vn = math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)
        ew = math.exp(self.w)
        if vn == 0.0:
            return quaternion(ew, 0.0, 0.0, 0.0)
        s = ew * math.sin(vn) / vn
        return quaternion(ew * math.cos(vn), s * self.x, s * self.y, s * self.z)

    def conjugate(self):
        return quaternion(self.w, -self.x, -self.y, -self.z)

    @property
    def imag(self):
        return np.array([self.x, self.y, self.z])

    @property
    def real(self):
        return self.w

    def __repr__(self):
        return f"quaternion({self.w}, {self.x}, {self.y}, {self.z})"
