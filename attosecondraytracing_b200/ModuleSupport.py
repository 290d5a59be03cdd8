"""Aperture shapes of optics -- the five support classes of ART/ModuleSupport.py.

Constructor signatures and attribute names are the reference's.  The per-ray inclusion test runs
in the CUDA trace kernel (csrc/art_device.cuh in_support); `_IncludeSupport` here answers for a
single host-side point (used when aligning the chief ray and by user scripts).
The plotting helpers of the reference (_get_grid, _ContourSupport, ...) are out of scope.
"""
from __future__ import annotations

import numpy as np

from . import ModuleGeometry as mgeo
from . import _cabi


class Support:
    """Base class; subclasses provide `_IncludeSupport`, `_CircumRect`, `_CircumCirc`, `_lower`."""

    def _lower(self):
        """(ART_SUPP_* kind, six parameters) for ArtElementDesc."""
        raise NotImplementedError


class SupportRound(Support):
    """Round support (ART/ModuleSupport.py:46)."""

    def __init__(self, Radius: float):
        self.radius = Radius

    def _IncludeSupport(self, Point):
        return mgeo.IncludeDisk(self.radius, Point)

    def _CircumRect(self):
        return np.array([self.radius * 2, self.radius * 2])

    def _CircumCirc(self):
        return self.radius

    def _lower(self):
        return _cabi.SUPP_ROUND, [self.radius, 0, 0, 0, 0, 0]


class SupportRoundHole(Support):
    """Round support with a round hole (ART/ModuleSupport.py:109)."""

    def __init__(self, Radius: float, RadiusHole: float, CenterHoleX: float, CenterHoleY: float):
        self.radius = Radius
        self.radiushole = RadiusHole
        self.centerholeX = CenterHoleX
        self.centerholeY = CenterHoleY

    def _IncludeSupport(self, Point):
        shifted = (Point[0] - self.centerholeX, Point[1] - self.centerholeY)
        return mgeo.IncludeDisk(self.radius, Point) and not mgeo.IncludeDisk(self.radiushole, shifted)

    def _CircumRect(self):
        return np.array([self.radius * 2, self.radius * 2])

    def _CircumCirc(self):
        return self.radius

    def _lower(self):
        return _cabi.SUPP_ROUND_HOLE, [self.radius, self.radiushole, self.centerholeX, self.centerholeY, 0, 0]


class SupportRectangle(Support):
    """Rectangular support (ART/ModuleSupport.py:200)."""

    def __init__(self, DimensionX: float, DimensionY: float):
        self.dimX = DimensionX
        self.dimY = DimensionY

    def _IncludeSupport(self, Point) -> bool:
        return mgeo.IncludeRectangle(self.dimX, self.dimY, Point)

    def _CircumRect(self):
        return np.array([self.dimX, self.dimY])

    def _CircumCirc(self):
        return np.sqrt(self.dimX**2 + self.dimY**2) / 2

    def _lower(self):
        return _cabi.SUPP_RECT, [self.dimX, self.dimY, 0, 0, 0, 0]


class SupportRectangleHole(Support):
    """Rectangular support with a round hole (ART/ModuleSupport.py:273)."""

    def __init__(self, DimensionX: float, DimensionY: float, RadiusHole: float, CenterHoleX: float,
                 CenterHoleY: float):
        self.dimX = DimensionX
        self.dimY = DimensionY
        self.radiushole = RadiusHole
        self.centerholeX = CenterHoleX
        self.centerholeY = CenterHoleY

    def _IncludeSupport(self, Point):
        shifted = (Point[0] - self.centerholeX, Point[1] - self.centerholeY)
        return mgeo.IncludeRectangle(self.dimX, self.dimY, Point) and not mgeo.IncludeDisk(self.radiushole, shifted)

    def _CircumRect(self):
        return np.array([self.dimX, self.dimY])

    def _CircumCirc(self):
        return np.sqrt(self.dimX**2 + self.dimY**2) / 2

    def _lower(self):
        return _cabi.SUPP_RECT_HOLE, [self.dimX, self.dimY, self.radiushole, self.centerholeX, self.centerholeY, 0]


class SupportRectangleRectHole(Support):
    """Rectangular support with a rectangular hole (ART/ModuleSupport.py:373)."""

    def __init__(self, DimensionX: float, DimensionY: float, HoleX: float, HoleY: float, CenterHoleX: float,
                 CenterHoleY: float):
        self.dimX = DimensionX
        self.dimY = DimensionY
        self.holeX = HoleX
        self.holeY = HoleY
        self.centerholeX = CenterHoleX
        self.centerholeY = CenterHoleY

    def _IncludeSupport(self, Point):
        shifted = (Point[0] - self.centerholeX, Point[1] - self.centerholeY)
        return mgeo.IncludeRectangle(self.dimX, self.dimY, Point) and not mgeo.IncludeRectangle(
            self.holeX, self.holeY, shifted)

    def _CircumRect(self):
        return np.array([self.dimX, self.dimY])

    def _CircumCirc(self):
        return np.sqrt(self.dimX**2 + self.dimY**2) / 2

    def _lower(self):
        return _cabi.SUPP_RECT_RECT_HOLE, [self.dimX, self.dimY, self.holeX, self.holeY, self.centerholeX,
                                           self.centerholeY]
