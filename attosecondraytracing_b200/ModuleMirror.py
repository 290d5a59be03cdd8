"""Mirror surfaces -- the six mirror classes of ART/ModuleMirror.py plus DeformedMirror.

These are scene-description objects with the reference's constructor signatures, attribute names
and `type` strings.  What the reference does per ray in Python (`_get_intersection`,
`ReflectionMirrorRayList`, ART/ModuleMirror.py:27-38, 878-939) happens in the fused CUDA trace
kernel (csrc/art_device.cuh); each class only lowers itself to the numbers that kernel needs
(`_lower`) and answers `get_centre()` / `get_normal(P)` for single host-side points (alignment).
`get_grid3D` (rendering meshes) is out of scope.
"""
from __future__ import annotations

import math

import numpy as np

from . import ModuleGeometry as mgeo
from . import _cabi


def _unit(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v)


class MirrorPlane:
    """Plane mirror in the element's xy-plane (ART/ModuleMirror.py:42)."""

    def __init__(self, Support):
        self.support = Support
        self.type = "Plane Mirror"

    def get_normal(self, Point):
        return np.array([0.0, 0.0, 1.0])

    def get_centre(self):
        return np.array([0.0, 0.0, 0.0])

    def _lower(self):
        return _cabi.SURF_PLANE, [0, 0, 0, 0]


class MirrorSpherical:
    """Spherical mirror, sphere centred on the element origin (ART/ModuleMirror.py:117).
    Radius > 0 concave ("SphericalCC Mirror"), < 0 convex ("SphericalCX Mirror")."""

    def __init__(self, Radius, Support):
        self.type = "SphericalCX Mirror" if Radius < 0 else "SphericalCC Mirror"
        self.radius = abs(Radius)
        self.support = Support

    def get_normal(self, Point):
        return _unit(-np.asarray(Point, dtype=np.float64))

    def get_centre(self):
        return np.array([0.0, 0.0, -self.radius])

    def _lower(self):
        return _cabi.SURF_SPHERICAL, [self.radius, 0, 0, 0]


class MirrorParabolic:
    """Off-axis paraboloid x^2 + y^2 = 2 p z (ART/ModuleMirror.py:212).  OffAxisAngle in degrees
    (stored in radians), p = feff (1 + cos(offaxisangle))."""

    def __init__(self, FocalEffective: float, OffAxisAngle: float, Support):
        self.support = Support
        self.type = "Parabolic Mirror"
        self._offaxisangle = np.deg2rad(OffAxisAngle)
        self._feff = FocalEffective
        self._update_p()

    def _update_p(self):
        self._p = self._feff * (1 + np.cos(self._offaxisangle))

    @property
    def offaxisangle(self):
        return self._offaxisangle

    @offaxisangle.setter
    def offaxisangle(self, OffAxisAngle):
        self._offaxisangle = np.deg2rad(OffAxisAngle)
        self._update_p()

    @property
    def feff(self):
        return self._feff

    @feff.setter
    def feff(self, FocalEffective):
        self._feff = FocalEffective
        self._update_p()

    @property
    def p(self):
        return self._p

    @p.setter
    def p(self, SemiLatusRectum):
        self._p = SemiLatusRectum
        self._feff = self._p / (1 + np.cos(self._offaxisangle))

    def get_normal(self, Point):
        return _unit(np.array([-Point[0], -Point[1], self._p]))

    def get_centre(self):
        # the point of the paraboloid seen from the focus under the off-axis angle (:357-365)
        return np.array([self._feff * np.sin(self._offaxisangle), 0.0,
                         self._p * 0.5 - self._feff * np.cos(self._offaxisangle)])

    def _lower(self):
        return _cabi.SURF_PARABOLIC, [self._p, 0, 0, 0]


class MirrorToroidal:
    """Toroid (sqrt(x^2+z^2) - R)^2 + y^2 = r^2 (ART/ModuleMirror.py:391); MajorRadius R is the
    distance to the centre of the minor circle."""

    def __init__(self, MajorRadius, MinorRadius, Support):
        self.majorradius = MajorRadius
        self.minorradius = MinorRadius
        self.support = Support
        self.type = "Toroidal Mirror"

    def get_normal(self, Point):
        # -grad of the implicit quartic (:480-498), common positive factor dropped
        x, y, z = (float(c) for c in Point)
        R, r = self.majorradius, self.minorradius
        s = x * x + y * y + z * z + (R * R - r * r)
        return _unit(-np.array([x * (s - 2 * R * R), y * s, z * (s - 2 * R * R)]))

    def get_centre(self):
        return np.array([0.0, 0.0, -self.majorradius - self.minorradius])

    def _lower(self):
        return _cabi.SURF_TOROIDAL, [self.majorradius, self.minorradius, 0, 0]


def ReturnOptimalToroidalRadii(Focal: float, AngleIncidence: float):
    """Astigmatism-free toroid radii for a focal length (mm) and incidence angle (deg)
    (ART/ModuleMirror.py:533-561)."""
    c = np.cos(AngleIncidence * np.pi / 180)
    return 2 * Focal * (1 / c - c), 2 * Focal * c


class MirrorEllipsoidal:
    """Ellipsoid of revolution (x/a)^2 + (y/b)^2 + (z/b)^2 = 1 (ART/ModuleMirror.py:565).

    Give (SemiMajorAxis, SemiMinorAxis) and/or (OffAxisAngle in degrees, f_object, f_image); the
    missing quantities follow from the focal geometry exactly as in the reference (:593-660)."""

    def __init__(self, Support, SemiMajorAxis=None, SemiMinorAxis=None, OffAxisAngle=None, f_object=None,
                 f_image=None):
        self.type = "Ellipsoidal Mirror"
        self.support = Support
        self.a = self.b = self._offaxisangle = None
        have_axes = SemiMajorAxis is not None and SemiMinorAxis is not None
        have_focals = f_object is not None and f_image is not None
        if have_axes:
            self.a, self.b = SemiMajorAxis, SemiMinorAxis
        if OffAxisAngle is not None:
            self._offaxisangle = np.deg2rad(OffAxisAngle)
            if have_focals:  # law of cosines in the triangle (focus, mirror centre, focus)
                d2 = f_object**2 + f_image**2 - 2 * f_object * f_image * np.cos(self._offaxisangle)
                self.a = (f_image + f_object) / 2
                self.b = np.sqrt(self.a**2 - d2 / 4)
        elif have_axes:
            d = 2 * np.sqrt(self.a**2 - self.b**2)
            if have_focals:
                self._offaxisangle = np.arccos((f_image**2 + f_object**2 - d**2) / (2 * f_image * f_object))
            else:  # mirror centre on the minor axis: both focal distances equal a
                self._offaxisangle = np.arccos(1 - d**2 / (2 * self.a**2))
        if self.a is None or self.b is None or self._offaxisangle is None:
            raise ValueError("Invalid mirror parameters")

    def get_normal(self, Point):
        return _unit(np.array([-Point[0] / self.a**2, -Point[1] / self.b**2, -Point[2] / self.b**2]))

    def get_centre(self):
        # Point of the ellipse (y = 0) from which the two foci subtend the off-axis angle: it lies on
        # the circle through both foci centred at (0, h) (inscribed-angle theorem), :695-714.
        d = 2 * np.sqrt(self.a**2 - self.b**2)
        h = -d / 2 / np.tan(self._offaxisangle)
        rad = np.sqrt(d**2 / 4 + h**2)
        sign = 1
        if math.isclose(self._offaxisangle, np.pi / 2):
            h = 0
        elif self._offaxisangle > np.pi / 2:
            h, sign = -h, -1
        qa = 1 - self.a**2 / self.b**2
        qb = -2 * h
        qc = self.a**2 + h**2 - rad**2
        z = (-qb + sign * np.sqrt(qb**2 - 4 * qa * qc)) / (2 * qa)
        if math.isclose(z**2, self.b**2):
            return np.array([0.0, 0.0, -self.b])
        return np.array([self.a * np.sqrt(1 - z**2 / self.b**2), 0.0, sign * z])

    def _lower(self):
        return _cabi.SURF_ELLIPSOIDAL, [self.a, self.b, 0, 0]


def ReturnOptimalEllipsoidalAxes(Focal: float, AngleIncidence: float):
    """Semi-axes for a focal length (mm) at an incidence angle (deg) (ART/ModuleMirror.py:755-777)."""
    return Focal, Focal * np.cos(np.deg2rad(AngleIncidence))


class MirrorCylindrical:
    """Cylinder y^2 + z^2 = R^2 with its axis along x (ART/ModuleMirror.py:781); sign of Radius as
    for MirrorSpherical."""

    def __init__(self, Radius, Support):
        self.type = "CylindricalCX Mirror" if Radius < 0 else "CylindricalCC Mirror"
        self.radius = abs(Radius)
        self.support = Support

    def get_normal(self, Point):
        return _unit(np.array([0.0, -Point[1], -Point[2]]))

    def get_centre(self):
        return np.array([0.0, 0.0, -self.radius])

    def _lower(self):
        return _cabi.SURF_CYLINDRICAL, [self.radius, 0, 0, 0]


class DeformedMirror:
    """A mirror with a list of surface defects (ART/ModuleMirror.py:945).  Only
    ModuleDefects.Zernike defects are supported by the CUDA path."""

    def __init__(self, Mirror, DeformationList):
        self.Mirror = Mirror
        self.DeformationList = DeformationList
        self.type = Mirror.type
        self.support = Mirror.support

    def get_normal(self, PointMirror):
        n = self.Mirror.get_normal(PointMirror)
        rel = np.asarray(PointMirror, dtype=np.float64) - self.get_centre()
        for d in self.DeformationList:
            n = _unit(mgeo.normal_add(n, d.get_normal(rel)))
        return n

    def get_centre(self):
        return self.Mirror.get_centre()

    def _lower(self):
        return self.Mirror._lower()


def _element_frame_trace(Optic, RayList, IgnoreDefects):
    """One optic acting on rays that are ALREADY in its own frame (what ReflectionMirrorRayList and
    TransmitMaskRayList of the reference take): a one-element chain whose element frame is the lab frame --
    position = the optic's centre, normal = ez, major axis = ex, so the frame rotation is the identity and
    p_e = (p - centre) + centre."""
    from .ModuleOpticalElement import OpticalElement
    from . import ModuleProcessing as mp
    element = OpticalElement(Optic, np.asarray(Optic.get_centre(), dtype=np.float64), np.array([0.0, 0.0, 1.0]),
                             np.array([1.0, 0.0, 0.0]))
    return mp.RayTracingCalculation(RayList, [element], IgnoreDefects=IgnoreDefects)[0]


def ReflectionMirrorRayList(Mirror, ListRay, IgnoreDefects=False):
    """The rays of ListRay (RayBundle or list[Ray], given in the mirror's own frame) reflected by Mirror:
    rays that miss the surface or its support are dropped, incidence angle and path are updated
    (ART/ModuleMirror.py:912-939; note its default IgnoreDefects=False, unlike RayTracingCalculation).
    Returns a RayBundle, computed by the CUDA trace kernel."""
    return _element_frame_trace(Mirror, ListRay, IgnoreDefects)
