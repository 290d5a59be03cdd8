"""Host-side vector helpers for scene construction (poses, alignment, sources).

Mirrors the handful of ART/ModuleGeometry.py functions that scene-building code and user scripts
call on single vectors; nothing here is applied per ray -- rays live on the GPU
(ModuleOpticalRay.RayBundle) and are transformed by the CUDA kernels.
"""
from __future__ import annotations

import math

import numpy as np


def Normalize(vector):
    """Unit vector along `vector` (ART/ModuleGeometry.py:17-19)."""
    v = np.asarray(vector, dtype=np.float64)
    return v / np.linalg.norm(v)


def VectorPerpendicular(vector):
    """A unit vector perpendicular to `vector` (ART/ModuleGeometry.py:22-29)."""
    v = np.asarray(vector, dtype=np.float64)
    if abs(v[0]) < 1e-15:
        return np.array([1.0, 0.0, 0.0])
    return Normalize(np.array([-(v[1] + v[2]) / v[0], 1.0, 1.0]))


def AngleBetweenTwoVectors(U, V):
    """Kahan's well-conditioned angle formula (ART/ModuleGeometry.py:40-44)."""
    U = np.asarray(U, dtype=np.float64)
    V = np.asarray(V, dtype=np.float64)
    u, v = np.linalg.norm(U), np.linalg.norm(V)
    return 2 * math.atan2(np.linalg.norm(U * v - V * u), np.linalg.norm(U * v + V * u))


def RotationMatrixAroundAxis(Axis, Angle):
    """3x3 matrix of the right-handed rotation by Angle about Axis."""
    k = Normalize(Axis)
    c, s = math.cos(Angle), math.sin(Angle)
    K = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return c * np.eye(3) + s * K + (1.0 - c) * np.outer(k, k)


def RotationAroundAxis(Axis, Angle, Vector):
    """Rotate Vector by Angle (rad) about Axis (ART/ModuleGeometry.py:321-329, there by quaternions)."""
    return RotationMatrixAroundAxis(Axis, Angle) @ np.asarray(Vector, dtype=np.float64)


def RotationMatrix(Axis1, Axis2):
    """Matrix M with RotationPoint(P, Axis1, Axis2) == M @ P (ART/ModuleGeometry.py:333-343):
    identity below 1e-10 rad, MINUS identity within 1e-10 of pi (as the reference), else the
    rotation by the angle between the axes about their cross product."""
    ang = AngleBetweenTwoVectors(Axis1, Axis2)
    if abs(ang) < 1e-10:
        return np.eye(3)
    if abs(ang - math.pi) < 1e-10:
        return -np.eye(3)
    return RotationMatrixAroundAxis(np.cross(Axis1, Axis2), ang)


def RotationPoint(Point, Axis1, Axis2):
    """Rotate Point by the rotation that takes Axis1 onto Axis2 (ART/ModuleGeometry.py:333-343)."""
    return RotationMatrix(Axis1, Axis2) @ np.asarray(Point, dtype=np.float64)


def SpiralVogel(NbPoint, Radius):
    """NbPoint points filling a disk of radius Radius on Vogel's spiral (ART/ModuleGeometry.py:61-76)."""
    k = np.arange(NbPoint, dtype=np.float64)
    golden = np.pi * (3 - np.sqrt(5))
    r = np.sqrt(k / NbPoint) * Radius
    return np.column_stack([np.cos(golden * k) * r, np.sin(golden * k) * r])


def IncludeRectangle(X, Y, Point):
    """Inclusive point-in-rectangle test (ART/ModuleGeometry.py:249-255)."""
    return bool(abs(Point[0]) <= abs(X / 2) and abs(Point[1]) <= abs(Y / 2))


def IncludeDisk(R, Point):
    """Inclusive point-in-disk test (ART/ModuleGeometry.py:259-268)."""
    return bool(Point[0] ** 2 + Point[1] ** 2 <= R**2)


def normal_add(N1, N2):
    """Add the slopes of two surface normals (ART/ModuleGeometry.py:394-407)."""
    n1, n2 = Normalize(N1), Normalize(N2)
    gx = -n1[0] / n1[2] - n2[0] / n2[2]
    gy = -n1[1] / n1[2] - n2[1] / n2[2]
    return np.array([-gx, -gy, 1.0])


# ----------------------------------------------------------------------------------------------
# The remaining helpers of ART/ModuleGeometry.py under their names.  The reference applies them ray by ray
# in Python; the trace itself does not use them here (the CUDA kernel has its own frame transforms and root
# selection).  Point lists may be (n, d) arrays; ray lists may be RayBundle objects (transformed with tensor
# operations on the bundle's device, returning a new RayBundle) or lists of Ray (returning a list).
# ----------------------------------------------------------------------------------------------
def IntersectionLinePlane(A, u, P, n):
    """Point where the line A + t u meets the plane through P with normal n (ART/ModuleGeometry.py:48-57;
    the sign of t is not checked)."""
    A, u, P, n = (np.asarray(v, dtype=np.float64) for v in (A, u, P, n))
    return A + u * (np.dot(P - A, n) / np.dot(u, n))


def _real_roots(coefficients):
    r = np.roots(coefficients)
    return [float(x.real) for x in r if abs(x.imag) < 1e-15]


def SolverQuadratic(a, b, c):
    """Real roots of a x^2 + b x + c (numpy.roots, |imag| < 1e-15; ART/ModuleGeometry.py:80-91)."""
    return _real_roots([a, b, c])


def SolverQuartic(a, b, c, d, e):
    """Real roots of a x^4 + b x^3 + c x^2 + d x + e (ART/ModuleGeometry.py:95-106)."""
    return _real_roots([a, b, c, d, e])


def KeepPositiveSolution(SolutionList):
    """The entries larger than 1e-12 (ART/ModuleGeometry.py:110-120)."""
    return [k for k in SolutionList if k > 1e-12]


def KeepNegativeSolution(SolutionList):
    """The entries smaller than -1e-12 (ART/ModuleGeometry.py:124-134)."""
    return [k for k in SolutionList if k < -1e-12]


def _dist2(A, B):
    d = np.asarray(B, dtype=np.float64) - np.asarray(A, dtype=np.float64)
    return float(np.dot(d, d))


def ClosestPoint(A, I1, I2):
    """Whichever of I1, I2 is nearer to A; I2 on a tie (ART/ModuleGeometry.py:138-147)."""
    return I1 if _dist2(A, I1) < _dist2(A, I2) else I2


def FarestPoint(A, I1, I2):
    """Whichever of I1, I2 is farther from A; I2 on a tie (ART/ModuleGeometry.py:151-160)."""
    return I1 if _dist2(A, I1) > _dist2(A, I2) else I2


def DiameterPointList(PointList):
    """Largest bounding-box extent of a 2-D or 3-D point cloud; None for an empty one
    (ART/ModuleGeometry.py:164-218)."""
    if len(PointList) == 0:
        return None
    pts = np.asarray(PointList, dtype=np.float64)
    return float(np.max(np.abs(pts.max(axis=0) - pts.min(axis=0))))


def CentrePointList(PointList):
    """The 2-D points shifted so that the MIDPOINT OF THEIR BOUNDING BOX (not their mean) is the origin
    (ART/ModuleGeometry.py:222-245).  Returns an (n, 2) array."""
    pts = np.asarray(PointList, dtype=np.float64).reshape(-1, 2)
    return pts - 0.5 * (pts.max(axis=0) + pts.min(axis=0))


def SymmetricalVector(V, SymmetryAxis):
    """V rotated by pi about SymmetryAxis (ART/ModuleGeometry.py:272-276)."""
    return RotationAroundAxis(SymmetryAxis, math.pi, V)


def TranslationPoint(Point, T):
    return np.asarray(Point, dtype=np.float64) + np.asarray(T, dtype=np.float64)


def TranslationPointList(PointList, T):
    """All points shifted by T (ART/ModuleGeometry.py:290-297); an (n, d) array."""
    return np.asarray(PointList, dtype=np.float64) + np.asarray(T, dtype=np.float64)


def RotationPointList(PointList, Axis1, Axis2):
    """All points rotated by the rotation that takes Axis1 onto Axis2 (ART/ModuleGeometry.py:347-354)."""
    return np.asarray(PointList, dtype=np.float64) @ RotationMatrix(Axis1, Axis2).T


def _transform_bundle(bundle, matrix=None, shift=None, rotate_points=True):
    """New RayBundle with points M p + T and directions M u (M, T optional); flags, numbers, paths,
    incidences and intensities carried over."""
    import torch
    from .ModuleOpticalRay import RayBundle
    src = bundle.materialize() if (matrix is not None and rotate_points) or shift is not None else bundle
    out = RayBundle(src.n, device=src.device, columns=src._names, wavelength=src.wavelength,
                    storage=src._storage.clone())
    out.alive = None if src.alive is None else src.alive.clone()
    out.number, out.origin = src.number, None if src.origin is None else src.origin.clone()
    if matrix is not None:
        M = torch.as_tensor(np.asarray(matrix, dtype=np.float64), device=src.device)
        groups = [("ux", "uy", "uz")] + ([("px", "py", "pz")] if rotate_points else [])
        for names in groups:
            if not all(n in src._names for n in names):
                continue
            v = torch.stack([src.col(n) for n in names])
            w = M @ v
            for i, n in enumerate(names):
                out.col(n).copy_(w[i])
        if rotate_points and out.origin is not None:
            out.origin = M @ out.origin
    if shift is not None:
        T = np.asarray(shift, dtype=np.float64)
        for i, n in enumerate(("px", "py", "pz")):
            out.col(n).add_(float(T[i]))
    return out


def _is_bundle(rays):
    from .ModuleOpticalRay import RayBundle
    return isinstance(rays, RayBundle)


def TranslationRay(Ray, T):
    """A copy of Ray with its point shifted by T (ART/ModuleGeometry.py:300-304)."""
    r = Ray.copy_ray()
    r.point = r.point + np.asarray(T, dtype=np.float64)
    return r


def TranslationRayList(RayList, T):
    """All rays shifted by T (ART/ModuleGeometry.py:308-315)."""
    if _is_bundle(RayList):
        return _transform_bundle(RayList, shift=T)
    return [TranslationRay(r, T) for r in RayList]


def RotationRay(Ray, Axis1, Axis2):
    """A copy of Ray with point and direction rotated by the rotation that takes Axis1 onto Axis2
    (ART/ModuleGeometry.py:357-368, which rotates the points A and A + u and takes their difference)."""
    M = RotationMatrix(Axis1, Axis2)
    r = Ray.copy_ray()
    r.point = M @ Ray.point
    r.vector = M @ Ray.vector
    return r


def RotationRayList(ListeRay, Axis1, Axis2):
    """All rays rotated by the rotation that takes Axis1 onto Axis2 (ART/ModuleGeometry.py:372-378)."""
    if _is_bundle(ListeRay):
        return _transform_bundle(ListeRay, matrix=RotationMatrix(Axis1, Axis2))
    return [RotationRay(r, Axis1, Axis2) for r in ListeRay]


def RotationAroundAxisRayList(ListeRay, Axis, Angle):
    """The DIRECTIONS of all rays rotated by Angle (rad) about Axis, points unchanged
    (ART/ModuleGeometry.py:382-390)."""
    M = RotationMatrixAroundAxis(Axis, Angle)
    if _is_bundle(ListeRay):
        return _transform_bundle(ListeRay, matrix=M, rotate_points=False)
    out = []
    for r in ListeRay:
        q = r.copy_ray()
        q.vector = M @ r.vector
        out.append(q)
    return out
