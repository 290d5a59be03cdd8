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
