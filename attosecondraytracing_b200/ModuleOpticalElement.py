"""OpticalElement -- an optic together with its pose in the lab frame (ART/ModuleOpticalElement.py:23).

The pose triple (position, normal, majoraxis) is what the tracer consumes: the lab->element
rotation is derived from it inside libart_b200 (art_element_rotation) with the reference's
branching.  The (mis-)alignment methods edit the pose on the host exactly as the reference's do.
"""
from __future__ import annotations

import numpy as np

from . import ModuleGeometry as mgeo


def _is_vec3(v):
    return type(v) == np.ndarray and len(v) == 3


class OpticalElement:
    def __init__(self, Type, Position, Normal, MajorAxis):
        self._type = Type
        self.position = Position
        self.normal = mgeo.Normalize(Normal)
        self.majoraxis = mgeo.Normalize(MajorAxis)

    # ---- pose -------------------------------------------------------------------------------
    @property
    def position(self):
        return self._position

    @position.setter
    def position(self, NewPosition):
        if not _is_vec3(NewPosition):
            raise TypeError("Position must be a 3D numpy.ndarray.")
        self._position = NewPosition

    @property
    def normal(self):
        return self._normal

    @normal.setter
    def normal(self, NewNormal):
        if not (_is_vec3(NewNormal) and np.linalg.norm(NewNormal) > 0):
            raise TypeError("Normal must be a 3D numpy.ndarray with finite length.")
        new = mgeo.Normalize(NewNormal)
        # keep the major axis perpendicular: co-rotate it with the normal (:125-141).  During
        # construction there is no major axis yet.
        major = getattr(self, "_majoraxis", None)
        if major is not None and abs(np.dot(new, major)) > 1e-12:
            try:
                self._majoraxis = mgeo.RotationAroundAxis(
                    np.cross(self._normal, NewNormal), mgeo.AngleBetweenTwoVectors(self._normal, NewNormal), major)
            except Exception:
                pass
        self._normal = new

    @property
    def majoraxis(self):
        return self._majoraxis

    @majoraxis.setter
    def majoraxis(self, NewMajorAxis):
        if not (_is_vec3(NewMajorAxis) and np.linalg.norm(NewMajorAxis) > 0):
            raise TypeError("MajorAxis must be a 3D numpy.ndarray with finite length.")
        new = mgeo.Normalize(NewMajorAxis)
        if abs(np.dot(self.normal, new)) > 1e-12:
            raise ValueError("The normal and major axis of optical elements need to be orthogonal!")
        self._majoraxis = new

    @property
    def type(self):
        return self._type

    def __hash__(self):
        pose = tuple(self.position.ravel()) + tuple(self.normal.ravel()) + tuple(self.majoraxis.ravel())
        return hash(pose) + hash(self.type)

    def _pose_key(self):
        return (self.position.tobytes(), self.normal.tobytes(), self.majoraxis.tobytes(), id(self._type))

    # ---- (mis-)alignment (:169-265); angles in degrees, distances in mm ---------------------------
    def rotate_pitch_by(self, angle):
        """Rotate about normal x majoraxis."""
        axis = np.cross(self.normal, self.majoraxis)
        self.normal = mgeo.RotationAroundAxis(axis, np.deg2rad(angle), self.normal)

    def rotate_roll_by(self, angle):
        """Rotate about the major axis."""
        self.normal = mgeo.RotationAroundAxis(self.majoraxis, np.deg2rad(angle), self.normal)

    def rotate_yaw_by(self, angle):
        """Rotate about the normal."""
        self.majoraxis = mgeo.RotationAroundAxis(self.normal, np.deg2rad(angle), self.majoraxis)

    def rotate_random_by(self, angle):
        """Rotate about a random axis (np.random, unseeded as in the reference)."""
        self.normal = mgeo.RotationAroundAxis(np.random.random(3), np.deg2rad(angle), self.normal)

    def shift_along_normal(self, distance):
        self.position = self.position + distance * self.normal

    def shift_along_major(self, distance):
        self.position = self.position + distance * self.majoraxis

    def shift_along_cross(self, distance):
        self.position = self.position + distance * mgeo.Normalize(np.cross(self.normal, self.majoraxis))

    def shift_along_random(self, distance):
        self.position = self.position + distance * mgeo.Normalize(np.random.random(3))
