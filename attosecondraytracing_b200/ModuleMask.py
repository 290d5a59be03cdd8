"""Mask -- a plane that blocks rays hitting its support and transmits the others (ART/ModuleMask.py:21).

The per-ray transmission test (ART/ModuleMask.py:51-61, 93-136) runs in the CUDA trace kernel.
"""
from __future__ import annotations

import numpy as np

from . import _cabi


class Mask:
    def __init__(self, Support):
        self.type = "Mask"
        self.support = Support

    def get_normal(self, Point):
        return np.array([0.0, 0.0, 1.0])

    def get_centre(self):
        return np.array([0.0, 0.0, 0.0])

    def _lower(self):
        """(ART_SURF_* kind, four surface parameters)."""
        return _cabi.SURF_MASK, [0, 0, 0, 0]


def TransmitMaskRayList(Mask, RayList):
    """The rays of RayList (RayBundle or list[Ray], given in the mask's own frame) that pass the mask, moved
    to the mask plane with incidence angle and path updated; rays that hit the support are dropped
    (ART/ModuleMask.py:112-136).  Returns a RayBundle, computed by the CUDA trace kernel."""
    from .ModuleMirror import _element_frame_trace
    return _element_frame_trace(Mask, RayList, True)
