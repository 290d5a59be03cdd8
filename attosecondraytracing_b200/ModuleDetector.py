"""Detector -- a virtual plane that records where and when rays arrive (ART/ModuleDetector.py:25).

Same constructor, attributes and methods as the reference; the per-ray work (plane hits, in-plane
coordinates, total path lengths) and the reductions behind the statistics run in the CUDA detector
kernel (art_detector_moments).  Methods that the reference returns as Python lists return numpy
arrays over the surviving rays, in bundle order.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from .ModuleOpticalRay import RayBundle

LightSpeed = 299792458000  # mm / s  (ART/ModuleDetector.py:21)


def _is_number(x):
    return type(x) in (int, float, np.float64)


class Detector:
    def __init__(self, RefPoint, Centre=None, Normal=None):
        self.centre = Centre
        self.normal = Normal
        self.refpoint = RefPoint
        self._cache = None

    # ---- properties (checks as ART/ModuleDetector.py:49-97) ----------------------------------------
    @property
    def centre(self):
        return self._centre

    @centre.setter
    def centre(self, Centre):
        if not (Centre is None or (type(Centre) == np.ndarray and Centre.shape == (3,))):
            raise TypeError("Detector Centre must be a 3D-vector, given as numpy.ndarray of shape (3,).")
        self._centre = Centre
        self._cache = None

    @property
    def normal(self):
        return self._normal

    @normal.setter
    def normal(self, Normal):
        if Normal is None:
            self._normal = None
        elif type(Normal) == np.ndarray and Normal.shape == (3,) and np.linalg.norm(Normal) > 0:
            self._normal = Normal / np.linalg.norm(Normal)
        else:
            raise TypeError("Detector Normal must be a 3D-vector of norm >0, given as numpy.ndarray of shape (3,).")
        self._cache = None

    @property
    def refpoint(self):
        return self._refpoint

    @refpoint.setter
    def refpoint(self, RefPoint):
        if not (type(RefPoint) == np.ndarray and RefPoint.shape == (3,)):
            raise TypeError("Detector RefPoint must a 3D-vector, given as numpy.ndarray of shape (3,).")
        self._refpoint = RefPoint

    # ---- placement ---------------------------------------------------------------------------------
    def copy_detector(self):
        return Detector(self.refpoint, self.centre, self.normal)

    def autoplace(self, RayList, DistanceDetector: float):
        """Normal to the central ray of RayList, DistanceDetector behind its origin (:109-137)."""
        from .ModuleProcessing import FindCentralRay
        central = FindCentralRay(_bundle(RayList))
        normal = -central.vector
        self.normal = normal
        self.centre = central.point - self.normal * DistanceDetector
        self.refpoint = central.point

    def get_distance(self):
        """Distance of the detector plane from the reference point along the normal (:139-147)."""
        return float(abs(np.dot(self.normal, self.centre - self.refpoint)))

    def shiftToDistance(self, NewDistance: float):
        if not _is_number(NewDistance):
            raise TypeError("The new Detector Distance must be int or float.")
        self.centre = self.centre - (NewDistance - self.get_distance()) * self.normal

    def shiftByDistance(self, Shift: float):
        if not _is_number(Shift):
            raise TypeError("The Detector Distance Shift must be int or float.")
        self.centre = self.centre - Shift * self.normal

    def _iscomplete(self):
        if self.centre is None or self.normal is None:
            raise TypeError("The detector has no centre and normal vectors defined yet.")
        return True

    # ---- detector response -------------------------------------------------------------------------
    def _struct(self, l0=0.0):
        d = _cabi.ArtDetector()
        _cabi.check(_cabi.lib().art_detector_make(_cabi.vec3(self.centre), _cabi.vec3(self.normal),
                                                  _cabi.vec3(self.refpoint), float(l0), C.byref(d)))
        return d

    def _evaluate(self, RayList):
        """One pass of the detector kernel; cached per (bundle, detector pose)."""
        self._iscomplete()
        b = _bundle(RayList)
        key = (b.content_key(), self.centre.tobytes(), self.normal.tobytes())
        if self._cache is not None and self._cache[0] == key:
            return self._cache[1]
        if b.device.type != "cuda":
            from .engine import require_cuda
            b = b.to(require_cuda())
        idx = b.alive_index()
        # pivot for the path moments: mean of the stored paths plus the distance to the plane
        l0 = float(b.col("path")[idx].mean()) + self.get_distance() if (b.has("path") and idx.numel()) else 0.0
        d = self._struct(l0)
        det = torch.from_numpy(np.frombuffer(bytes(d), dtype=np.float64).copy()).to(b.device).reshape(1, -1)
        from .engine import DeviceChain
        mom, x, y, l = DeviceChain.moments(_Scratchless(b.device), b, det, want_points=True)
        torch.cuda.current_stream().synchronize()
        res = {"bundle": b, "idx": idx, "det": det, "moments": mom, "x": x, "y": y, "l": l,
               "rot": np.array(d.rot[:]).reshape(3, 3)}
        self._cache = (key, res)
        return res

    def get_PointList3D(self, RayList):
        """(n,3) lab-frame points where the rays meet the detector plane (:191-210)."""
        r = self._evaluate(RayList)
        xy = torch.stack([r["x"][r["idx"]], r["y"][r["idx"]]], dim=1).cpu().numpy()
        return xy @ r["rot"][:2, :] + self.centre

    def get_PointList2D(self, RayList):
        """(n,2) points in the detector plane, origin at Detector.centre (:212-234)."""
        r = self._evaluate(RayList)
        return torch.stack([r["x"][r["idx"]], r["y"][r["idx"]]], dim=1).cpu().numpy()

    def get_PointList2DCentre(self, RayList):
        """As get_PointList2D with the origin at the bounding-box midpoint of the cloud (:236-252)."""
        r = self._evaluate(RayList)
        m = r["moments"].cpu().numpy()[0]
        c = np.array([0.5 * (m[_cabi.M_XMAX] + m[_cabi.M_XMIN]), 0.5 * (m[_cabi.M_YMAX] + m[_cabi.M_YMIN])])
        return self.get_PointList2D(RayList) - c

    def get_Delays(self, RayList):
        """(n,) delays in fs relative to the mean travel time of the bundle (:254-279)."""
        r = self._evaluate(RayList)
        b = r["bundle"]
        out = torch.full_like(r["l"], float("nan"))
        _cabi.check(_cabi.lib().art_delays(C.c_void_p(r["l"].data_ptr()),
                                           C.c_void_p(b.alive.data_ptr()) if b.alive is not None else None, b.n, 1,
                                           C.c_void_p(r["det"].data_ptr()), C.c_void_p(r["moments"].data_ptr()),
                                           C.c_void_p(out.data_ptr()),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out[r["idx"]].cpu().numpy()

    def get_scan_sums(self, RayList):
        """The 32 scan sums of the bundle on this detector (art_detector_scan_moments): the input of the
        closed-form detector-distance optimiser (ModuleProcessing.FindOptimalDistance)."""
        from .engine import DeviceChain
        r = self._evaluate(RayList)
        scan = DeviceChain.scan_moments(_Scratchless(r["bundle"].device), r["bundle"], r["det"])
        torch.cuda.current_stream().synchronize()
        return scan.cpu().numpy()[0]

    def get_histograms(self, RayList, bins=(64, 64), delay_bins=128):
        """Binned detector response (art_detector_histogram) as the dict of `engine.split_histogram`: what
        SpotDiagram / DelayGraph (ART/ModuleAnalysisAndPlots.py:133, 360) scatter ray by ray, as
        fixed-size arrays that do not grow with the ray count."""
        from .engine import DeviceChain, split_histogram
        r = self._evaluate(RayList)
        b = r["bundle"]
        hist = DeviceChain.histogram(_Scratchless(b.device), b, r["det"], r["moments"], bins=bins,
                                     delay_bins=delay_bins)
        torch.cuda.current_stream().synchronize()
        return split_histogram(hist.cpu().numpy(), r["moments"].cpu().numpy()[0], bins=bins, delay_bins=delay_bins)

    def get_statistics(self, RayList, RayListIn=None):
        """All bundle statistics from ONE kernel pass (dict: SpotSizeSD mm, DurationSD fs, weighted
        variants, Diameter, NA, ...) without materialising per-ray lists."""
        from .engine import summary_from_moments
        r = self._evaluate(RayList)
        return summary_from_moments(r["moments"].cpu().numpy()[0])


class _Scratchless:
    """Stands in for a DeviceChain when a Detector is evaluated on its own: no ArtChain handle, the
    library's per-device scratch is used."""

    def __init__(self, device):
        self.device = device
        self._handle = None

    def _check_bundle(self, bundle):
        if bundle.device.type != "cuda":
            raise RuntimeError("the ray bundle must live on the CUDA device")


def _bundle(RayList):
    if isinstance(RayList, RayBundle):
        return RayList
    from .engine import require_cuda
    return RayBundle.from_rays(list(RayList), device=require_cuda())
