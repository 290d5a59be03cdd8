"""Result summaries of ART/ModuleAnalysisAndPlots.py (:62-129) and the DATA of its SpotDiagram /
DelayGraph figures as device-side histograms and of MirrorProjection as per-ray arrays.  The matplotlib / pyvista drawing itself is out of scope
of this package."""
from __future__ import annotations

import numpy as np

from . import ModuleProcessing as mp
from .ModuleOpticalRay import RayBundle


def _intensity_sum(RayList):
    if isinstance(RayList, RayBundle):
        idx = RayList.alive_index()
        col = RayList.col("intensity") if RayList.has("intensity") else getattr(RayList, "shared_intensity", None)
        if col is None:
            raise TypeError("the bundle carries no intensities")
        return float(col[idx].sum())
    return sum(r.intensity for r in RayList)


def getETransmission(RayListIn, RayListOut) -> float:
    """Energy transmission in percent: summed intensity out over summed intensity in (:62-77)."""
    return 100 * _intensity_sum(RayListOut) / _intensity_sum(RayListIn)


def GetResultSummary(Detector, RayListAnalysed, verbose=False):
    """(focal spot size SD in mm, duration SD in fs) of the bundle on the detector (:81-129)."""
    s = Detector.get_statistics(RayListAnalysed)
    if verbose:
        print("At the detector distance of {:.3f} mm we get:\n".format(Detector.get_distance())
              + "Spatial std : {:.3f} μm and min-max: {:.3f} μm\n".format(s["SpotSizeSD"] * 1e3, s["Diameter"] * 1e3)
              + "Temporal std : {:.3e} fs and min-max : {:.3e} fs".format(s["DurationSD"], s["delay_max_fs"] - s["delay_min_fs"]))
    return s["SpotSizeSD"], s["DurationSD"]


def SpotDiagramData(RayListAnalysed, Detector, bins=(64, 64), ColorCoded=None):
    """The content of SpotDiagram (ART/ModuleAnalysisAndPlots.py:133-250) as a 2-D histogram instead of one
    scatter point per ray: returns (x_edges_um, y_edges_um, count, colour) with the axes in micrometres
    centred on the bounding-box midpoint like the reference's figure; `colour` is the per-bin mean of the
    quantity named by ColorCoded ("Intensity": arb. u., "Delay": fs) or None."""
    if ColorCoded not in (None, "Intensity", "Delay"):
        raise ValueError('ColorCoded must be None, "Intensity" or "Delay" (incidences are not binned)')
    h = Detector.get_histograms(RayListAnalysed, bins=bins, delay_bins=1)
    colour = None
    with np.errstate(invalid="ignore", divide="ignore"):
        if ColorCoded == "Intensity":
            colour = h["spot_intensity"] / h["spot_count"]
        elif ColorCoded == "Delay":
            colour = h["spot_delay"]
    return h["x_edges"] * 1e3, h["y_edges"] * 1e3, h["spot_count"], colour


def DelayGraphData(RayListAnalysed, Detector, delay_bins=128):
    """Delay distribution of the bundle on the detector (the third axis of DelayGraph,
    ART/ModuleAnalysisAndPlots.py:360-440): (delay_edges_fs, count, summed_intensity)."""
    h = Detector.get_histograms(RayListAnalysed, bins=(1, 1), delay_bins=delay_bins)
    return h["delay_edges"], h["delay_count"], h["delay_intensity"]


def MirrorProjectionData(OpticalChain, ReflectionNumber: int, Detector=None, ColorCoded=None):
    """The content of MirrorProjection (ART/ModuleAnalysisAndPlots.py:443-520): the impact points of the
    bundle after optical element `ReflectionNumber` in that element's SUPPORT frame (the element frame
    without the shift by the optic's centre) and the quantity the figure colour-codes.  Returns (x, y, z):
    x, y in mm; z = intensities, incidence angles in degrees, delays at `Detector` in fs, or None."""
    from . import ModuleGeometry as mgeo
    if ColorCoded not in (None, "Intensity", "Incidence", "Delay"):
        raise ValueError('ColorCoded must be None, "Intensity", "Incidence" or "Delay"')
    if ColorCoded == "Delay" and Detector is None:
        raise ValueError("If you want to project ray delays, you must specify a detector.")
    element = OpticalChain.optical_elements[ReflectionNumber]
    rays = OpticalChain.get_output_rays()[ReflectionNumber]
    ez, ex = np.array([0.0, 0.0, 1.0]), np.array([1.0, 0.0, 0.0])
    moved = mgeo.TranslationRayList(rays, -np.asarray(element.position, dtype=np.float64))
    moved = mgeo.RotationRayList(moved, element.normal, ez)
    moved = mgeo.RotationRayList(moved, mgeo.RotationPoint(element.majoraxis, element.normal, ez), ex)
    idx = moved.alive_index()
    x = moved.col("px")[idx].cpu().numpy()
    y = moved.col("py")[idx].cpu().numpy()
    z = None
    if ColorCoded == "Intensity":
        col = rays.col("intensity") if rays.has("intensity") else getattr(rays, "shared_intensity", None)
        if col is None:
            raise TypeError("the bundle carries no intensities")
        z = col[idx].cpu().numpy()
    elif ColorCoded == "Incidence":
        z = np.rad2deg(rays.col("incidence")[idx].cpu().numpy())
    elif ColorCoded == "Delay":
        z = np.asarray(Detector.get_Delays(rays))
    return x, y, z
