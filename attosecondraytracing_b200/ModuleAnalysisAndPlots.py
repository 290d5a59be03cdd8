"""Result summaries of ART/ModuleAnalysisAndPlots.py (:62-129).  The plotting / rendering functions
of that module (matplotlib, pyvista) are out of scope of this package."""
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
