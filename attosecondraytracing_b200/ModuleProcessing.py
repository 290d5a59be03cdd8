"""The trace driver and bundle statistics -- the ART/ModuleProcessing.py entry points of the hot path.

    RayTracingCalculation(source_rays, optical_elements, IgnoreDefects=True)   ART/ModuleProcessing.py:250
    OEPlacement(SourceProperties, OpticsList, DistanceList, IncidenceAngleList, ...)   :133
    FindCentralRay :464, StandardDeviation :485, WeightedStandardDeviation :510, ReturnNumericalAperture :536

`RayTracingCalculation` lowers the element list to the packed table of libart_b200 and launches the
fused CUDA trace kernel; it returns one RayBundle per optical element (a lazy, list-like view of the
surviving rays).  There is no CPU path.
"""
from __future__ import annotations

import copy
import math

import numpy as np
import torch

from . import ModuleGeometry as mgeo
from .ModuleMirror import DeformedMirror
from .ModuleOpticalElement import OpticalElement
from .ModuleOpticalRay import Ray, RayBundle


# ----------------------------------------------------------------------------------------------
# the trace
# ----------------------------------------------------------------------------------------------
def _as_device_bundle(source_rays, device=None):
    from .engine import require_cuda
    dev = require_cuda(device)
    if isinstance(source_rays, RayBundle):
        return source_rays if source_rays.device.type == "cuda" else source_rays.to(dev)
    return RayBundle.from_rays(list(source_rays), device=dev)


def RayTracingCalculation(source_rays, optical_elements, IgnoreDefects=True, device=None):
    """Trace `source_rays` (RayBundle or list[Ray]) through `optical_elements` (list[OpticalElement]).

    Returns a list with one RayBundle per optical element: item k is the bundle after element k, in
    the lab frame, rays that missed removed, order and Ray.number preserved -- the
    `list[list[Ray]]` of ART/ModuleProcessing.py:250-313."""
    from .engine import DeviceChain
    src = _as_device_bundle(source_rays, device)
    chain = DeviceChain(list(optical_elements), device=src.device)
    try:
        outs, _ = chain.trace(src, ignore_defects=IgnoreDefects, history=True, want_central=False)
        torch.cuda.current_stream().synchronize()
    finally:
        chain.close()
    return outs


# ----------------------------------------------------------------------------------------------
# automatic placement (host-side scene construction; one chief ray, no bundle involved)
# ----------------------------------------------------------------------------------------------
def _chief_direction_after(optic, element, direction):
    """Direction of the chief ray after `optic`.  By construction the chief ray meets the optic at
    its centre point `get_centre()` (the element is positioned there), so the reflection is the
    mirror law about the surface normal at that point -- the result the reference obtains by tracing
    a one-ray chain (ART/ModuleProcessing.py:114-118)."""
    from . import _cabi
    R = np.array(_cabi.element_rotation(element.normal, element.majoraxis))
    u = mgeo.Normalize(R @ direction)
    C = optic.get_centre()
    # every mirror class tests its support at the (x, y) of the hit relative to the support centre,
    # which for the chief ray is (0, 0); the reference dies with IndexError when it is blocked (:118)
    if not optic.support._IncludeSupport(np.zeros(3)):
        raise IndexError("the chief ray does not hit the support of " + optic.type + " -- cannot align the chain")
    P = C.copy()
    base = optic.Mirror if isinstance(optic, DeformedMirror) else optic
    if isinstance(optic, DeformedMirror):
        # DeformedMirror._get_intersection, ART/ModuleMirror.py:969-980
        h = sum(d.get_offset(P - C) for d in optic.DeformationList)
        alpha = mgeo.AngleBetweenTwoVectors(-u, base.get_normal(P))
        P = P - u * h / math.cos(alpha)
    n = base.get_normal(P)  # get_output_rays() default: IgnoreDefects=True
    out = mgeo.Normalize(u - 2.0 * np.dot(n, u) * n)
    return mgeo.Normalize(R.T @ out)


def place_optical_elements(OpticsList, DistanceList, IncidenceAngleList, IncidencePlaneAngleList=None):
    """The alignment loop of `_singleOEPlacement` (ART/ModuleProcessing.py:82-128): returns the list
    of OpticalElement for a source at the origin pointing along +x."""
    if IncidencePlaneAngleList is None:
        IncidencePlaneAngleList = [0.0] * len(OpticsList)
    plane = [np.deg2rad(a % 360) for a in IncidencePlaneAngleList]
    inc = [np.deg2rad(a % 360) for a in IncidenceAngleList]
    centre = np.array([0.0, 0.0, 0.0])
    chief = np.array([1.0, 0.0, 0.0])
    rot_axis = np.array([0.0, 1.0, 0.0])  # perpendicular to the incidence plane (initially x-z)
    elements = []
    for k, optic in enumerate(OpticsList):
        if optic.type in ("SphericalCX Mirror", "CylindricalCX Mirror"):
            inc[k] = np.pi - inc[k]  # convex mirrors are hit from the back side (:94-95)
        centre = chief * DistanceList[k] + centre
        if abs(plane[k] - np.pi) < 1e-10:
            rot_axis = -rot_axis
        else:
            rot_axis = mgeo.RotationAroundAxis(chief, -plane[k], rot_axis)
        normal = mgeo.RotationAroundAxis(rot_axis, -np.pi / 2 + inc[k], np.cross(chief, rot_axis))
        major = np.cross(rot_axis, normal)
        el = OpticalElement(optic, centre, normal, major)
        elements.append(el)
        if "Mirror" in optic.type:
            chief = _chief_direction_after(optic, el, chief)
        elif optic.type == "Mask":
            pass  # the chief ray passes undeviated (:119-126)
        else:
            raise NameError("I don`t recognize the type of optical element " + optic.type + ".")
    return elements


def _singleOEPlacement(SourceProperties, OpticsList, DistanceList, IncidenceAngleList, IncidencePlaneAngleList,
                       Description, device=None):
    from . import ModuleOpticalChain as moc
    from . import ModuleSource as msource
    elements = place_optical_elements(OpticsList, DistanceList, IncidenceAngleList, IncidencePlaneAngleList)
    source = msource.synthetic_source(SourceProperties, first_optic_support=OpticsList[0].support, device=device)
    return moc.OpticalChain(source, elements, Description)


def _which_indeces(lst):
    return [i for i, x in enumerate(lst) if isinstance(x, (list, np.ndarray))]


def OEPlacement(SourceProperties, OpticsList, DistanceList, IncidenceAngleList, IncidencePlaneAngleList=None,
                Description="", device=None):
    """Place the optics in the lab frame from distances (mm) and incidence angles (deg) and return an
    OpticalChain (ART/ModuleProcessing.py:133-246).  One entry of one of the three lists may itself
    be a list / array: then a list of OpticalChains is returned, one per value, carrying
    `loop_variable_name` / `loop_variable_value`."""
    if IncidencePlaneAngleList is None:
        IncidencePlaneAngleList = np.zeros(len(OpticsList)).tolist()
    nd, ni, npl = _which_indeces(DistanceList), _which_indeces(IncidenceAngleList), _which_indeces(IncidencePlaneAngleList)
    nested = ni + npl + nd
    if len(nested) > 1:
        raise ValueError("Only one element of one of the lists IncidenceAngleList, IncidencePlaneAngleList, or "
                         "DistanceList can be a list or array itself. Otherwise things get too tangled...")
    if not nested:
        return _singleOEPlacement(SourceProperties, OpticsList, DistanceList, IncidenceAngleList,
                                  IncidencePlaneAngleList, Description, device)
    i = nested[0]
    name = OpticsList[i].type + "_idx_" + str(i)
    if ni:
        name, loop_list = name + " incidence angle (deg)", IncidenceAngleList
    elif nd:
        name, loop_list = name + " distance (mm)", DistanceList
    else:
        name, loop_list = name + " incidence-plane angle rotation (deg)", IncidencePlaneAngleList
    values = copy.deepcopy(loop_list[i])
    chains = []
    for x in values:
        loop_list[i] = x
        ch = _singleOEPlacement(SourceProperties, OpticsList, DistanceList, IncidenceAngleList,
                                IncidencePlaneAngleList, Description, device)
        ch.loop_variable_name = name
        ch.loop_variable_value = x
        chains.append(ch)
    return chains


# ----------------------------------------------------------------------------------------------
# bundle statistics
# ----------------------------------------------------------------------------------------------
def _central_sums(RayList):
    """(sum U, sum P, count) of the surviving rays of a RayBundle -- a device reduction."""
    b = RayList
    idx = b.alive_index()
    n = int(idx.numel())
    cols = torch.stack([b.col(c)[idx].sum() for c in ("ux", "uy", "uz", "px", "py", "pz")]).cpu().numpy()
    return cols[:3], cols[3:], n


def FindCentralRay(RayList):
    """Ray through the mean point along the normalised mean direction of the bundle
    (ART/ModuleProcessing.py:464-482)."""
    if isinstance(RayList, RayBundle):
        su, sp, n = _central_sums(RayList)
        if n == 0:
            return None
        return Ray(Point=sp / n, Vector=su / n)
    P = np.mean([r.point for r in RayList], axis=0)
    V = np.mean([r.vector for r in RayList], axis=0)
    return Ray(Point=P, Vector=V)


def StandardDeviation(List):
    """Population standard deviation of a list of numbers, or sqrt(var_x + var_y (+ var_z)) of a list
    of points (ART/ModuleProcessing.py:485-507).  Accepts numpy arrays / torch tensors."""
    a = List.detach().cpu().numpy() if isinstance(List, torch.Tensor) else np.asarray(List)
    if a.ndim == 1:
        return float(np.std(a))
    return float(np.sqrt(np.var(a, axis=0).sum()))


def WeightedStandardDeviation(List, Weights):
    """Intensity-weighted analogue (ART/ModuleProcessing.py:510-532)."""
    a = List.detach().cpu().numpy() if isinstance(List, torch.Tensor) else np.asarray(List)
    w = Weights.detach().cpu().numpy() if isinstance(Weights, torch.Tensor) else np.asarray(Weights)
    avg = np.average(a, axis=0, weights=w)
    var = np.average((a - avg) ** 2, axis=0, weights=w)
    return float(np.sqrt(np.sum(var)))


def ReturnNumericalAperture(RayList, RefractiveIndex: float = 1):
    """sin of the largest angle between a ray and the central ray, times the refractive index
    (ART/ModuleProcessing.py:536-566).  Device reduction for RayBundles."""
    if not isinstance(RayList, RayBundle):
        RayList = RayBundle.from_rays(list(RayList))
    b = RayList
    su, _, n = _central_sums(b)
    cv = mgeo.Normalize(su / n)
    idx = b.alive_index()
    U = torch.stack([b.col("ux")[idx], b.col("uy")[idx], b.col("uz")[idx]], dim=1)
    c = torch.as_tensor(cv, dtype=torch.float64, device=U.device)
    ang = 2 * torch.atan2(torch.linalg.norm(U - c, dim=1), torch.linalg.norm(U + c, dim=1))
    return float(torch.sin(ang.max())) * RefractiveIndex


def ReturnAiryRadius(Wavelength: float, NumericalAperture: float) -> float:
    """Radius 1.22 * lambda / (2 NA) of the Airy disk in the unit of Wavelength; 0 for NA <= 1e-3 or an unknown
    wavelength, where the diffraction limit is meaningless (ART/ModuleProcessing.py:570-593)."""
    if Wavelength is None or not NumericalAperture > 1e-3:
        return 0
    return 1.22 * 0.5 * Wavelength / NumericalAperture


def FindOptimalDistance(Detector, RayList, OptFor="intensity", Amplitude=None, Precision=3, IntensityWeighted=False,
                        verbose=False):
    """Detector position that minimises the spot size ("size"), the duration ("duration") or
    SpotSize^2 * Duration ("intensity") within +-Amplitude of the current distance
    (ART/ModuleProcessing.py:369-460).  Returns (moved copy of Detector, OptSizeSpot mm, OptDuration fs).

    The reference re-intersects every ray with 80 trial detectors (and therefore subsamples to 1000
    random rays, ARTmain.py:168); here ONE kernel pass over ALL rays yields 32 sums from which the
    statistics at any detector shift follow in closed form, and the reference's search schedule is
    evaluated on those."""
    from .engine import optimal_shift_from_scan
    scan = Detector.get_scan_sums(RayList)
    st = Detector.get_statistics(RayList)
    first = Detector.get_distance()
    if verbose:
        print(f"Searching optimal detector position for *{OptFor}* ...", end="", flush=True)
    # the default Amplitude uses the angle to the bundle's OWN central ray (:424-434), which differs from the
    # detector's reference axis for a manually placed / tilted detector
    na = ReturnNumericalAperture(RayList, 1) if Amplitude is None else st["NA"]
    s, spot, dur, amp = optimal_shift_from_scan(scan, first, st["SpotSizeSD"], na, OptFor, Amplitude, Precision,
                                                IntensityWeighted)
    moving = Detector.copy_detector()
    moving.shiftByDistance(float(s))
    if not first - amp + 10**-Precision < moving.get_distance() < first + amp - 10**-Precision:
        print("There`s no minimum-size/duration focus in the searched range.")
    if verbose:
        print("\r\033[K", end="", flush=True)
    return moving, spot, dur


def save_compressed(obj, filename: str = None):
    """Pickle `obj` into an lzma-compressed file `<filename>_<i>.xz` (ART/ModuleProcessing.py:612-627).
    RayBundles inside `obj` are stored as host columns (not as millions of Ray objects)."""
    import lzma
    import os
    import pickle
    from datetime import datetime
    if not type(filename) == str:
        filename = "kept_data_" + datetime.now().strftime("%Y-%m-%d-%Hh%M")
    i = 0
    while os.path.exists(filename + f"_{i}.xz"):
        i += 1
    filename = filename + f"_{i}"
    with lzma.open(filename + ".xz", "wb") as f:
        pickle.dump(obj, f)
    print("Saved results to " + filename + ".xz.")
    print("->To reload from disk do: kept_data = mp.load_compressed('" + filename + "')")
    return filename


def load_compressed(filename: str):
    """Load an object saved by save_compressed (ART/ModuleProcessing.py:630-635); bundles come back on the host."""
    import lzma
    import pickle
    with lzma.open(filename + ".xz", "rb") as f:
        return pickle.load(f)


def _hash_list_of_objects(objs):
    """Summed hashes, as ART/ModuleProcessing.py:597-602 (kept for small host-side lists)."""
    return sum(hash(o) for o in objs)
