"""Light sources -- the ray-bundle generators of ART/ModuleSource.py, producing RayBundles.

The reference builds `list[Ray]` in Python loops; here the Vogel-spiral bundles are evaluated in
closed form by a CUDA kernel (art_source_generate) straight into the device columns, so a 10^8-ray
bundle (or one rank's slice of it) never exists on the host.  Ray numbers are the spiral indices.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import ModuleGeometry as mgeo
from . import _cabi
from .engine import _ptr, _stream, require_cuda
from .ModuleOpticalRay import RayBundle

_SRC_COLUMNS = ("px", "py", "pz", "ux", "uy", "uz", "intensity")


def _generate(kind, n_total, first, count, rho, axis, origin, wavelength, device, stride=1, extended=(0, 0, 0.0)):
    device = require_cuda(device)
    # a point source keeps ONE origin for all rays instead of three point columns (24 B/ray less)
    cols = _SRC_COLUMNS[3:] if kind == 0 else _SRC_COLUMNS
    b = RayBundle(count, device=device, columns=cols, wavelength=wavelength)
    if kind == 0:
        b.origin = torch.as_tensor(np.asarray(origin, dtype=np.float64), device=device).clone()
    if first != 0 or stride != 1:
        b.number = first + stride * torch.arange(count, device=device, dtype=torch.int64)
    v = b.view()
    v.intensity = None
    if kind == 0:
        v.px = v.py = v.pz = None  # the generator writes directions only
    with torch.cuda.device(device):
        _cabi.check(_cabi.lib().art_source_generate(kind, n_total, first, count, stride, float(rho), _cabi.vec3(axis),
                                                    _cabi.vec3(origin), int(extended[0]), int(extended[1]),
                                                    float(extended[2]), C.byref(v), _stream()))
    b.col("intensity").fill_(1.0)
    return b


def _slice_count(n_rays, first, stride):
    return max(0, (n_rays - first + stride - 1) // stride)


def PointSource(S, Axis, Divergence, NbRays, Wavelength=None, device=None, first=0, count=None, stride=1):
    """Rays from the point S filling a cone of half-angle Divergence (rad) about Axis on Vogel's
    spiral (ART/ModuleSource.py:54-81).  `first`/`count`/`stride` select rays first, first+stride, ...
    of the NbRays-ray bundle (one rank's share)."""
    count = _slice_count(NbRays, first, stride) if count is None else count
    return _generate(0, NbRays, first, count, np.tan(Divergence), Axis, S, Wavelength, device, stride)


def ExtendedSource(S, Axis, Diameter, Divergence, NbRays, Wavelength=None, device=None, first=0, count=None,
                   stride=1):
    """An extended source (ART/ModuleSource.py:85-131): max(30, 250*Diameter) -- at most NbRays/300 --
    point sources on a Vogel spiral over a disk of the given Diameter, each emitting the same cone of
    max(300, NbRays / n_sources) rays.  The bundle holds n_sources * rays_per_source rays (not exactly
    NbRays, as in the reference), numbered source by source."""
    n_ps = min(max(30, int(250 * Diameter)), int(NbRays / 300))
    per = max(300, int(NbRays / n_ps))
    n_total = n_ps * per
    count = _slice_count(n_total, first, stride) if count is None else count
    return _generate(2, n_total, first, count, np.tan(Divergence), Axis, S, Wavelength, device, stride,
                     extended=(n_ps, per, Diameter / 2))


def PlaneWaveDisk(Centre, Axis, Radius, NbRays, Wavelength=None, device=None, first=0, count=None, stride=1):
    """Collimated rays from a disk (ART/ModuleSource.py:135-169).  As in the reference the bundle
    holds NbRays-1 rays: points 0 .. NbRays-2 of the NbRays-point spiral."""
    count = _slice_count(NbRays - 1, first, stride) if count is None else count
    if count > 0 and first + (count - 1) * stride > NbRays - 2:
        raise ValueError("PlaneWaveDisk(NbRays) has NbRays-1 rays")
    return _generate(1, NbRays, first, count, Radius, Axis, Centre, Wavelength, device, stride)


def PlaneWaveSquare(Centre, Axis, SideLength, NbRays, Wavelength=None, device=None):
    """Collimated rays on a square grid of int(sqrt(NbRays))^2 points of side SideLength, plus a central ray.

    ART/ModuleSource.py:173-208 is meant to build this bundle but raises for every NbRays >= 4 (it compares whole
    coordinate arrays, `abs(x) > 1e-4`, inside the loop over their entries).  This is the evident intent of that
    code: ray 0 on the axis, then the grid points (i, j) in row order with |i| > 1e-4 and |j| > 1e-4, numbered
    consecutively, rotated from ez onto Axis and moved to Centre.  A small host-side construction (no kernel)."""
    from . import ModuleGeometry as mgeo
    from .ModuleOpticalRay import RayBundle
    m = int(np.sqrt(NbRays))
    grid = np.linspace(-SideLength / 2, SideLength / 2, m)
    keep = grid[np.abs(grid) > 1e-4]
    xx, yy = np.meshgrid(keep, keep, indexing="ij")
    P = np.concatenate([np.zeros((1, 3)), np.stack([xx.ravel(), yy.ravel(), np.zeros(xx.size)], axis=1)])
    U = np.tile(np.array([0.0, 0.0, 1.0]), (P.shape[0], 1))
    bundle = RayBundle.from_numpy(P, U, wavelength=Wavelength, device="cpu")
    bundle = mgeo.TranslationRayList(mgeo.RotationRayList(bundle, np.array([0.0, 0.0, 1.0]), np.asarray(Axis, dtype=np.float64)),
                                     np.asarray(Centre, dtype=np.float64))
    return bundle if device is None else bundle.to(device)


def ApplyGaussianIntensityToRayList(RayList, IntensityFraction=1 / np.e**2, group=None, axis=None, scale=None):
    """Gaussian intensity profile, 1 on the axis falling to IntensityFraction at the edge of the
    bundle (ART/ModuleSource.py:219-261): by angle for a diverging bundle (max angle > 1e-12),
    by distance of the ray origin from the lab origin otherwise.

    RayList: RayBundle on the device (filled in place and returned).  With a torch.distributed
    `group` (or the default group when initialised and group=True) the axis and the extent are
    all-reduced so that every rank normalises by the FULL bundle."""
    b = RayList
    require_cuda(b.device)
    import torch.distributed as dist
    use_dist = group is not None and dist.is_available() and dist.is_initialized()
    grp = None if group is True else group
    if axis is None:
        su = torch.stack([b.col("ux").sum(), b.col("uy").sum(), b.col("uz").sum(),
                          torch.tensor(float(b.n), dtype=torch.float64, device=b.device)])
        if use_dist:
            dist.all_reduce(su, op=dist.ReduceOp.SUM, group=grp)
        su = su.cpu().numpy()
        axis = mgeo.Normalize(su[:3] / su[3])  # FindCentralRay(...).vector, ART/ModuleProcessing.py:464-482
    v = b.view()
    if b.origin is not None:
        v.px = v.py = v.pz = None  # uniform origin: |P| is not needed (a diverging bundle uses angles)
    ext = torch.zeros(2, dtype=torch.float64, device=b.device)
    with torch.cuda.device(b.device):
        _cabi.check(_cabi.lib().art_source_extents(C.byref(v), _cabi.vec3(axis), _ptr(ext), _stream()))
    if use_dist:
        dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=grp)
    max_angle, max_dist = (float(x) for x in ext.cpu().numpy())
    mode = 0 if max_angle > 1e-12 else 1
    if scale is None:
        scale = max_angle if mode == 0 else max_dist
    with torch.cuda.device(b.device):
        _cabi.check(_cabi.lib().art_source_intensity(C.byref(v), _cabi.vec3(axis), mode, float(scale),
                                                     float(IntensityFraction), _stream()))
    return b


def synthetic_source(SourceProperties, first_optic_support=None, device=None, first=0, count=None, group=None,
                     intensity=True, stride=1):
    """The source bundle `OEPlacement` launches (ART/ModuleProcessing.py:55-79): from the origin along
    +x; a plane-wave disk when Divergence == 0 (radius SourceSize/2, or from the first optic's
    support when SourceSize == 0), else a point source; Gaussian intensities down to 1/e^2 at the edge.
    Divergence and SourceSize both non-zero give an ExtendedSource."""
    div = SourceProperties["Divergence"]
    size = SourceProperties["SourceSize"]
    n = int(SourceProperties["NumberRays"])
    wl = SourceProperties.get("Wavelength")
    origin = np.zeros(3)
    axis = np.array([1.0, 0.0, 0.0])
    if div == 0:
        if size == 0:
            sup = first_optic_support
            radius = 0.5 * min(sup.dimX, sup.dimY) if hasattr(sup, "dimX") else sup.radius
        else:
            radius = size / 2
        b = PlaneWaveDisk(origin, axis, radius, n, Wavelength=wl, device=device, first=first, count=count, stride=stride)
    else:
        if size != 0:
            b = ExtendedSource(origin, axis, size, div, n, Wavelength=wl, device=device, first=first, count=count,
                               stride=stride)
        else:
            b = PointSource(origin, axis, div, n, Wavelength=wl, device=device, first=first, count=count, stride=stride)
    if intensity:
        ApplyGaussianIntensityToRayList(b, 1 / np.e**2, group=group)
    return b


def source_descriptor(SourceProperties, first_optic_support=None, first=0, count=None, stride=1, intensity=True):
    """The ArtSourceDesc (include/art_b200.h) of the bundle `synthetic_source` would build: what
    `DeviceChain.run_source` hands to the library instead of rays -- the reference's host input is this
    description too (SourceProperties, ART/ModuleProcessing.py:58-79)."""
    div = SourceProperties["Divergence"]
    size = SourceProperties["SourceSize"]
    n = int(SourceProperties["NumberRays"])
    d = _cabi.ArtSourceDesc()
    d.intensity = 1 if intensity else 0
    d.intensity_fraction = 1 / np.e**2
    d.axis = (1.0, 0.0, 0.0)
    d.origin = (0.0, 0.0, 0.0)
    d.first, d.stride = int(first), int(stride)
    if div == 0:
        if size == 0:
            sup = first_optic_support
            radius = 0.5 * min(sup.dimX, sup.dimY) if hasattr(sup, "dimX") else sup.radius
        else:
            radius = size / 2
        d.kind, d.n_total, d.rho = 1, n, float(radius)
        n_rays = n - 1  # PlaneWaveDisk emits NbRays - 1 rays
    elif size != 0:
        n_ps = min(max(30, int(250 * size)), int(n / 300))
        per = max(300, int(n / n_ps))
        d.kind, d.n_total, d.rho = 2, n_ps * per, float(np.tan(div))
        d.n_point_sources, d.rays_per_source, d.source_radius = n_ps, per, size / 2
        n_rays = n_ps * per
    else:
        d.kind, d.n_total, d.rho = 0, n, float(np.tan(div))
        n_rays = n
    d.count = _slice_count(n_rays, first, stride) if count is None else int(count)
    return d
