"""Reference-side binding of libart_b200.so: what a maintainer of ART adds to route
`ModuleProcessing.RayTracingCalculation` (ART/ModuleProcessing.py:250) through the B200 library.

Depends on numpy and ctypes only -- no torch, none of this repository's Python package.  It works on
the REFERENCE's own objects by duck typing (`OpticalElement.type/.position/.normal/.majoraxis`, the
mirror / mask / support / Zernike attributes of ART/ModuleMirror.py, ModuleMask.py, ModuleSupport.py,
ModuleDefects.py, and `Ray.point/.vector/.path/.number/...`), flattens the list[Ray] into FP64
columns, calls `art_trace_host`, and rebuilds the list[list[Ray]] the reference returns.

    import art_b200_binding as b200
    b200.load("/path/to/libart_b200.so")
    ART.ModuleProcessing.RayTracingCalculation = b200.make_RayTracingCalculation(ART.ModuleOpticalRay.Ray)
"""
import ctypes as C

import numpy as np

_lib = None
c_dp = C.POINTER(C.c_double)


class ArtElementDesc(C.Structure):
    _fields_ = [("surface", C.c_int32), ("support", C.c_int32), ("surface_params", C.c_double * 4),
                ("support_params", C.c_double * 6), ("centre", C.c_double * 3), ("position", C.c_double * 3),
                ("normal", C.c_double * 3), ("majoraxis", C.c_double * 3), ("n_defects", C.c_int32),
                ("first_defect", C.c_int32), ("n_gridmaps", C.c_int32), ("first_gridmap", C.c_int32)]


class ArtZernikeDesc(C.Structure):
    _fields_ = [("radius", C.c_double), ("n_coefficients", C.c_int32), ("n", C.POINTER(C.c_int32)),
                ("m", C.POINTER(C.c_int32)), ("c", c_dp)]


class ArtBundleView(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("px", "py", "pz", "ux", "uy", "uz", "path", "incidence", "intensity",
                                          "alive")] + [("n", C.c_int64)]


SURFACE = {"Plane Mirror": 0, "SphericalCC Mirror": 1, "SphericalCX Mirror": 1, "Parabolic Mirror": 2,
           "Toroidal Mirror": 3, "Ellipsoidal Mirror": 4, "CylindricalCC Mirror": 5, "CylindricalCX Mirror": 5, "Mask": 6}
TRACE_IGNORE_DEFECTS = 1


def load(path):
    global _lib
    L = C.CDLL(path)
    L.art_last_error.restype = C.c_char_p
    L.art_chain_create.argtypes = [C.POINTER(ArtElementDesc), C.c_int32, C.c_int32, C.POINTER(ArtZernikeDesc), C.c_int32,
                                   C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
    L.art_chain_destroy.argtypes = [C.c_void_p]
    L.art_trace_host.argtypes = [C.c_void_p, C.POINTER(ArtBundleView), C.POINTER(ArtBundleView),
                                 C.POINTER(ArtBundleView), C.c_uint32]
    sizes = (C.c_int32 * 6)()
    L.art_abi_sizes(sizes)
    assert list(sizes)[:3] == [C.sizeof(ArtElementDesc), C.sizeof(ArtZernikeDesc), C.sizeof(ArtBundleView)], "ABI mismatch"
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise RuntimeError("libart_b200: " + _lib.art_last_error().decode())


def _support(s):
    """(ART_SUPP_* kind, params) from a reference support object (ART/ModuleSupport.py)."""
    name = type(s).__name__
    if name == "SupportRound":
        return 0, [s.radius]
    if name == "SupportRoundHole":
        return 1, [s.radius, s.radiushole, s.centerholeX, s.centerholeY]
    if name == "SupportRectangle":
        return 2, [s.dimX, s.dimY]
    if name == "SupportRectangleHole":
        return 3, [s.dimX, s.dimY, s.radiushole, s.centerholeX, s.centerholeY]
    if name == "SupportRectangleRectHole":
        return 4, [s.dimX, s.dimY, s.holeX, s.holeY, s.centerholeX, s.centerholeY]
    raise NameError("unsupported support " + name)


def _surface_params(optic):
    base = getattr(optic, "Mirror", optic)  # DeformedMirror wraps the base mirror
    kind = SURFACE[optic.type]
    if kind in (1, 5):
        return kind, [base.radius]
    if kind == 2:
        return kind, [base.p]
    if kind == 3:
        return kind, [base.majorradius, base.minorradius]
    if kind == 4:
        return kind, [base.a, base.b]
    return kind, []


def lower(optical_elements):
    """ArtElementDesc / ArtZernikeDesc arrays for a list of reference OpticalElement objects."""
    n = len(optical_elements)
    els = (ArtElementDesc * n)()
    zern, keep = [], []
    for k, oe in enumerate(optical_elements):
        optic = oe.type
        if not ("Mirror" in optic.type or optic.type == "Mask"):
            raise NameError("I don`t recognize the type of optical element " + optic.type + ".")
        d = els[k]
        d.surface, sp = _surface_params(optic)
        d.support, ap = _support(optic.support)
        d.surface_params[:len(sp)] = [float(x) for x in sp]
        d.support_params[:len(ap)] = [float(x) for x in ap]
        d.centre[:] = [float(x) for x in optic.get_centre()]
        d.position[:] = [float(x) for x in oe.position]
        d.normal[:] = [float(x) for x in oe.normal]
        d.majoraxis[:] = [float(x) for x in oe.majoraxis]
        defects = getattr(optic, "DeformationList", [])
        d.first_defect, d.n_defects = len(zern), len(defects)
        for z in defects:
            if type(z).__name__ != "Zernike":
                raise NotImplementedError("only ModuleDefects.Zernike defects are supported")
            zern.append(z)
    zd = (ArtZernikeDesc * max(1, len(zern)))()
    for i, z in enumerate(zern):
        keys = list(z.coefficients)
        an = (C.c_int32 * len(keys))(*[int(k[0]) for k in keys])
        am = (C.c_int32 * len(keys))(*[int(k[1]) for k in keys])
        ac = (C.c_double * len(keys))(*[float(z.coefficients[k]) for k in keys])
        keep += [an, am, ac]
        zd[i].radius, zd[i].n_coefficients, zd[i].n, zd[i].m, zd[i].c = float(z.R), len(keys), an, am, ac
    return els, zd, len(zern), keep


def _view(cols, alive, n):
    v = ArtBundleView()
    for name, arr in cols.items():
        setattr(v, name, arr.ctypes.data if arr is not None else None)
    v.alive = alive.ctypes.data if alive is not None else None
    v.n = n
    return v


def trace_columns(P, U, optical_elements, IgnoreDefects=True, path=None):
    """Trace rays given as (n,3) arrays; returns per element a dict of alive, P, U, path, incidence."""
    if _lib is None:
        raise RuntimeError("call load(path_to_libart_b200.so) first")
    n = P.shape[0]
    els, zd, nz, keep = lower(optical_elements)
    chain = C.c_void_p()
    # gridded defects (MeasuredMap / Fourrier) need their maps in device memory: not bound here
    _check(_lib.art_chain_create(els, len(optical_elements), 1, zd, nz, None, 0, C.byref(chain)))
    try:
        names = ("px", "py", "pz", "ux", "uy", "uz")
        U = U / np.linalg.norm(U, axis=1)[:, None]  # the Ray.vector setter normalises
        src = {k: np.ascontiguousarray(a) for k, a in zip(names, list(P.T) + list(U.T))}
        src.update(path=None if path is None else np.ascontiguousarray(path, dtype=np.float64), incidence=None,
                   intensity=None)
        vin = _view(src, None, n)
        K = len(optical_elements)
        outs, views = [], (ArtBundleView * K)()
        for k in range(K):
            cols = {c: np.empty(n) for c in names + ("path", "incidence")}
            cols["intensity"] = None
            alive = np.zeros(n, dtype=np.uint8)
            views[k] = _view(cols, alive, n)
            outs.append((cols, alive))
        _check(_lib.art_trace_host(chain, C.byref(vin), None, views, TRACE_IGNORE_DEFECTS if IgnoreDefects else 0))
    finally:
        _lib.art_chain_destroy(chain)
    res = []
    for cols, alive in outs:
        a = alive.astype(bool)
        res.append({"alive": a, "P": np.stack([cols["px"], cols["py"], cols["pz"]], axis=1),
                    "U": np.stack([cols["ux"], cols["uy"], cols["uz"]], axis=1), "path": cols["path"],
                    "incidence": cols["incidence"]})
    return res


def make_RayTracingCalculation(Ray):
    """A replacement for ModuleProcessing.RayTracingCalculation that returns list[list[Ray]] built with
    the given Ray class (ART.ModuleOpticalRay.Ray)."""

    def RayTracingCalculation(source_rays, optical_elements, IgnoreDefects=True):
        n = len(source_rays)
        P = np.array([r.point for r in source_rays], dtype=np.float64).reshape(n, 3)
        U = np.array([r.vector for r in source_rays], dtype=np.float64).reshape(n, 3)
        path0 = np.array([float(np.sum(r.path)) for r in source_rays])
        res = trace_columns(P, U, optical_elements, IgnoreDefects, path=path0)
        output_rays = []
        for b in res:
            rays = []
            for i in np.nonzero(b["alive"])[0]:
                s = source_rays[i]
                # the reference appends one segment per element; only np.sum(path) is ever consumed
                rays.append(Ray(b["P"][i].copy(), b["U"][i].copy(), Path=(float(b["path"][i]),), Number=s.number,
                                Wavelength=s.wavelength, Incidence=float(b["incidence"][i]), Intensity=s.intensity))
            output_rays.append(rays)
        return output_rays

    return RayTracingCalculation
