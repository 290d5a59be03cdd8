"""ctypes binding of libart_b200.so (include/art_b200.h) -- the only door from Python to the CUDA path.

There is no fallback: if the shared library is missing or cannot be loaded, `lib()` raises.
Build it with `python -m attosecondraytracing_b200.build` (nvcc, sm_100a), which `__graft_entry__.build()`
calls as well.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# ART_B200_LIB points the loader at another build of the same library (kernel tuning variants)
LIB_PATH = os.environ.get("ART_B200_LIB") or os.path.join(HERE, "libart_b200.so")

ART_MAX_ELEMENTS = 16

# surface kinds
SURF_PLANE, SURF_SPHERICAL, SURF_PARABOLIC, SURF_TOROIDAL, SURF_ELLIPSOIDAL, SURF_CYLINDRICAL, SURF_MASK = range(7)
# support kinds
SUPP_ROUND, SUPP_ROUND_HOLE, SUPP_RECT, SUPP_RECT_HOLE, SUPP_RECT_RECT_HOLE = range(5)

# central-sum row
(C_SUX, C_SUY, C_SUZ, C_SPX, C_SPY, C_SPZ, C_SPATH, C_N, C_SW_OUT, C_SW_IN) = range(10)
CENTRAL_LEN = 10
# moments row
(M_N, M_SX, M_SY, M_SXX, M_SYY, M_SD, M_SDD, M_SW, M_SWX, M_SWY, M_SWXX, M_SWYY, M_SWD, M_SWDD,
 M_XMIN, M_XMAX, M_YMIN, M_YMAX, M_DMIN, M_DMAX, M_TMAX) = range(21)
MOMENTS_LEN = 24
SCAN_LEN = 32
(S_N, S_SW, S_X, S_Y, S_AX, S_AY, S_XX, S_YY, S_AXAX, S_AYAY, S_XAX, S_YAY, S_D, S_G, S_DD, S_GG, S_DG) = range(17)
S_WEIGHTED = 17
MOMENT_SUM = list(range(0, 14)) + [21, 22, 23]
MOMENT_MIN = [M_XMIN, M_YMIN, M_DMIN]
MOMENT_MAX = [M_XMAX, M_YMAX, M_DMAX, M_TMAX]

TRACE_IGNORE_DEFECTS = 1
TRACE_NO_INCIDENCE = 2
TRACE_UNIFORM_POINT = 4
TRACE_NO_FOLD = 8

c_double_p = C.POINTER(C.c_double)
c_u8_p = C.POINTER(C.c_uint8)


class ArtElementDesc(C.Structure):
    _fields_ = [
        ("surface", C.c_int32), ("support", C.c_int32),
        ("surface_params", C.c_double * 4), ("support_params", C.c_double * 6),
        ("centre", C.c_double * 3), ("position", C.c_double * 3),
        ("normal", C.c_double * 3), ("majoraxis", C.c_double * 3),
        ("n_defects", C.c_int32), ("first_defect", C.c_int32),
        ("n_gridmaps", C.c_int32), ("first_gridmap", C.c_int32),
    ]


class ArtGridMapDesc(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("x0", C.c_double), ("x1", C.c_double), ("y0", C.c_double),
        ("y1", C.c_double), ("h", C.c_void_p), ("dx", C.c_void_p), ("dy", C.c_void_p),
    ]


class ArtZernikeDesc(C.Structure):
    _fields_ = [
        ("radius", C.c_double), ("n_coefficients", C.c_int32),
        ("n", C.POINTER(C.c_int32)), ("m", C.POINTER(C.c_int32)), ("c", c_double_p),
    ]


class ArtBundleView(C.Structure):
    _fields_ = [
        ("px", C.c_void_p), ("py", C.c_void_p), ("pz", C.c_void_p),
        ("ux", C.c_void_p), ("uy", C.c_void_p), ("uz", C.c_void_p),
        ("path", C.c_void_p), ("incidence", C.c_void_p), ("intensity", C.c_void_p),
        ("alive", C.c_void_p), ("n", C.c_int64),
    ]


class ArtDetector(C.Structure):
    _fields_ = [
        ("centre", C.c_double * 3), ("normal", C.c_double * 3), ("refpoint", C.c_double * 3),
        ("cvec", C.c_double * 3), ("rot", C.c_double * 9), ("l0", C.c_double), ("n_rays", C.c_double),
    ]


DETECTOR_DOUBLES = C.sizeof(ArtDetector) // 8  # 23


class ArtSourceDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("intensity", C.c_int32), ("n_total", C.c_int64), ("first", C.c_int64),
        ("count", C.c_int64), ("stride", C.c_int64), ("rho", C.c_double), ("axis", C.c_double * 3),
        ("origin", C.c_double * 3), ("n_point_sources", C.c_int64), ("rays_per_source", C.c_int64),
        ("source_radius", C.c_double), ("intensity_fraction", C.c_double),
    ]


E_PEER_TIMEOUT = -5  # ART_E_PEER_TIMEOUT


HIST_FIXED_ONE = 67108864.0  # ART_HIST_FIXED_ONE


PEER_MAX_RANKS = 16       # ART_PEER_MAX_RANKS
PEER_MAX_VARIANTS = 64    # ART_PEER_MAX_VARIANTS
PEER_STATS = 4            # ART_PEER_STATS


def peer_buffer_bytes(world):
    """ART_PEER_BUFFER_BYTES of the header."""
    return 16 * (2 * world * PEER_MAX_VARIANTS * MOMENTS_LEN) + 8 * (world + 2 + PEER_STATS)


def hist_len(nx, ny, nt):
    """ART_HIST_LEN of the header."""
    return 3 * nx * ny + 2 * nt


class ArtError(RuntimeError):
    """A libart_b200 call returned a negative status."""

    def __init__(self, code, message):
        super().__init__(f"libart_b200 error {code}: {message}")
        self.code = code


# every exported symbol of include/art_b200.h with its signature
_SIGNATURES = {
    "art_version": (C.c_int32, []),
    "art_last_error": (C.c_char_p, []),
    "art_abi_sizes": (C.c_int32, [C.POINTER(C.c_int32)]),
    "art_device_count": (C.c_int32, [C.POINTER(C.c_int32)]),
    "art_element_rotation": (C.c_int32, [c_double_p, c_double_p, c_double_p]),
    "art_chain_create": (C.c_int32, [C.POINTER(ArtElementDesc), C.c_int32, C.c_int32, C.POINTER(ArtZernikeDesc),
                                     C.c_int32, C.POINTER(ArtGridMapDesc), C.c_int32, C.POINTER(C.c_void_p)]),
    "art_chain_destroy": (C.c_int32, [C.c_void_p]),
    "art_trace": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(ArtBundleView), C.POINTER(ArtBundleView),
                              C.POINTER(ArtBundleView), C.c_uint32, C.c_void_p, C.c_void_p]),
    "art_trace_detect": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(ArtBundleView),
                                     C.POINTER(ArtBundleView), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "art_detector_autoplace": (C.c_int32, [C.c_void_p, C.c_double, C.c_int32, C.c_void_p, C.c_void_p]),
    "art_detector_make": (C.c_int32, [c_double_p, c_double_p, c_double_p, C.c_double, C.POINTER(ArtDetector)]),
    "art_detector_moments": (C.c_int32, [C.c_void_p, C.POINTER(ArtBundleView), C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "art_detector_scan_moments": (C.c_int32, [C.c_void_p, C.POINTER(ArtBundleView), C.c_int32, C.c_void_p, C.c_void_p,
                                              C.c_void_p]),
    "art_detector_histogram": (C.c_int32, [C.POINTER(ArtBundleView), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_double, C.c_void_p, C.c_void_p]),
    "art_peer_exchange": (C.c_int32, [C.POINTER(C.c_uint64), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_double, C.c_void_p, C.c_void_p]),
    "art_peer_exchange_fold": (C.c_int32, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int32, C.c_int32, C.c_int32,
                                           C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "art_peer_status": (C.c_int32, [C.POINTER(C.c_uint64), C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.c_void_p]),
    "art_peer_stats": (C.c_int32, [C.POINTER(C.c_uint64), C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.c_int32,
                                   C.c_void_p]),
    "art_moments_merge": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "art_sweep": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(ArtBundleView), C.c_uint32, C.c_double,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "art_delays": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "art_source_generate": (C.c_int32, [C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_double, c_double_p,
                                        c_double_p, C.c_int64, C.c_int64, C.c_double, C.POINTER(ArtBundleView),
                                        C.c_void_p]),
    "art_source_extents": (C.c_int32, [C.POINTER(ArtBundleView), c_double_p, C.c_void_p, C.c_void_p]),
    "art_source_intensity": (C.c_int32, [C.POINTER(ArtBundleView), c_double_p, C.c_int32, C.c_double, C.c_double,
                                         C.c_void_p]),
    "art_run_host": (C.c_int32, [C.c_void_p, C.POINTER(ArtBundleView), C.POINTER(ArtBundleView), C.c_uint32,
                                 C.c_double, C.POINTER(ArtDetector), c_double_p, c_double_p,
                                 C.POINTER(ArtDetector)]),
    "art_run_host_sharded": (C.c_int32, [C.c_void_p, C.POINTER(ArtBundleView), C.POINTER(ArtBundleView), C.c_uint32,
                                         C.c_double, C.POINTER(ArtDetector), c_double_p, c_double_p,
                                         C.POINTER(ArtDetector), C.POINTER(C.c_uint64), C.c_int32, C.c_int32]),
    "art_run_source_host": (C.c_int32, [C.c_void_p, C.POINTER(ArtSourceDesc), C.c_uint32, C.c_double,
                                        C.POINTER(ArtDetector), c_double_p, c_double_p, C.POINTER(ArtDetector),
                                        C.POINTER(C.c_uint64), C.c_int32, C.c_int32]),
    "art_trace_host": (C.c_int32, [C.c_void_p, C.POINTER(ArtBundleView), C.POINTER(ArtBundleView),
                                   C.POINTER(ArtBundleView), C.c_uint32]),
    "art_probe_fp64": (C.c_int32, [c_double_p]),
    "art_probe_hbm": (C.c_int32, [c_double_p]),
    "art_launch_count": (C.c_int64, []),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib():
    """The loaded library (loads on first use).  Raises if libart_b200.so is absent: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing -- build the CUDA library first (python -m attosecondraytracing_b200.build); "
                "this package has no CPU path")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        sizes = (C.c_int32 * 6)()
        L.art_abi_sizes(sizes)
        mine = [C.sizeof(t) for t in (ArtElementDesc, ArtZernikeDesc, ArtBundleView, ArtDetector, ArtGridMapDesc,
                                      ArtSourceDesc)]
        if list(sizes) != mine:
            raise RuntimeError(f"struct layout mismatch between _cabi.py {mine} and libart_b200.so {list(sizes)}")
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise ArtError(rc, lib().art_last_error().decode("utf-8", "replace"))


def vec3(v):
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


def element_rotation(normal, majoraxis):
    """Row-major lab->element rotation (host arithmetic inside the library), as a nested list."""
    out = (C.c_double * 9)()
    check(lib().art_element_rotation(vec3(normal), vec3(majoraxis), out))
    return [[out[3 * i + j] for j in range(3)] for i in range(3)]
