"""TEST INFRASTRUCTURE: builds / loads tests/hostcheck/libart_hostcheck.so, which runs the per-ray
device code of csrc/art_device.cuh on the host (see tests/hostcheck/hostcheck.cu)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cu")
OUT = os.path.join(HERE, "hostcheck", "libart_hostcheck.so")
CSRC = os.path.join(os.path.dirname(HERE), "attosecondraytracing_b200", "csrc")


def build():
    # ART_HOSTCHECK_SO: a pre-built variant of the checker (kernel tuning flags) instead of the default build
    if os.environ.get("ART_HOSTCHECK_SO"):
        return os.environ["ART_HOSTCHECK_SO"]
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("art_device.cuh", "art_optics.cuh", "art_lowering.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    subprocess.run([nvcc, "-std=c++17", "-O2", "-Wno-deprecated-gpu-targets", "-shared", "-Xcompiler", "-fPIC",
                    "-o", OUT, SRC], check=True, capture_output=True)
    return OUT


def trace(lowered, P, U, flags):
    """Run the chain of `lowered` (a _lowering.LoweredChain, variant 0) on the host for rays (P, U).
    Returns a list (per element) of dicts alive, P, U, path, inc."""
    L = C.CDLL(build())
    L.hc_last_error.restype = C.c_char_p
    n = P.shape[0]
    ne = lowered.n_elements
    cols = [np.ascontiguousarray(P[:, i]) for i in range(3)] + [np.ascontiguousarray(U[:, i]) for i in range(3)]
    outs = [np.empty(ne * n) for _ in range(8)]
    alive = np.zeros(ne * n, dtype=np.uint8)
    dp = C.POINTER(C.c_double)
    rc = L.hc_trace(lowered.elements, C.c_int(ne), lowered.defects, C.c_int(lowered.n_defects), lowered.gridmaps,
                    C.c_int(lowered.n_gridmaps), C.c_longlong(n),
                    *[c.ctypes.data_as(dp) for c in cols], C.c_uint(flags),
                    *[o.ctypes.data_as(dp) for o in outs], alive.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc != 0:
        raise RuntimeError(L.hc_last_error().decode())
    res = []
    for k in range(ne):
        s = slice(k * n, (k + 1) * n)
        res.append({"alive": alive[s].astype(bool),
                    "P": np.stack([outs[0][s], outs[1][s], outs[2][s]], axis=1),
                    "U": np.stack([outs[3][s], outs[4][s], outs[5][s]], axis=1),
                    "path": outs[6][s], "inc": outs[7][s]})
    return res
