"""
ORACLE (extended precision) -- TEST INFRASTRUCTURE ONLY.  Not part of the product.

An `np.longdouble` (x87 80-bit, eps = 1.1e-19) evaluation of the same path as oracle/art_oracle.py, used
ONLY as an arbiter: on the 5 m-arm telescope (BASELINE config 5) the reference's own float64 rounding
noise -- directions formed as R(p+u) - R(p) with |p| ~ 5000 mm, ART/ModuleGeometry.py:357-368 -- is
5e-9 .. 1.2e-8 mm in the intersection points, above the 1e-9 mm bar of the north star (SURVEY.md
Appendix C.1).  Comparing both the CUDA path and the reference with this evaluation shows which side the
noise is on: the tests hold the CUDA path to 1e-9 mm against THIS oracle where they can hold it only to
3e-8 mm against the reference.

What is evaluated in extended precision: the element-frame matrices (from the float64 poses the reference
produced), frame changes, the surface intersections (closed form for the quadrics; the reference's quartic
for the toroid, its float64 np.roots candidates polished by Newton in longdouble), normals, Zernike
offsets / normals (Andersen's recurrences, ART/recursive_zernike_generator.py), reflection, path
accumulation, detector placement, in-plane points and delays.  Which rays survive is NOT re-decided
here beyond the reference's own rule applied to the extended-precision numbers; callers compare on the rays
both sides keep.

Every function cites the reference file:line it follows (relative to /root/reference).
"""
from __future__ import annotations

import numpy as np

import art_oracle as orc

LD = np.longdouble
LIGHTSPEED = LD(299792458000)  # mm/s, ART/ModuleDetector.py:21


def available():
    """True when np.longdouble really is wider than float64 (x86: 64-bit mantissa)."""
    return np.finfo(LD).eps < 1e-18


def _ld(a):
    return np.asarray(a, dtype=LD)


def norm(v):
    v = _ld(v)
    return np.sqrt(np.sum(v * v, axis=-1))


def normalize(v):
    v = _ld(v)
    return v / norm(v)[..., None]


def angle_between(U, V):
    """Kahan's formula, ART/ModuleGeometry.py:40-44."""
    U, V = _ld(U), _ld(V)
    u, v = norm(U)[..., None], norm(V)[..., None]
    return 2 * np.arctan2(norm(U * v - V * u), norm(U * v + V * u))


def rotation_matrix(axis1, axis2):
    """RotationPoint as a matrix, ART/ModuleGeometry.py:333-343, 321-329 (same 1e-10 branches)."""
    a1, a2 = _ld(axis1), _ld(axis2)
    ang = angle_between(a1, a2)
    if abs(ang) < 1e-10:
        return np.eye(3, dtype=LD)
    if abs(ang - LD(np.pi)) < 1e-10:
        return -np.eye(3, dtype=LD)
    k = normalize(np.cross(a1, a2))
    c, s = np.cos(ang), np.sin(ang)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]], dtype=LD)
    return c * np.eye(3, dtype=LD) + s * K + (1 - c) * np.outer(k, k)


def element_frame_matrix(normal, majoraxis):
    """ART/ModuleProcessing.py:289-294."""
    R1 = rotation_matrix(normal, orc.EZ)
    R2 = rotation_matrix(R1 @ _ld(majoraxis), orc.EX)
    return R2 @ R1


def optic_centre(optic):
    """get_centre(): the float64 value the reference computed (ART/ModuleMirror.py:89, 185, 357, 500, 695, 851)
    -- it enters the reference's arithmetic as a float64 constant, so it does here."""
    return _ld(orc.optic_centre(optic))


def optic_normal(optic, P):
    """get_normal(P), ART/ModuleMirror.py:84, 180, 349, 480-498, 685, 846."""
    k = optic["kind"]
    P = _ld(P)
    if k in ("plane", "mask"):
        return np.broadcast_to(_ld(orc.EZ), P.shape).copy()
    if k == "spherical":
        return normalize(-P)
    if k == "parabolic":
        return normalize(np.stack([-P[..., 0], -P[..., 1], np.full(P.shape[:-1], LD(optic["p"]))], axis=-1))
    if k == "toroidal":
        x, y, z = P[..., 0], P[..., 1], P[..., 2]
        R, r = LD(optic["majorradius"]), LD(optic["minorradius"])
        A = R**2 - r**2
        gx = 4 * (x**3 + x * y**2 + x * z**2 + x * A) - 8 * x * R**2
        gy = 4 * (y**3 + y * x**2 + y * z**2 + y * A)
        gz = 4 * (z**3 + z * x**2 + z * y**2 + z * A) - 8 * z * R**2
        return normalize(-np.stack([gx, gy, gz], axis=-1))
    if k == "ellipsoidal":
        a, b = LD(optic["a"]), LD(optic["b"])
        return normalize(np.stack([-P[..., 0] / a**2, -P[..., 1] / b**2, -P[..., 2] / b**2], axis=-1))
    if k == "cylindrical":
        return normalize(np.stack([np.zeros(P.shape[:-1], dtype=LD), -P[..., 1], -P[..., 2]], axis=-1))
    raise ValueError(k)


def _quadratic_roots(a, b, c):
    """Both real roots of a t^2 + b t + c in the cancellation-free form; NaN when
    complex.  a == 0 degrades to the single root -c/b."""
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        disc = b * b - 4 * a * c
        sq = np.sqrt(np.where(disc >= 0, disc, LD(np.nan)))
        q = -(b + np.where(b >= 0, sq, -sq)) / 2
        t1 = q / a
        t2 = c / q
    return np.stack([t1, t2], axis=1)


def _select(P, U, roots, side, sup):
    """Candidate rule, ART/ModuleGeometry.py:110-147 + ART/ModuleMirror.py:27-38 (as oracle._select_hit)."""
    with np.errstate(invalid="ignore", over="ignore"):
        valid = np.isfinite(roots) & (roots > 1e-12)
        t = np.where(valid, roots, LD(0))
        pts = P[:, None, :] + t[..., None] * U[:, None, :]
        cand = valid & side(pts) & sup(pts)
        count = cand.sum(axis=1)
        tt = np.where(cand, t, LD(np.inf))
        best = np.min(tt, axis=1)
    hit = (count == 1) | (count == 2)
    return hit, np.where(hit, best, LD(np.nan))


def optic_intersection(optic, P, U):
    """_get_intersection of each optic class in extended precision.  Returns (hit, t)."""
    k = optic["kind"]
    sup = optic["support"]
    x, y, z = P[:, 0], P[:, 1], P[:, 2]
    ux, uy, uz = U[:, 0], U[:, 1], U[:, 2]

    def sup_xy(q):
        return orc.support_include(sup, q[..., 0], q[..., 1])

    if k in ("plane", "mask"):  # ART/ModuleMirror.py:73-82, ART/ModuleMask.py:51-61
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            t = -z / uz
            I = U * t[:, None] + P
            inc = orc.support_include(sup, I[:, 0], I[:, 1])
            hit = (t > 0) & (~inc if k == "mask" else inc)
        return hit, np.where(hit, t, LD(np.nan))
    if k == "spherical":  # :163-178
        roots = _quadratic_roots(np.sum(U * U, axis=1), 2 * np.sum(U * P, axis=1),
                                 np.sum(P * P, axis=1) - LD(optic["radius"]) ** 2)
        return _select(P, U, roots, lambda q: q[..., 2] < 0, sup_xy)
    if k == "parabolic":  # :325-347
        p = LD(optic["p"])
        roots = _quadratic_roots(ux**2 + uy**2, 2 * (ux * x + uy * y) - 2 * p * uz, x**2 + y**2 - 2 * p * z)
        C = optic_centre(optic)
        return _select(P, U, roots, lambda q: np.ones(q.shape[:-1], bool),
                       lambda q: orc.support_include(sup, q[..., 0] - C[0], q[..., 1] - C[1]))
    if k == "ellipsoidal":  # :662-683
        a_, b_ = LD(optic["a"]), LD(optic["b"])
        roots = _quadratic_roots((uy**2 + uz**2) / b_**2 + (ux / a_) ** 2,
                                 2 * ((uy * y + uz * z) / b_**2 + (ux * x) / a_**2),
                                 (y**2 + z**2) / b_**2 + (x / a_) ** 2 - 1)
        C = optic_centre(optic)
        return _select(P, U, roots, lambda q: q[..., 2] < 0,
                       lambda q: orc.support_include(sup, q[..., 0] - C[0], q[..., 1] - C[1]))
    if k == "cylindrical":  # :824-844
        roots = _quadratic_roots(uy**2 + uz**2, 2 * (uy * y + uz * z), y**2 + z**2 - LD(optic["radius"]) ** 2)
        return _select(P, U, roots, lambda q: q[..., 2] < 0, sup_xy)
    if k == "toroidal":  # :443-478: the reference's quartic; float64 np.roots candidates polished in longdouble
        R, r = LD(optic["majorradius"]), LD(optic["minorradius"])
        G = 4 * R**2 * (ux**2 + uz**2)
        H = 8 * R**2 * (ux * x + uz * z)
        I = 4 * R**2 * (x**2 + z**2)
        J = np.sum(U * U, axis=1)
        K = 2 * np.sum(U * P, axis=1)
        L = np.sum(P * P, axis=1) + R**2 - r**2
        co = [J**2, 2 * J * K, 2 * J * L + K**2 - G, 2 * K * L - H, L**2 - I]
        roots = _ld(orc._roots_rows(np.stack([np.asarray(c, dtype=np.float64) for c in co], axis=1)))
        for _ in range(4):
            f = (((co[0][:, None] * roots + co[1][:, None]) * roots + co[2][:, None]) * roots + co[3][:, None]) * roots \
                + co[4][:, None]
            df = ((4 * co[0][:, None] * roots + 3 * co[1][:, None]) * roots + 2 * co[2][:, None]) * roots + co[3][:, None]
            with np.errstate(invalid="ignore", divide="ignore"):
                step = f / df
            roots = np.where(np.isfinite(step), roots - step, roots)
        return _select(P, U, roots, lambda q: q[..., 2] < -R, sup_xy)
    raise ValueError(k)


def _zernike(defect, Q, want):
    R = LD(defect["R"])
    xy = _ld(Q) / R
    Z, GX, GY = orc.zernike_gradient(xy[:, 0], xy[:, 1], defect["max_order"], dtype=LD)
    if want == "offset":
        out = np.zeros(Q.shape[0], dtype=LD)
        for k, c in defect["coefficients"].items():
            out = out + LD(c) * Z[k]
        return out
    dX = np.zeros(Q.shape[0], dtype=LD)
    dY = np.zeros(Q.shape[0], dtype=LD)
    for k, c in defect["coefficients"].items():
        dX = dX + LD(c) * GX[k]
        dY = dY + LD(c) * GY[k]
    return np.stack([-dX / R, -dY / R, np.ones_like(dX)], axis=-1)


def normal_add(N1, N2):
    """ART/ModuleGeometry.py:394-407."""
    n1, n2 = normalize(N1), normalize(N2)
    gX = (-n1[..., 0] / n1[..., 2]) + (-n2[..., 0] / n2[..., 2])
    gY = (-n1[..., 1] / n1[..., 2]) + (-n2[..., 1] / n2[..., 2])
    return np.stack([-gX, -gY, np.ones_like(gX)], axis=-1)


def trace_chain(P, U, elements, ignore_defects=True, numbers=None):
    """RayTracingCalculation, ART/ModuleProcessing.py:250-313, in extended precision; the return layout of
    art_oracle.trace_chain with longdouble arrays.  Zernike defects only (no gridded maps)."""
    P = _ld(P)
    U = normalize(U)
    n = P.shape[0]
    index = np.arange(n)
    numbers = np.arange(n) if numbers is None else np.asarray(numbers)
    path = np.zeros(n, dtype=LD)
    out = []
    for el in elements:
        optic = el["optic"]
        pos = _ld(el["position"])
        R = element_frame_matrix(el["normal"], el["majoraxis"])
        C = optic_centre(optic)
        p1 = (P - pos) @ R.T + C
        u1 = normalize(U @ R.T)
        hit, t = optic_intersection(optic, p1, u1)
        keep = np.nonzero(hit)[0]
        p1, u1, t = p1[keep], u1[keep], t[keep]
        index, path = index[keep], path[keep]
        Ph = u1 * t[:, None] + p1
        defects = optic.get("defects") or []
        if optic["kind"] == "mask":  # ART/ModuleMask.py:93-108
            incidence = angle_between(u1, np.broadcast_to(_ld(orc.EZ), u1.shape))
            u2 = u1
        else:
            if defects:  # ART/ModuleMirror.py:969-980
                h = np.zeros(Ph.shape[0], dtype=LD)
                for D in defects:
                    h = h + _zernike(D, Ph - C, "offset")
                alpha = angle_between(-u1, optic_normal(optic, Ph))
                Ph = Ph - u1 * (h / np.cos(alpha))[:, None]
            nrm = optic_normal(optic, Ph)
            if defects and not ignore_defects:  # ART/ModuleMirror.py:952-961
                for D in defects:
                    nrm = normal_add(nrm, _zernike(D, Ph - C, "normal"))
                    nrm = nrm / norm(nrm)[:, None]
            u2 = normalize(u1 - 2 * np.sum(nrm * u1, axis=1)[:, None] * nrm)  # ART/ModuleMirror.py:878-906
            incidence = angle_between(-u1, nrm)
        path = path + norm(Ph - p1)
        P = (Ph - C) @ R + pos
        U = normalize(u2 @ R)
        out.append({"index": index.copy(), "number": numbers[index], "P": P.copy(), "U": U.copy(),
                    "path": path.copy(), "incidence": incidence})
    return out


def detector_autoplace(P, U, distance):
    """Detector.autoplace + FindCentralRay, ART/ModuleDetector.py:109-137, ART/ModuleProcessing.py:464-482."""
    cp, cv = np.mean(_ld(P), axis=0), normalize(np.mean(_ld(U), axis=0))
    normal = normalize(-cv)
    return {"centre": cp - normal * LD(distance), "normal": normal, "refpoint": cp}


def detector_response(det, P, U, path):
    """(centred in-plane points, delays in fs): get_PointList2DCentre and get_Delays,
    ART/ModuleDetector.py:191-279 + CentrePointList ART/ModuleGeometry.py:222-245."""
    P, U, path = _ld(P), _ld(U), _ld(path)
    n, c = _ld(det["normal"]), _ld(det["centre"])
    t = ((c - P) @ n) / (U @ n)
    H = U * t[:, None] + P
    xy = ((H - c) @ rotation_matrix(n, orc.EZ).T)[:, :2]
    xy = xy - (np.amax(xy, axis=0) + np.amin(xy, axis=0)) / 2
    L = norm(P - H) + path
    return xy, (L - np.mean(L)) / LIGHTSPEED * LD(1e15)
