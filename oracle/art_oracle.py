"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product.

A CPU (numpy, float64) restatement of the reference's hot path: ART v0.93
`ModuleProcessing.RayTracingCalculation` and everything it calls, the detector response and
the bundle statistics.  It exists so that the CUDA path can be checked ray by ray; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import it.  `attosecondraytracing_b200/` never does.

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the reference ITSELF, generated in the build container by
`oracle/gen_golden.py` (which imports /root/reference through `oracle/refshim`) and committed
under `tests/golden/*.npz`.  `tests/test_oracle_vs_golden.py` checks every function here
against those fixtures.

Every function cites the reference file:line (relative to /root/reference) it restates.
Rays are vectorised over the leading axis; the per-ray semantics are the reference's.
"""
from __future__ import annotations

import math

import numpy as np

LIGHTSPEED = 299792458000  # mm/s, ART/ModuleDetector.py:21

EZ = np.array([0.0, 0.0, 1.0])
EX = np.array([1.0, 0.0, 0.0])


# --------------------------------------------------------------------------------------
# geometry primitives (ART/ModuleGeometry.py)
# --------------------------------------------------------------------------------------
def norm(v):
    """np.linalg.norm over the last axis."""
    v = np.asarray(v, dtype=np.float64)
    return np.sqrt(np.sum(v * v, axis=-1))


def normalize(v):
    """ART/ModuleGeometry.py:17-19."""
    v = np.asarray(v, dtype=np.float64)
    return v / norm(v)[..., None]


def angle_between(U, V):
    """Kahan's angle formula, ART/ModuleGeometry.py:40-44 (vectorised over the last axis)."""
    U = np.asarray(U, dtype=np.float64)
    V = np.asarray(V, dtype=np.float64)
    u = norm(U)[..., None]
    v = norm(V)[..., None]
    return 2 * np.arctan2(norm(U * v - V * u), norm(U * v + V * u))


def rotation_about_axis_matrix(axis, angle):
    """Matrix of RotationAroundAxis, ART/ModuleGeometry.py:321-329.

    The reference builds q = exp(angle/2 * axis/|axis|) and returns (q v conj(q)).imag, i.e. a
    right-handed rotation by `angle` about `axis`; this is Rodrigues' matrix of that rotation.
    """
    k = normalize(np.asarray(axis, dtype=np.float64))
    c, s = math.cos(angle), math.sin(angle)
    K = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return c * np.eye(3) + s * K + (1.0 - c) * np.outer(k, k)


def rotation_matrix(axis1, axis2):
    """Matrix M with RotationPoint(P, axis1, axis2) == M @ P, ART/ModuleGeometry.py:333-343.

    angle < 1e-10 -> identity; |angle - pi| < 1e-10 -> MINUS identity (a point inversion, as
    the reference does); otherwise a rotation by the Kahan angle about cross(axis1, axis2).
    """
    a1 = np.asarray(axis1, dtype=np.float64)
    a2 = np.asarray(axis2, dtype=np.float64)
    ang = float(angle_between(a1, a2))
    if abs(ang) < 1e-10:
        return np.eye(3)
    if abs(ang - np.pi) < 1e-10:
        return -np.eye(3)
    return rotation_about_axis_matrix(np.cross(a1, a2), ang)


def element_frame_matrix(normal, majoraxis):
    """Lab -> element rotation of RayTracingCalculation, ART/ModuleProcessing.py:289-294.

    R1 takes normal -> ez; mPrime = R1 majoraxis; R2 takes mPrime -> ex; R = R2 R1.
    The way back (:306-309) is the transpose.
    """
    R1 = rotation_matrix(normal, EZ)
    mprime = R1 @ np.asarray(majoraxis, dtype=np.float64)
    R2 = rotation_matrix(mprime, EX)
    return R2 @ R1


# --------------------------------------------------------------------------------------
# supports (ART/ModuleSupport.py)
# --------------------------------------------------------------------------------------
def include_rectangle(X, Y, x, y):
    """ART/ModuleGeometry.py:249-255 (inclusive)."""
    return (np.abs(x) <= abs(X / 2)) & (np.abs(y) <= abs(Y / 2))


def include_disk(R, x, y):
    """ART/ModuleGeometry.py:259-268 (inclusive)."""
    return (x**2 + y**2) <= R**2


def support_include(support, x, y):
    """`_IncludeSupport` of the five support classes, ART/ModuleSupport.py:68,151,228,322,431.

    `support` = (kind, params...) with kind in
      "round" (R) | "roundhole" (R, Rh, cx, cy) | "rect" (X, Y) |
      "recthole" (X, Y, Rh, cx, cy) | "rectrecthole" (X, Y, hX, hY, cx, cy)
    """
    kind = support[0]
    p = support[1:]
    if kind == "round":
        return include_disk(p[0], x, y)
    if kind == "roundhole":
        return include_disk(p[0], x, y) & ~include_disk(p[1], x - p[2], y - p[3])
    if kind == "rect":
        return include_rectangle(p[0], p[1], x, y)
    if kind == "recthole":
        return include_rectangle(p[0], p[1], x, y) & ~include_disk(p[2], x - p[3], y - p[4])
    if kind == "rectrecthole":
        return include_rectangle(p[0], p[1], x, y) & ~include_rectangle(p[2], p[3], x - p[4], y - p[5])
    raise ValueError(f"unknown support kind {kind!r}")


def support_circum_circ(support):
    """`_CircumCirc`, ART/ModuleSupport.py:95-96,182-183,256-257,353-354,462-463."""
    kind = support[0]
    if kind in ("round", "roundhole"):
        return support[1]
    return np.sqrt(support[1] ** 2 + support[2] ** 2) / 2


# --------------------------------------------------------------------------------------
# polynomial roots exactly as the reference gets them (np.roots + filters)
# --------------------------------------------------------------------------------------
def _roots_rows(coeffs):
    """Row-wise np.roots with the reference's 'real root' filter.

    ART/ModuleGeometry.py:80-106: np.roots(...) then keep roots with |imag| < 1e-15.
    Returns (N, deg) float array, NaN where the root is complex or absent.

    np.roots strips leading/trailing zero coefficients and takes the eigenvalues of the
    companion matrix; for the generic rows (no zero at either end) the companion matrices are
    stacked and handed to np.linalg.eigvals (the same LAPACK geev per matrix); the rare
    degenerate rows go through np.roots itself.
    """
    coeffs = np.asarray(coeffs, dtype=np.float64)
    n, m = coeffs.shape
    deg = m - 1
    out = np.full((n, deg), np.nan)
    finite = np.all(np.isfinite(coeffs), axis=1)
    generic = finite & (coeffs[:, 0] != 0) & (coeffs[:, -1] != 0)
    idx = np.nonzero(generic)[0]
    if idx.size:
        A = np.zeros((idx.size, deg, deg))
        A[:, 0, :] = -coeffs[idx, 1:] / coeffs[idx, :1]
        for k in range(deg - 1):
            A[:, k + 1, k] = 1.0
        ev = np.linalg.eigvals(A)
        real = np.abs(ev.imag) < 1e-15
        out[idx] = np.where(real, ev.real, np.nan)
    for i in np.nonzero(finite & ~generic)[0]:
        r = np.roots(coeffs[i])
        r = np.array([z.real for z in r if abs(z.imag) < 1e-15])
        out[i, : r.size] = r
    return out


def _select_hit(P, U, roots, side_test, support_test):
    """Candidate rule of every curved mirror + `_IntersectionRayMirror`.

    ART/ModuleGeometry.py:110-120 (t > 1e-12), the per-surface loop e.g.
    ART/ModuleMirror.py:172-178, and ART/ModuleMirror.py:27-38 + ClosestPoint
    ART/ModuleGeometry.py:138-147: exactly one candidate -> it; exactly two -> the one nearer to
    the ray origin (the second one on a tie); none or three and more -> miss.
    Returns (hit mask, t).
    """
    with np.errstate(invalid="ignore", over="ignore"):
        valid = np.isfinite(roots) & (roots > 1e-12)
        t = np.where(valid, roots, 0.0)
        pts = P[:, None, :] + t[..., None] * U[:, None, :]
        cand = valid & side_test(pts) & support_test(pts)
        count = cand.sum(axis=1)
        d = pts - P[:, None, :]
        dist2 = np.where(cand, np.sum(d * d, axis=-1), np.inf)
        # first strict minimum loses a tie to the later one: scan from the back
        nr = roots.shape[1]
        pick = nr - 1 - np.argmin(dist2[:, ::-1], axis=1)
        hit = (count == 1) | (count == 2)
        tt = np.take_along_axis(t, pick[:, None], axis=1)[:, 0]
    return hit, np.where(hit, tt, np.nan)


# --------------------------------------------------------------------------------------
# surfaces (ART/ModuleMirror.py, ART/ModuleMask.py)
# --------------------------------------------------------------------------------------
def optic_centre(optic):
    """`get_centre()` of each optic class (element-frame coordinates of the support centre)."""
    k = optic["kind"]
    if k in ("plane", "mask"):  # ART/ModuleMirror.py:89-91, ART/ModuleMask.py:68-70
        return np.zeros(3)
    if k in ("spherical", "cylindrical"):  # ART/ModuleMirror.py:185-187, :851-853
        return np.array([0.0, 0.0, -optic["radius"]])
    if k == "parabolic":  # ART/ModuleMirror.py:357-365
        f, a, p = optic["feff"], optic["offaxisangle"], optic["p"]
        return np.array([f * np.sin(a), 0.0, p * 0.5 - f * np.cos(a)])
    if k == "toroidal":  # ART/ModuleMirror.py:500-502
        return np.array([0.0, 0.0, -optic["majorradius"] - optic["minorradius"]])
    if k == "ellipsoidal":  # ART/ModuleMirror.py:695-714
        a_, b_, oa = optic["a"], optic["b"], optic["offaxisangle"]
        foci = 2 * np.sqrt(a_**2 - b_**2)
        h = -foci / 2 / np.tan(oa)
        R = np.sqrt(foci**2 / 4 + h**2)
        sign = 1
        if math.isclose(oa, np.pi / 2):
            h = 0
        elif oa > np.pi / 2:
            h = -h
            sign = -1
        a = 1 - a_**2 / b_**2
        b = -2 * h
        c = a_**2 + h**2 - R**2
        z = (-b + sign * np.sqrt(b**2 - 4 * a * c)) / (2 * a)
        if math.isclose(z**2, b_**2):
            return np.array([0.0, 0.0, -b_])
        x = a_ * np.sqrt(1 - z**2 / b_**2)
        return np.array([x, 0.0, sign * z])
    raise ValueError(k)


def optic_normal(optic, P):
    """`get_normal(P)` of each optic class; P is (N,3) in the element frame."""
    k = optic["kind"]
    P = np.asarray(P, dtype=np.float64)
    if k in ("plane", "mask"):  # ART/ModuleMirror.py:84-87
        return np.broadcast_to(EZ, P.shape).copy()
    if k == "spherical":  # ART/ModuleMirror.py:180-183
        return normalize(-P)
    if k == "parabolic":  # ART/ModuleMirror.py:349-355
        G = np.stack([-P[..., 0], -P[..., 1], np.full(P.shape[:-1], optic["p"])], axis=-1)
        return normalize(G)
    if k == "toroidal":  # ART/ModuleMirror.py:480-498
        x, y, z = P[..., 0], P[..., 1], P[..., 2]
        R, r = optic["majorradius"], optic["minorradius"]
        A = R**2 - r**2
        gx = 4 * (x**3 + x * y**2 + x * z**2 + x * A) - 8 * x * R**2
        gy = 4 * (y**3 + y * x**2 + y * z**2 + y * A)
        gz = 4 * (z**3 + z * x**2 + z * y**2 + z * A) - 8 * z * R**2
        return normalize(-np.stack([gx, gy, gz], axis=-1))
    if k == "ellipsoidal":  # ART/ModuleMirror.py:685-693
        a, b = optic["a"], optic["b"]
        G = np.stack([-P[..., 0] / a**2, -P[..., 1] / b**2, -P[..., 2] / b**2], axis=-1)
        return normalize(G)
    if k == "cylindrical":  # ART/ModuleMirror.py:846-849
        G = np.stack([np.zeros(P.shape[:-1]), -P[..., 1], -P[..., 2]], axis=-1)
        return normalize(G)
    raise ValueError(k)


def optic_intersection(optic, P, U):
    """`_get_intersection` of each optic class, vectorised.  Returns (hit mask, t)."""
    k = optic["kind"]
    sup = optic["support"]
    x, y, z = P[:, 0], P[:, 1], P[:, 2]
    ux, uy, uz = U[:, 0], U[:, 1], U[:, 2]

    def sup_xy(pts):
        return support_include(sup, pts[..., 0], pts[..., 1])

    if k in ("plane", "mask"):
        # ART/ModuleMirror.py:73-82 and ART/ModuleMask.py:51-61: t>0 with no epsilon;
        # the mask transmits iff the point is NOT on its support.
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            t = -z / uz
            I = U * t[:, None] + P
            inc = support_include(sup, I[:, 0], I[:, 1])
            hit = (t > 0) & (~inc if k == "mask" else inc)
        return hit, np.where(hit, t, np.nan)

    if k == "spherical":  # ART/ModuleMirror.py:163-178
        a = np.sum(U * U, axis=1)
        b = 2 * np.sum(U * P, axis=1)
        c = np.sum(P * P, axis=1) - optic["radius"] ** 2
        roots = _roots_rows(np.stack([a, b, c], axis=1))
        return _select_hit(P, U, roots, lambda q: q[..., 2] < 0, sup_xy)

    if k == "parabolic":  # ART/ModuleMirror.py:325-347 (no z test; support about the centre)
        p = optic["p"]
        da = ux**2 + uy**2
        db = 2 * (ux * x + uy * y) - 2 * p * uz
        dc = x**2 + y**2 - 2 * p * z
        roots = _roots_rows(np.stack([da, db, dc], axis=1))
        C = optic_centre(optic)
        return _select_hit(
            P, U, roots, lambda q: np.ones(q.shape[:-1], bool),
            lambda q: support_include(sup, q[..., 0] - C[0], q[..., 1] - C[1]),
        )

    if k == "toroidal":  # ART/ModuleMirror.py:443-478
        R, r = optic["majorradius"], optic["minorradius"]
        G = 4.0 * R**2 * (ux**2 + uz**2)
        H = 8.0 * R**2 * (ux * x + uz * z)
        I = 4.0 * R**2 * (x**2 + z**2)
        J = np.sum(U * U, axis=1)
        K = 2.0 * np.sum(U * P, axis=1)
        L = np.sum(P * P, axis=1) + R**2 - r**2
        a = J**2
        b = 2 * J * K
        c = 2 * J * L + K**2 - G
        d = 2 * K * L - H
        e = L**2 - I
        roots = _roots_rows(np.stack([a, b, c, d, e], axis=1))
        return _select_hit(P, U, roots, lambda q: q[..., 2] < -R, sup_xy)

    if k == "ellipsoidal":  # ART/ModuleMirror.py:662-683
        a_, b_ = optic["a"], optic["b"]
        da = (uy**2 + uz**2) / b_**2 + (ux / a_) ** 2
        db = 2 * ((uy * y + uz * z) / b_**2 + (ux * x) / a_**2)
        dc = (y**2 + z**2) / b_**2 + (x / a_) ** 2 - 1
        roots = _roots_rows(np.stack([da, db, dc], axis=1))
        C = optic_centre(optic)
        return _select_hit(
            P, U, roots, lambda q: q[..., 2] < 0,
            lambda q: support_include(sup, q[..., 0] - C[0], q[..., 1] - C[1]),
        )

    if k == "cylindrical":  # ART/ModuleMirror.py:824-844
        a = uy**2 + uz**2
        b = 2 * (uy * y + uz * z)
        c = y**2 + z**2 - optic["radius"] ** 2
        roots = _roots_rows(np.stack([a, b, c], axis=1))
        return _select_hit(P, U, roots, lambda q: q[..., 2] < 0, sup_xy)

    raise ValueError(k)


# --------------------------------------------------------------------------------------
# Zernike defects (ART/ModuleDefects.py:149-177, ART/recursive_zernike_generator.py:4-254)
# --------------------------------------------------------------------------------------
def zernike_gradient(x, y, max_order, dtype=np.float64):
    """Andersen's Cartesian recurrences as the reference evaluates them, vectorised over rays.

    ART/recursive_zernike_generator.py:36-253.  Returns three dicts keyed (n, m), m = 0..n:
    value, d/dx, d/dy (unnormalised polynomials; (1,0) = y, (1,1) = x).
    """
    if max_order < 2:  # :37-38
        max_order = 2
    x = np.asarray(x, dtype=dtype)  # dtype: np.longdouble for the extended-precision arbiter (art_oracle_ld.py)
    y = np.asarray(y, dtype=dtype)
    one = np.ones_like(x)
    zero = np.zeros_like(x)
    Z = {(0, 0): one, (1, 0): y, (1, 1): x}  # :52-54
    GX = {(0, 0): zero, (1, 0): zero, (1, 1): one}  # :57-59
    GY = {(0, 0): zero, (1, 0): one, (1, 1): zero}  # :61-63
    for n in range(2, max_order + 1):
        for m in range(0, n + 1):
            if m == 0:  # :80-95
                z = x * Z[(n - 1, 0)] + y * Z[(n - 1, n - 1)]
                gx = n * Z[(n - 1, 0)]
                gy = n * Z[(n - 1, n - 1)]
            elif m == n:  # :97-109
                z = x * Z[(n - 1, n - 1)] - y * Z[(n - 1, 0)]
                gx = n * Z[(n - 1, n - 1)]
                gy = -1.0 * n * Z[(n - 1, 0)]
            elif n % 2 != 0 and m == (n - 1) / 2:  # :111-144
                z = (y * Z[(n - 1, n - 1 - m)] + x * Z[(n - 1, m - 1)]
                     - y * Z[(n - 1, n - m)] - Z[(n - 2, m - 1)])
                gx = n * Z[(n - 1, m - 1)] + GX[(n - 2, m - 1)]
                gy = n * Z[(n - 1, n - 1 - m)] - n * Z[(n - 1, n - m)] + GY[(n - 2, m - 1)]
            elif n % 2 != 0 and m == (n - 1) / 2 + 1:  # :146-176
                z = (x * Z[(n - 1, m)] + y * Z[(n - 1, n - 1 - m)]
                     + x * Z[(n - 1, m - 1)] - Z[(n - 2, m - 1)])
                gx = n * Z[(n - 1, m)] + n * Z[(n - 1, m - 1)] + GX[(n - 2, m - 1)]
                gy = n * Z[(n - 1, n - 1 - m)] + GY[(n - 2, m - 1)]
            elif n % 2 == 0 and m == n / 2:  # :178-208
                z = 2.0 * x * Z[(n - 1, m)] + 2.0 * y * Z[(n - 1, m - 1)] - Z[(n - 2, m - 1)]
                gx = 2.0 * n * Z[(n - 1, m)] + GX[(n - 2, m - 1)]
                gy = 2.0 * n * Z[(n - 1, n - 1 - m)] + GY[(n - 2, m - 1)]
            else:  # :210-248
                z = (x * Z[(n - 1, m)] + y * Z[(n - 1, n - 1 - m)] + x * Z[(n - 1, m - 1)]
                     - y * Z[(n - 1, n - m)] - Z[(n - 2, m - 1)])
                gx = n * Z[(n - 1, m)] + n * Z[(n - 1, m - 1)] + GX[(n - 2, m - 1)]
                gy = n * Z[(n - 1, n - 1 - m)] - n * Z[(n - 1, n - m)] + GY[(n - 2, m - 1)]
            Z[(n, m)] = z
            GX[(n, m)] = gx
            GY[(n, m)] = gy
    return Z, GX, GY


def zernike_offset(defect, Q):
    """Zernike.get_offset, ART/ModuleDefects.py:168-174.  Q = point minus optic centre, (N,3)."""
    R = defect["R"]
    xy = Q / R
    Z, _, _ = zernike_gradient(xy[:, 0], xy[:, 1], defect["max_order"])
    out = np.zeros(Q.shape[0])
    for k, c in defect["coefficients"].items():
        out = out + c * Z[k]
    return out


def zernike_normal(defect, Q):
    """Zernike.get_normal, ART/ModuleDefects.py:156-166 (unnormalised (-dX,-dY,1))."""
    R = defect["R"]
    xy = Q / R
    _, GX, GY = zernike_gradient(xy[:, 0], xy[:, 1], defect["max_order"])
    dX = np.zeros(Q.shape[0])
    dY = np.zeros(Q.shape[0])
    for k, c in defect["coefficients"].items():
        dX = dX + c * GX[k]
        dY = dY + c * GY[k]
    dX = dX / R
    dY = dY / R
    return np.stack([-dX, -dY, np.ones_like(dX)], axis=-1)


# --------------------------------------------------------------------------------------
# gridded defects (ART/ModuleDefects.py:34-146: MeasuredMap, Fourrier)
# --------------------------------------------------------------------------------------
def _bilinear(V, x0, x1, y0, y1, x, y):
    """scipy.interpolate.RegularGridInterpolator((X, Y), V, method="linear") on the uniform grid
    X = linspace(x0, x1, V.shape[0]), Y = linspace(y0, y1, V.shape[1]) (ART/ModuleDefects.py:108-110)."""
    nx, ny = V.shape
    fx = (np.asarray(x) - x0) / (x1 - x0) * (nx - 1)
    fy = (np.asarray(y) - y0) / (y1 - y0) * (ny - 1)
    ix = np.clip(np.floor(fx).astype(int), 0, nx - 2)
    iy = np.clip(np.floor(fy).astype(int), 0, ny - 2)
    tx, ty = fx - ix, fy - iy
    return ((1 - tx) * (1 - ty) * V[ix, iy] + tx * (1 - ty) * V[ix + 1, iy]
            + (1 - tx) * ty * V[ix, iy + 1] + tx * ty * V[ix + 1, iy + 1])


def gridmap_offset(defect, Q):
    """get_offset of MeasuredMap / Fourrier, ART/ModuleDefects.py:60-61,131-137."""
    return _bilinear(defect["h"], defect["x0"], defect["x1"], defect["y0"], defect["y1"], Q[:, 0], Q[:, 1])


def gridmap_normal(defect, Q):
    """get_normal of MeasuredMap / Fourrier, ART/ModuleDefects.py:52-58,119-129: (dX, dY, 1)/norm --
    NOT negated, unlike Zernike.get_normal."""
    dX = _bilinear(defect["dx"], defect["x0"], defect["x1"], defect["y0"], defect["y1"], Q[:, 0], Q[:, 1])
    dY = _bilinear(defect["dy"], defect["x0"], defect["x1"], defect["y0"], defect["y1"], Q[:, 0], Q[:, 1])
    nrm = np.sqrt(dX**2 + dY**2 + 1)
    dX, dY = dX / nrm, dY / nrm
    return np.stack([dX, dY, np.sqrt(1 - dX**2 - dY**2)], axis=-1)


def defect_offset(D, Q):
    return gridmap_offset(D, Q) if D["kind"] == "gridmap" else zernike_offset(D, Q)


def defect_normal(D, Q):
    return gridmap_normal(D, Q) if D["kind"] == "gridmap" else zernike_normal(D, Q)


def normal_add(N1, N2):
    """ART/ModuleGeometry.py:394-407."""
    n1 = normalize(N1)
    n2 = normalize(N2)
    gX = (-n1[..., 0] / n1[..., 2]) + (-n2[..., 0] / n2[..., 2])
    gY = (-n1[..., 1] / n1[..., 2]) + (-n2[..., 1] / n2[..., 2])
    return np.stack([-gX, -gY, np.ones_like(gX)], axis=-1)


# --------------------------------------------------------------------------------------
# the trace (ART/ModuleProcessing.py:250-313)
# --------------------------------------------------------------------------------------
def trace_chain(P, U, elements, ignore_defects=True, numbers=None):
    """RayTracingCalculation, ART/ModuleProcessing.py:250-313.

    P, U: (N,3) lab-frame source points / unit directions.  `elements`: list of dicts
      {"optic": {...}, "position": (3,), "normal": (3,), "majoraxis": (3,)}
    with optic = {"kind", <surface params>, "support": (kind, ...), "defects": [ ... ]}.

    Returns one dict per element, describing the bundle AFTER that element for the rays that
    are still alive, in the source order: {"index" (into the source arrays), "number", "P", "U",
    "path" (sum of segments), "incidence"}.
    """
    P = np.asarray(P, dtype=np.float64)
    U = normalize(np.asarray(U, dtype=np.float64))  # Ray.vector setter, ART/ModuleOpticalRay.py:85-90
    n = P.shape[0]
    index = np.arange(n)
    if numbers is None:
        numbers = np.arange(n)
    path = np.zeros(n)
    out = []
    for el in elements:
        optic = el["optic"]
        pos = np.asarray(el["position"], dtype=np.float64)
        R = element_frame_matrix(el["normal"], el["majoraxis"])
        C = optic_centre(optic)
        # :289-295 lab -> element frame (the Ray.vector setter renormalises after each rotation)
        p1 = (P - pos) @ R.T + C
        u1 = normalize(U @ R.T)
        # :298-301 the optic acts
        hit, t = optic_intersection(optic, p1, u1)
        keep = np.nonzero(hit)[0]
        p1, u1, t = p1[keep], u1[keep], t[keep]
        index, path = index[keep], path[keep]
        Ph = u1 * t[:, None] + p1
        defects = optic.get("defects") or []
        if optic["kind"] == "mask":
            # _TransmitMaskRay, ART/ModuleMask.py:93-108
            incidence = angle_between(u1, np.broadcast_to(EZ, u1.shape))
            u2 = u1
        else:
            if defects:
                # DeformedMirror._get_intersection, ART/ModuleMirror.py:969-980
                h = np.zeros(Ph.shape[0])
                for D in defects:
                    h = h + defect_offset(D, Ph - C)
                alpha = angle_between(-u1, optic_normal(optic, Ph))
                Ph = Ph - u1 * (h / np.cos(alpha))[:, None]
            # ReflectionMirrorRayList / _ReflectionMirrorRay, ART/ModuleMirror.py:878-939
            nrm = optic_normal(optic, Ph)
            if defects and not ignore_defects:
                # DeformedMirror.get_normal, ART/ModuleMirror.py:952-961
                for D in defects:
                    nrm = normal_add(nrm, defect_normal(D, Ph - C))
                    nrm = nrm / norm(nrm)[:, None]
            # SymmetricalVector(-u, n): rotation of -u by pi about n == u - 2 (n.u) n
            u2 = normalize(u1 - 2.0 * np.sum(nrm * u1, axis=1)[:, None] * nrm)
            incidence = angle_between(-u1, nrm)
        path = path + norm(Ph - p1)
        # :306-309 element -> lab frame
        P = (Ph - C) @ R + pos
        U = normalize(u2 @ R)
        out.append({
            "index": index.copy(), "number": np.asarray(numbers)[index], "P": P.copy(), "U": U.copy(),
            "path": path.copy(), "incidence": incidence,
        })
    return out


def support_edge_distance(support, x, y):
    """Distance (mm) from (x, y) to the nearest edge LINE / CIRCLE of the support (same parameters as
    support_include).  Used only to REPORT rays whose survival may legitimately differ between two
    evaluations because they sit within rounding noise of an aperture edge ("bit-exact
    survival"); the survival rule itself is support_include."""
    kind, p = support[0], support[1:]
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)

    def circle(R, cx=0.0, cy=0.0):
        return np.abs(np.hypot(x - cx, y - cy) - abs(R))

    def rect(X, Y, cx=0.0, cy=0.0):
        return np.minimum(np.abs(np.abs(x - cx) - abs(X / 2)), np.abs(np.abs(y - cy) - abs(Y / 2)))

    if kind == "round":
        return circle(p[0])
    if kind == "roundhole":
        return np.minimum(circle(p[0]), circle(p[1], p[2], p[3]))
    if kind == "rect":
        return rect(p[0], p[1])
    if kind == "recthole":
        return np.minimum(rect(p[0], p[1]), circle(p[2], p[3], p[4]))
    if kind == "rectrecthole":
        return np.minimum(rect(p[0], p[1]), rect(p[2], p[3], p[4], p[5]))
    raise ValueError(f"unknown support kind {kind!r}")


def edge_margins(P, U, elements, ignore_defects=True):
    """For every source ray and element: how far (mm) the ray's hit point on that element's surface lies
    from the nearest aperture edge, NaN once the ray is lost for another reason (no intersection with the
    unbounded surface).  The rays are traced with every support opened wide, so a ray that the true support
    blocks still gets its margin at the blocking element; callers read the column of the element at which two
    evaluations disagree.  Returns an (N, K) array."""
    P = np.asarray(P, dtype=np.float64)
    n = P.shape[0]
    out = np.full((n, len(elements)), np.nan)
    wide = []
    for el in elements:
        optic = dict(el["optic"])
        optic["support"] = ("rect", 1e30, 1e30)
        if optic["kind"] == "mask":
            optic["kind"] = "plane"   # the plane hit itself; a mask only inverts the support test
        w = dict(el)
        w["optic"] = optic
        wide.append(w)
    prev_P, prev_U, prev_idx = P, normalize(np.asarray(U, dtype=np.float64)), np.arange(n)
    for k, el in enumerate(elements):
        step = trace_chain(prev_P, prev_U, [wide[k]], ignore_defects=ignore_defects)[0]
        idx = prev_idx[step["index"]]
        R = element_frame_matrix(el["normal"], el["majoraxis"])
        C = optic_centre(el["optic"])
        q = (step["P"] - np.asarray(el["position"], dtype=np.float64)) @ R.T + C   # hit point, element frame
        sup = el["optic"]["support"]
        off = C if el["optic"]["kind"] in ("parabolic", "ellipsoidal") else np.zeros(3)
        out[idx, k] = support_edge_distance(sup, q[:, 0] - off[0], q[:, 1] - off[1])
        if el["optic"]["kind"] == "mask":
            prev_P, prev_U = step["P"], prev_U[step["index"]]   # a mask does not deflect
        else:
            prev_P, prev_U = step["P"], step["U"]
        prev_idx = idx
    return out


def count_interactions(n_source, traced):
    """Interaction count of the metric: sum over elements of the rays ENTERING that element."""
    total, entering = 0, n_source
    for b in traced:
        total += entering
        entering = b["index"].size
    return total


# --------------------------------------------------------------------------------------
# detector and statistics (ART/ModuleDetector.py, ART/ModuleProcessing.py:464-532)
# --------------------------------------------------------------------------------------
def find_central_ray(P, U):
    """FindCentralRay, ART/ModuleProcessing.py:464-482 (the Ray ctor normalises the mean vector)."""
    return np.mean(P, axis=0), normalize(np.mean(U, axis=0))


def detector_autoplace(P, U, distance):
    """Detector.autoplace, ART/ModuleDetector.py:109-137.  Returns dict(centre, normal, refpoint)."""
    cp, cv = find_central_ray(P, U)
    normal = normalize(-cv)
    return {"centre": cp - normal * distance, "normal": normal, "refpoint": cp}


def detector_hits3d(det, P, U):
    """get_PointList3D + IntersectionLinePlane, ART/ModuleDetector.py:191-210, ModuleGeometry.py:48-57."""
    n = det["normal"]
    t = ((det["centre"] - P) @ n) / (U @ n)
    return U * t[:, None] + P


def detector_points2d(det, P, U):
    """get_PointList2D, ART/ModuleDetector.py:212-234."""
    H = detector_hits3d(det, P, U) - det["centre"]
    M = rotation_matrix(det["normal"], EZ)
    return (H @ M.T)[:, :2]


def centre_point_list(xy):
    """CentrePointList, ART/ModuleGeometry.py:222-245 (bounding-box midpoint, not the mean)."""
    c = (np.amax(xy, axis=0) + np.amin(xy, axis=0)) * 0.5
    return xy - c


def detector_points2d_centre(det, P, U):
    """get_PointList2DCentre, ART/ModuleDetector.py:236-252."""
    return centre_point_list(detector_points2d(det, P, U))


def detector_optical_paths(det, P, U, path):
    """Total path to the detector plane, ART/ModuleDetector.py:271-275."""
    H = detector_hits3d(det, P, U)
    return norm(P - H) + path


def detector_delays(det, P, U, path):
    """get_Delays, ART/ModuleDetector.py:254-279 (fs, relative to the unweighted mean path)."""
    L = detector_optical_paths(det, P, U, path)
    return (L - np.mean(L)) / LIGHTSPEED * 1e15


def standard_deviation(a):
    """StandardDeviation, ART/ModuleProcessing.py:485-507 (population; points: sqrt(sum var))."""
    a = np.asarray(a)
    if a.ndim == 1:
        return np.std(a)
    return np.sqrt(np.var(a, axis=0).sum())


def weighted_standard_deviation(a, w):
    """WeightedStandardDeviation, ART/ModuleProcessing.py:510-532."""
    a = np.asarray(a)
    avg = np.average(a, axis=0, weights=w)
    var = np.average((a - avg) ** 2, axis=0, weights=w)
    return np.sqrt(np.sum(var))


def diameter_point_list(xy):
    """DiameterPointList (2D branch), ART/ModuleGeometry.py:164-190."""
    ext = np.abs(np.amax(xy, axis=0) - np.amin(xy, axis=0))
    return np.max(ext)


def e_transmission(intensity_in, intensity_out):
    """getETransmission, ART/ModuleAnalysisAndPlots.py:62-77."""
    return 100 * np.sum(intensity_out) / np.sum(intensity_in)


def result_summary(det, P, U, path):
    """GetResultSummary, ART/ModuleAnalysisAndPlots.py:81-129 -> (SpotSizeSD mm, DurationSD fs)."""
    xy = detector_points2d_centre(det, P, U)
    return standard_deviation(xy), standard_deviation(detector_delays(det, P, U, path))


def detector_histograms(det, P, U, path, intensity=None, bins=(64, 64), delay_bins=128):
    """Binned SpotDiagram / DelayGraph data (ART/ModuleAnalysisAndPlots.py:133-250, 360-440 scatter
    get_PointList2DCentre and get_Delays ray by ray; the reference itself has no histogram): numpy's
    histogram2d / histogram of those very lists over their bounding box / range.  Returns the dict layout
    of attosecondraytracing_b200.engine.split_histogram."""
    xy = detector_points2d_centre(det, P, U)
    dl = detector_delays(det, P, U, path)
    w = np.ones(len(dl)) if intensity is None else np.asarray(intensity, dtype=np.float64)
    rng = [[xy[:, 0].min(), xy[:, 0].max()], [xy[:, 1].min(), xy[:, 1].max()]]
    cnt, xe, ye = np.histogram2d(xy[:, 0], xy[:, 1], bins=bins, range=rng)
    wsum, _, _ = np.histogram2d(xy[:, 0], xy[:, 1], bins=bins, range=rng, weights=w)
    dsum, _, _ = np.histogram2d(xy[:, 0], xy[:, 1], bins=bins, range=rng, weights=dl)
    tc, te = np.histogram(dl, bins=delay_bins, range=(dl.min(), dl.max()))
    tw, _ = np.histogram(dl, bins=delay_bins, range=(dl.min(), dl.max()), weights=w)
    with np.errstate(invalid="ignore", divide="ignore"):
        dmean = dsum / cnt
    return {"x_edges": xe, "y_edges": ye, "spot_count": cnt.astype(np.int64), "spot_intensity": wsum,
            "spot_delay": dmean, "delay_edges": te, "delay_count": tc.astype(np.int64), "delay_intensity": tw}


def find_optimal_distance(det, P, U, path, opt_for="intensity", amplitude=None, precision=3, weights=None):
    """FindOptimalDistance + _FindOptimalDistanceBIS, ART/ModuleProcessing.py:317-460, by brute force as
    the reference does it (every trial detector re-intersects every ray).  Returns
    (detector dict moved, OptSizeSpot, OptDuration)."""
    det = {k: np.array(v, dtype=np.float64) for k, v in det.items()}

    def distance(d):
        return abs(np.dot(d["normal"], d["centre"] - d["refpoint"]))

    def stats(d):
        xy = detector_points2d_centre(d, P, U)
        dl = detector_delays(d, P, U, path)
        if weights is None:
            return standard_deviation(xy), standard_deviation(dl)
        return weighted_standard_deviation(xy, weights), weighted_standard_deviation(dl, weights)

    first = distance(det)
    size = 2 * standard_deviation(detector_points2d_centre(det, P, U))
    na = numerical_aperture(U)
    if amplitude is None:
        amplitude = min(4 * np.ceil(size / np.tan(np.arcsin(na))), first)
    step = amplitude / 10
    spot = dur = np.nan
    for k in range(precision + 1):
        a_k, s_k = amplitude * 0.1**k, step * 0.1**k
        det["centre"] = det["centre"] - (-a_k) * det["normal"]  # shiftByDistance(-Amplitude)
        n = int(2 * a_k / s_k)
        spots, durs, fits = [], [], []
        for _ in range(n):
            sp, du = stats(det)
            spots.append(sp)
            durs.append(du)
            fits.append(sp**2 * du if opt_for == "intensity" else (du if opt_for == "duration" else sp))
            det["centre"] = det["centre"] - s_k * det["normal"]
        ind = fits.index(min(fits))
        spot = spots[ind] if opt_for != "duration" else np.nan
        dur = durs[ind] if opt_for in ("intensity", "duration") else np.nan
        det["centre"] = det["centre"] - (-(n - ind) * s_k) * det["normal"]
    return det, spot, dur


def numerical_aperture(U, refractive_index=1.0):
    """ReturnNumericalAperture, ART/ModuleProcessing.py:536-566."""
    cv = normalize(np.mean(U, axis=0))
    ang = angle_between(np.broadcast_to(cv, U.shape), U)
    return np.sin(np.amax(ang)) * refractive_index


# --------------------------------------------------------------------------------------
# sources (ART/ModuleSource.py) -- the synthetic bundles of the benchmark
# --------------------------------------------------------------------------------------
def spiral_vogel(nb_point, radius, k=None):
    """SpiralVogel, ART/ModuleGeometry.py:61-76; `k` optionally restricts to some indices."""
    golden = np.pi * (3 - np.sqrt(5))
    kk = np.arange(nb_point) if k is None else np.asarray(k)
    r = np.sqrt(kk / nb_point) * radius
    theta = golden * kk
    return np.stack([np.cos(theta) * r, np.sin(theta) * r], axis=1)


def _rotate_rays(P, U, axis1, axis2):
    """RotationRay, ART/ModuleGeometry.py:357-368: u' = R(p+u) - R(p), renormalised."""
    M = rotation_matrix(axis1, axis2)
    Pp = P @ M.T
    Up = (P + U) @ M.T - Pp
    return Pp, normalize(Up)


def point_source(S, axis, divergence, nb_rays, k=None):
    """PointSource / _Cone, ART/ModuleSource.py:23-81.  Returns (P, U, number)."""
    xy = spiral_vogel(nb_rays, 1 * np.tan(divergence), k)
    U = normalize(np.column_stack([xy, np.ones(xy.shape[0])]))
    P = np.zeros_like(U)
    P, U = _rotate_rays(P, U, EZ, np.asarray(axis, dtype=np.float64))
    num = np.arange(nb_rays) if k is None else np.asarray(k)
    return P + np.asarray(S, dtype=np.float64), U, num


def plane_wave_disk(centre, axis, radius, nb_rays, k=None):
    """PlaneWaveDisk, ART/ModuleSource.py:135-169: NbRays-1 rays numbered 0..NbRays-2."""
    kk = np.arange(nb_rays - 1) if k is None else np.asarray(k)
    xy = spiral_vogel(nb_rays, radius, kk)
    P = np.column_stack([xy, np.zeros(xy.shape[0])])
    U = np.broadcast_to(EZ, P.shape).copy()
    P, U = _rotate_rays(P, U, EZ, np.asarray(axis, dtype=np.float64))
    return P + np.asarray(centre, dtype=np.float64), U, kk


def extended_source_layout(diameter, nb_rays):
    """(number of point sources, rays per point source) of ExtendedSource, ART/ModuleSource.py:105-113."""
    n_ps = max(30, int(250 * diameter))
    n_ps = min(n_ps, int(nb_rays / 300))
    return n_ps, max(300, int(nb_rays / n_ps))


def extended_source(S, axis, diameter, divergence, nb_rays):
    """ExtendedSource, ART/ModuleSource.py:85-131: point sources on a Vogel spiral over a disk, each
    emitting the same cone; ray number = k * rays_per_source + l.  Returns (P, U, number)."""
    n_ps, per = extended_source_layout(diameter, nb_rays)
    centres = spiral_vogel(n_ps, diameter / 2)
    cone = normalize(np.column_stack([spiral_vogel(per, np.tan(divergence)), np.ones(per)]))
    P = np.repeat(np.column_stack([centres, np.zeros(n_ps)]), per, axis=0)
    U = np.tile(cone, (n_ps, 1))
    P, U = _rotate_rays(P, U, EZ, np.asarray(axis, dtype=np.float64))
    return P + np.asarray(S, dtype=np.float64), U, np.arange(n_ps * per)


def gaussian_intensity(P, U, fraction=1 / np.e**2):
    """ApplyGaussianIntensityToRayList, ART/ModuleSource.py:219-261."""
    axis = normalize(np.mean(U, axis=0))
    ang = angle_between(np.broadcast_to(axis, U.shape), U)
    div = np.max(ang) if ang.size else 0.0
    div = max(0.0, div)
    if div > 1e-12:
        return np.exp(-2 * (np.tan(ang) / div) ** 2 * -0.5 * np.log(fraction))
    d = norm(P)
    return np.exp(-2 * (d / np.max(d)) ** 2 * -0.5 * np.log(fraction))


def source_for(source_properties, first_support=None, k=None):
    """Source selection of _singleOEPlacement, ART/ModuleProcessing.py:55-79 (no ExtendedSource).

    Returns (P, U, number, intensity) of the bundle launched from the origin along +x.
    With `k` only those indices are generated; the intensity normalisation (max angle / max
    distance) is then still taken over the FULL bundle, which for the Vogel spiral is its last ray.
    """
    div = source_properties["Divergence"]
    size = source_properties["SourceSize"]
    n = source_properties["NumberRays"]
    if div == 0:
        if size == 0:
            radius = first_support
        else:
            radius = size / 2
        P, U, num = plane_wave_disk(np.zeros(3), EX, radius, n, k)
        Pf, Uf, _ = (P, U, num) if k is None else plane_wave_disk(np.zeros(3), EX, radius, n)
    else:
        if size != 0:
            raise NotImplementedError("ExtendedSource is not part of the synthetic benchmark bundles")
        P, U, num = point_source(np.zeros(3), EX, div, n, k)
        Pf, Uf, _ = (P, U, num) if k is None else point_source(np.zeros(3), EX, div, n)
    if k is None:
        inten = gaussian_intensity(P, U)
    else:
        # same formula, normalisation from the full bundle
        frac = 1 / np.e**2
        axis = normalize(np.mean(Uf, axis=0))
        angf = angle_between(np.broadcast_to(axis, Uf.shape), Uf)
        dmax = max(0.0, np.max(angf))
        if dmax > 1e-12:
            ang = angle_between(np.broadcast_to(axis, U.shape), U)
            inten = np.exp(-2 * (np.tan(ang) / dmax) ** 2 * -0.5 * np.log(frac))
        else:
            inten = np.exp(-2 * (norm(P) / np.max(norm(Pf))) ** 2 * -0.5 * np.log(frac))
    return P, U, num, inten
