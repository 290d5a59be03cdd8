"""Pins the oracle (oracle/art_oracle.py) against outputs of the reference itself.

The fixtures under tests/golden were produced by running the unmodified reference
(oracle/gen_golden.py); the reference has no tests or golden vectors of its own.
Bars: survival / ray numbers bit-exact, points <= 1e-9 mm (3e-8 on the 5 m telescope, see
golden_util.point_tol), per-ray delays <= 0.01 as.
"""
import numpy as np
import pytest

import art_oracle as orc
from golden_util import DELAY_TOL_FS, Golden, dir_tol, golden_names, point_tol

NAMES = golden_names()


@pytest.mark.parametrize("name", NAMES)
def test_trace_matches_reference(name):
    g = Golden(name)
    traced = orc.trace_chain(g["src_P"], g["src_U"], g.oracle_elements(), ignore_defects=g.ignore_defects,
                             numbers=g["src_num"])
    assert len(traced) == g.n_elements
    tol = point_tol(name)
    for k, b in enumerate(traced):
        ref = g.out(k)
        assert np.array_equal(b["number"], ref["num"]), f"{name}: survivors differ after element {k}"
        if ref["num"].size == 0:
            continue
        assert np.max(np.abs(b["P"] - ref["P"])) <= tol, (name, k, np.max(np.abs(b["P"] - ref["P"])))
        assert np.max(np.abs(b["U"] - ref["U"])) <= dir_tol(name)
        assert np.max(np.abs(b["path"] - ref["path"])) <= tol * 2
        assert np.max(np.abs(b["incidence"] - ref["inc"])) <= 1e-9
    assert orc.count_interactions(g["src_P"].shape[0], traced) == g.spec["interactions"]


@pytest.mark.parametrize("name", [n for n in NAMES if "det_centre" in Golden(n)])
def test_detector_and_statistics_match_reference(name):
    g = Golden(name)
    last = g.out(g.n_elements - 1)
    P, U, path = last["P"], last["U"], last["path"]
    det = orc.detector_autoplace(P, U, g.spec["detector_distance"])
    assert np.max(np.abs(det["centre"] - g["det_centre"])) <= 1e-10
    assert np.max(np.abs(det["normal"] - g["det_normal"])) <= 1e-14
    assert np.max(np.abs(det["refpoint"] - g["det_refpoint"])) <= 1e-10
    det = {"centre": g["det_centre"], "normal": g["det_normal"], "refpoint": g["det_refpoint"]}
    xy = orc.detector_points2d_centre(det, P, U)
    assert np.max(np.abs(xy - g["det_xy_centre"])) <= 1e-9
    delays = orc.detector_delays(det, P, U, path)
    assert np.max(np.abs(delays - g["det_delays"])) <= DELAY_TOL_FS
    sd, dur = orc.result_summary(det, P, U, path)
    assert abs(sd - g["SpotSizeSD"]) <= 1e-9
    assert abs(dur - g["DurationSD"]) <= DELAY_TOL_FS
    src_I = g["src_I"]
    idx = np.searchsorted(g["src_num"], last["num"])
    w = src_I[idx]
    assert abs(orc.e_transmission(src_I, w) - g["ETransmission"]) <= 1e-9
    assert abs(orc.weighted_standard_deviation(xy, w) - g["SpotSizeSD_w"]) <= 1e-9
    assert abs(orc.weighted_standard_deviation(delays, w) - g["DurationSD_w"]) <= DELAY_TOL_FS
    assert abs(orc.numerical_aperture(U) - g["NA"]) <= 1e-12
    assert abs(orc.diameter_point_list(xy) - g["Diameter"]) <= 1e-9


@pytest.mark.parametrize("name", [n for n in NAMES if "_sub" not in n])
def test_sources_match_reference(name):
    """oracle.source_for reproduces the reference's PointSource / PlaneWaveDisk + Gaussian weights."""
    g = Golden(name)
    P, U, num, inten = orc.source_for(g.spec["source"])
    assert np.array_equal(num, g["src_num"])
    assert np.max(np.abs(P - g["src_P"])) <= 1e-12
    assert np.max(np.abs(U - g["src_U"])) <= 1e-14  # reference: u = R(p+u) - R(p), noise ~ |p| eps
    assert np.max(np.abs(inten - g["src_I"])) <= 1e-13


def test_subset_sources_are_the_full_bundle_rows():
    """Index-restricted generation == the same rows of the full bundle (what the scale tests rely on)."""
    sp = {"Divergence": 0.015, "SourceSize": 0, "NumberRays": 20000}
    P, U, num, inten = orc.source_for(sp)
    k = np.array([0, 5, 777, 19999])
    Pk, Uk, numk, ik = orc.source_for(sp, k=k)
    assert np.array_equal(numk, k)
    assert np.array_equal(Pk, P[k]) and np.array_equal(Uk, U[k])
    assert np.max(np.abs(ik - inten[k])) <= 1e-15
    sp = {"Divergence": 0, "SourceSize": 30, "NumberRays": 5000}
    P, U, num, inten = orc.source_for(sp)
    assert P.shape[0] == 4999  # PlaneWaveDisk off-by-one, ART/ModuleSource.py:162
    k = np.array([0, 17, 4998])
    Pk, Uk, numk, ik = orc.source_for(sp, k=k)
    assert np.array_equal(Pk, P[k]) and np.max(np.abs(ik - inten[k])) <= 1e-15


# ----------------------------------------------------------------------------------------------
# detector-distance optimiser
# ----------------------------------------------------------------------------------------------
def _optdist_cases():
    import os
    from golden_util import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "optdist.npz"))
    return z, sorted(k for k in z.files if k != "meta" and not k.endswith("__centre"))


def numpy_scan_sums(det, l0, P, U, path, w):
    """The ART_S_* scan sums (include/art_b200.h) by oracle arithmetic -- what scan_kernel accumulates."""
    xy = orc.detector_points2d(det, P, U)
    L = orc.detector_optical_paths(det, P, U, path)
    d = L - l0
    cv = -np.asarray(det["normal"])
    cu = U @ cv
    g = 1.0 / cu
    gp = 0.5 * np.sum((U - cv) ** 2, axis=1) * g
    M = orc.rotation_matrix(det["normal"], orc.EZ)
    ru = U @ M.T
    ax, ay = g * ru[:, 0], g * ru[:, 1]
    x, y = xy[:, 0], xy[:, 1]
    t = [x, y, ax, ay, x * x, y * y, ax * ax, ay * ay, x * ax, y * ay, d, gp, d * d, gp * gp, d * gp]
    row = np.zeros(32)
    row[0], row[1] = P.shape[0], w.sum()
    for j, v in enumerate(t):
        row[2 + j] = v.sum()
        row[17 + j] = (w * v).sum()
    return row


@pytest.mark.parametrize("key", _optdist_cases()[1])
def test_find_optimal_distance_matches_reference(key):
    """Both the oracle's brute-force restatement and the product's closed-form search
    (engine.optimal_shift_from_scan on the 32 scan sums) land where the reference's FindOptimalDistance did."""
    from attosecondraytracing_b200.engine import optimal_shift_from_scan
    z, _ = _optdist_cases()
    name, opt_for, wflag = key.split("__")
    ref_dist, ref_spot, ref_dur = z[key]
    g = Golden(name)
    last = g.out(g.n_elements - 1)
    P, U, path = last["P"], last["U"], last["path"]
    det = {"centre": g["det_centre"], "normal": g["det_normal"], "refpoint": g["det_refpoint"]}
    w = g["src_I"][np.searchsorted(g["src_num"], last["num"])]
    weights = w if wflag == "w" else None
    # finest step of the search = Amplitude * 1e-4; neighbouring positions can tie within rounding
    first = float(abs(np.dot(det["normal"], det["centre"] - det["refpoint"])))
    d2, spot, dur = orc.find_optimal_distance(det, P, U, path, opt_for, None, 3, weights)
    dist = abs(np.dot(d2["normal"], d2["centre"] - d2["refpoint"]))
    tol_d = 2.5e-4 * first
    assert abs(dist - ref_dist) <= tol_d, (dist, ref_dist)
    if opt_for != "duration":
        assert abs(spot - ref_spot) <= 1e-6 * max(ref_spot, 1e-9) + 1e-9
    assert abs(dur - ref_dur) <= 1e-5 * max(ref_dur, 1.0)
    # closed form on the scan sums
    l0 = path.mean() + first
    scan = numpy_scan_sums(det, l0, P, U, path, w)
    sd0 = orc.standard_deviation(orc.detector_points2d_centre(det, P, U))
    s, spot2, dur2, amp = optimal_shift_from_scan(scan, first, sd0, orc.numerical_aperture(U), opt_for, None, 3,
                                                  wflag == "w")
    assert abs((first + s) - ref_dist) <= tol_d, (first + s, ref_dist)
    if opt_for != "duration":
        assert abs(spot2 - ref_spot) <= 1e-6 * max(ref_spot, 1e-9) + 1e-9
    assert abs(dur2 - ref_dur) <= 1e-5 * max(ref_dur, 1.0)


@pytest.mark.parametrize("case", ["a", "b"])
def test_extended_source_matches_reference(case):
    import os
    from golden_util import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "extsource.npz"))
    diameter, divergence, nb = z[case + "_params"]
    P, U, num = orc.extended_source(np.zeros(3), orc.EX, float(diameter), float(divergence), int(nb))
    assert np.array_equal(num, z[case + "_num"])
    assert np.max(np.abs(P - z[case + "_P"])) <= 1e-14 and np.max(np.abs(U - z[case + "_U"])) <= 1e-14
    assert np.max(np.abs(orc.gaussian_intensity(P, U) - z[case + "_I"])) <= 1e-12


@pytest.mark.parametrize("key,tag,ignore", [("toroid", "out", False), ("sphere_cx", "out", False),
                                            ("parabola_hole", "out", False), ("mask", "out", True),
                                            ("sphere_zernike", "out", False), ("sphere_zernike", "outign", True)])
def test_oracle_element_frame_functions_match_reference(key, tag, ignore):
    """The oracle on rays given in the optic's own frame (one element whose frame is the lab frame) against
    ReflectionMirrorRayList / TransmitMaskRayList of the unmodified reference (tests/golden/raylist.npz)."""
    import bench
    from golden_util import RayListGolden, compare_bundle
    g = RayListGolden()
    P, U, num = g.source(key)
    els = bench._oracle_elements([g.identity_element(key)])
    res = orc.trace_chain(P, U, els, ignore_defects=ignore, numbers=num)[0]
    compare_bundle("raylist_" + key, 0, g.out(key, tag), res["number"], res["P"], res["U"], res["path"], res["incidence"])
