"""CPU tests of the host side: the C ABI loads and exports what include/art_b200.h declares, the
lowering of scene objects, the alignment (OEPlacement) and misalignment arithmetic against the
poses the reference produced, the host Zernike evaluation against the oracle's recurrences, and the
statistics derived from a moments row.  No GPU compute is called here."""
import copy
import ctypes as C
import os
import re

import numpy as np
import pytest

import art_oracle as orc
from golden_util import Golden, build_optic, golden_names, golden_optical_elements
from test_distributed_cpu import oracle_central, oracle_moments

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = golden_names()


def test_library_exports_every_declared_symbol():
    from attosecondraytracing_b200 import _cabi
    header = open(os.path.join(ROOT, "include", "art_b200.h")).read()
    declared = set(re.findall(r"\b(art_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    lib = _cabi.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libart_b200.so does not export {name}"
    assert declared == set(_cabi.EXPORTED_SYMBOLS), declared ^ set(_cabi.EXPORTED_SYMBOLS)
    assert lib.art_version() == 100
    sizes = (C.c_int32 * 6)()
    assert lib.art_abi_sizes(sizes) == 0
    assert list(sizes) == [C.sizeof(t) for t in (_cabi.ArtElementDesc, _cabi.ArtZernikeDesc, _cabi.ArtBundleView,
                                                  _cabi.ArtDetector, _cabi.ArtGridMapDesc,
                                                  _cabi.ArtSourceDesc)] == [200, 40, 88, 184, 64, 128]


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from attosecondraytracing_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU path"):
        _cabi.lib()


def test_tracing_without_cuda_raises():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import attosecondraytracing_b200.ModuleProcessing as mp
    from attosecondraytracing_b200.ModuleOpticalRay import RayBundle
    g = Golden("cfg1_par")
    with pytest.raises(RuntimeError, match="no CPU path"):
        mp.RayTracingCalculation(RayBundle.from_numpy(g["src_P"], g["src_U"]), golden_optical_elements(g))


@pytest.mark.parametrize("name", NAMES)
def test_alignment_reproduces_reference_poses(name):
    """place_optical_elements (OEPlacement) + the misalignment methods give the reference's poses.
    On the 5 m telescope the reference's own chief-ray noise is ~3e-9 mm."""
    import attosecondraytracing_b200.ModuleProcessing as mp
    g = Golden(name)
    s = g.spec
    optics = [build_optic(o) for o in s["optics"]]
    oes = mp.place_optical_elements(optics, s["distances"], s["incidences"], s["plane_angles"])
    for op in s.get("post", []):
        getattr(oes[op["element"]], op["op"])(op["value"])
    tol = 1e-8 if (name.startswith("cfg5") or name.startswith("tele")) else 1e-10
    for k, oe in enumerate(oes):
        assert np.max(np.abs(oe.position - g[f"el{k}_position"])) <= tol
        assert np.max(np.abs(oe.normal - g[f"el{k}_normal"])) <= 1e-11
        assert np.max(np.abs(oe.majoraxis - g[f"el{k}_majoraxis"])) <= 1e-11
        assert np.max(np.abs(optics[k].get_centre() - np.array(s["derived_optics"][k]["centre"]))) <= 1e-12
        assert optics[k].type == s["derived_optics"][k]["type"]


def test_element_rotation_matches_oracle_including_branches():
    from attosecondraytracing_b200 import _cabi
    rng = np.random.default_rng(3)
    cases = [(np.array([0, 0, 1.0]), np.array([1.0, 0, 0])), (np.array([0, 0, -1.0]), np.array([1.0, 0, 0])),
             (np.array([0, 0, 1.0]), np.array([-1.0, 0, 0])), (np.array([0, 0, -1.0]), np.array([0, 1.0, 0]))]
    for _ in range(20):
        n = orc.normalize(rng.normal(size=3))
        m = orc.normalize(np.cross(n, rng.normal(size=3)))
        cases.append((n, m))
    for n, m in cases:
        R = np.array(_cabi.element_rotation(n, m))
        assert np.max(np.abs(R - orc.element_frame_matrix(n, m))) <= 1e-14


def test_lowering_fills_descriptors():
    from attosecondraytracing_b200 import _cabi
    from attosecondraytracing_b200._lowering import LoweredChain
    g = Golden("cfg4_zern_def")
    low = LoweredChain([golden_optical_elements(g)])
    d = low.elements[0]
    assert d.surface == _cabi.SURF_PARABOLIC and d.support == _cabi.SUPP_RECT
    assert d.n_defects == 1 and low.n_defects == 1
    assert low.defects[0].n_coefficients == len(g.spec["optics"][0]["defects"][0]["coefficients"])
    assert abs(low.defects[0].radius - np.sqrt(40**2 + 40**2) / 2) < 1e-12
    g3 = Golden("cfg3_2tor")
    low3 = LoweredChain([golden_optical_elements(g3)] * 3)
    assert low3.n_variants == 3 and low3.n_elements == 3
    assert [low3.elements[i].surface for i in range(3)] == [_cabi.SURF_MASK, _cabi.SURF_TOROIDAL, _cabi.SURF_TOROIDAL]


def test_host_zernike_matches_oracle_recurrences():
    """ModuleDefects.Zernike (radial-polynomial form, as in the kernel) == the reference's Cartesian
    recurrences (oracle.zernike_gradient) for every (n, m) up to order 12."""
    import attosecondraytracing_b200.ModuleDefects as mdef
    import attosecondraytracing_b200.ModuleSupport as msupp
    sup = msupp.SupportRound(10.0)
    rng = np.random.default_rng(0)
    pts = rng.uniform(-7, 7, size=(6, 2))
    Z, GX, GY = orc.zernike_gradient(pts[:, 0] / 10.0, pts[:, 1] / 10.0, 12)
    for n in range(0, 13):
        for m in range(0, n + 1):
            z = mdef.Zernike(sup, {(2, 0): 0.0, (n, m): 1.0})
            for i, p in enumerate(pts):
                q = np.array([p[0], p[1], 0.0])
                assert abs(z.get_offset(q) - Z[(n, m)][i]) <= 1e-12 * max(1.0, abs(Z[(n, m)][i])), (n, m)
                nrm = z.get_normal(q)
                assert abs(-nrm[0] * 10.0 - GX[(n, m)][i]) <= 1e-11 * max(1.0, abs(GX[(n, m)][i])), (n, m)
                assert abs(-nrm[1] * 10.0 - GY[(n, m)][i]) <= 1e-11 * max(1.0, abs(GY[(n, m)][i])), (n, m)


def test_supports_include_semantics():
    import attosecondraytracing_b200.ModuleSupport as msupp
    rng = np.random.default_rng(1)
    pts = rng.uniform(-40, 40, size=(400, 2))
    specs = [("round", (30,)), ("roundhole", (30, 5, 10, 5)), ("rect", (60, 30)), ("recthole", (60, 60, 5, 9, -8)),
             ("rectrecthole", (30, 20, 10, 6, 2, -1))]
    cls = {"round": msupp.SupportRound, "roundhole": msupp.SupportRoundHole, "rect": msupp.SupportRectangle,
           "recthole": msupp.SupportRectangleHole, "rectrecthole": msupp.SupportRectangleRectHole}
    for kind, p in specs:
        sup = cls[kind](*p)
        want = orc.support_include((kind,) + p, pts[:, 0], pts[:, 1])
        got = np.array([sup._IncludeSupport(np.array([x, y, 0.0])) for x, y in pts])
        assert np.array_equal(got, want), kind
        assert abs(sup._CircumCirc() - orc.support_circum_circ((kind,) + p)) < 1e-14
    # edges are inclusive (ART/ModuleGeometry.py:249-268)
    assert msupp.SupportRound(5)._IncludeSupport(np.array([3.0, 4.0, 0]))
    assert msupp.SupportRectangle(4, 2)._IncludeSupport(np.array([2.0, -1.0, 0]))


def test_optical_element_pose_rules():
    import attosecondraytracing_b200.ModuleOpticalElement as moe
    import attosecondraytracing_b200.ModuleMirror as mmirror
    import attosecondraytracing_b200.ModuleSupport as msupp
    m = mmirror.MirrorPlane(msupp.SupportRound(10))
    oe = moe.OpticalElement(m, np.zeros(3), np.array([0, 0, 2.0]), np.array([3.0, 0, 0]))
    assert np.allclose(oe.normal, [0, 0, 1]) and np.allclose(oe.majoraxis, [1, 0, 0])
    with pytest.raises(ValueError):
        oe.majoraxis = np.array([0.0, 0.1, 1.0])
    with pytest.raises(TypeError):
        oe.position = [0, 0, 0]
    oe.rotate_pitch_by(10.0)  # about normal x major = +y; the major axis co-rotates and stays perpendicular
    assert abs(np.dot(oe.normal, oe.majoraxis)) < 1e-15
    assert np.allclose(oe.normal, [np.sin(np.deg2rad(10)), 0, np.cos(np.deg2rad(10))])
    oe.shift_along_cross(2.0)
    assert np.allclose(oe.position, 2.0 * np.cross(oe.normal, oe.majoraxis))
    h0 = hash(oe)
    oe.rotate_yaw_by(1.0)
    assert hash(oe) != h0


def test_chain_and_detector_argument_checks():
    import attosecondraytracing_b200.ModuleOpticalChain as moc
    import attosecondraytracing_b200.ModuleDetector as mdet
    from attosecondraytracing_b200.ModuleOpticalRay import Ray, RayBundle
    g = Golden("cfg1_par")
    oes = golden_optical_elements(g)
    with pytest.raises(TypeError):
        moc.OpticalChain("rays", oes)
    with pytest.raises(TypeError):
        moc.OpticalChain(RayBundle.from_numpy(g["src_P"], g["src_U"]), "elements")
    rays = [Ray(g["src_P"][i].copy(), g["src_U"][i].copy(), Number=i, Intensity=1.0) for i in range(5)]
    ch = moc.OpticalChain(rays, oes, "five rays")
    assert len(ch.source_rays) == 5 and ch.source_rays[3].number == 3
    chains = ch.get_OE_loop_list(0, "pitch", np.linspace(-0.1, 0.1, 5))
    assert len(chains) == 5 and chains[0].source_rays is ch.source_rays
    assert chains[1].loop_variable_name.endswith("pitch rotation (deg)")
    assert not np.allclose(chains[0].optical_elements[0].normal, chains[4].optical_elements[0].normal)
    with pytest.raises(ValueError):
        ch.get_OE_loop_list(0, "wobble", [1.0])
    with pytest.raises(TypeError):
        mdet.Detector(np.zeros(3), Normal=np.zeros(3))
    d = mdet.Detector(np.array([0.0, 0, 0]), np.array([0.0, 0, 5.0]), np.array([0.0, 0, -2.0]))
    assert np.allclose(d.normal, [0, 0, -1]) and abs(d.get_distance() - 5.0) < 1e-15
    d.shiftByDistance(1.0)
    assert abs(d.get_distance() - 6.0) < 1e-15
    d.shiftToDistance(2.5)
    assert abs(d.get_distance() - 2.5) < 1e-15


def test_ray_bundle_list_facade_on_cpu():
    from attosecondraytracing_b200.ModuleOpticalRay import RayBundle
    import torch
    g = Golden("cfg3_2tor")
    b = RayBundle.from_numpy(g["src_P"], g["src_U"], intensity=g["src_I"])
    assert len(b) == 1000 and b[999].number == 999 and abs(b[7].intensity - g["src_I"][7]) < 1e-16
    b.alive = torch.ones(b.n, dtype=torch.uint8)
    b.alive[::2] = 0
    b.invalidate()
    assert len(b) == 500 and b[0].number == 1 and [r.number for r in b][:3] == [1, 3, 5]
    with pytest.raises(IndexError):
        b[500]
    d = b.to_numpy()
    assert np.array_equal(d["number"], np.arange(1, 1000, 2)) and np.array_equal(d["P"], g["src_P"][1::2])


@pytest.mark.parametrize("name", ["cfg3_2tor", "cfg1_par", "cfg5_tele", "sph_recthole"])
def test_statistics_from_moments_row(name):
    """engine.summary_from_moments turns the 24 sums / extents into the reference's statistics."""
    from attosecondraytracing_b200.engine import summary_from_moments
    g = Golden(name)
    last = g.out(g.n_elements - 1)
    det = {"centre": g["det_centre"], "normal": g["det_normal"], "refpoint": g["det_refpoint"]}
    idx = np.searchsorted(g["src_num"], last["num"])
    w = g["src_I"][idx]
    l0 = last["path"].mean() + g.spec["detector_distance"]
    m = oracle_moments(det, l0, last["P"], last["U"], last["path"], w)
    c = oracle_central(last["P"], last["U"], last["path"], w, g["src_I"].sum())
    s = summary_from_moments(m, c)
    assert abs(s["SpotSizeSD"] - g["SpotSizeSD"]) <= 1e-9
    assert abs(s["DurationSD"] - g["DurationSD"]) <= 1e-5
    assert abs(s["SpotSizeSD_w"] - g["SpotSizeSD_w"]) <= 1e-9
    assert abs(s["DurationSD_w"] - g["DurationSD_w"]) <= 1e-5
    assert abs(s["ETransmission"] - g["ETransmission"]) <= 1e-9
    assert abs(s["Diameter"] - g["Diameter"]) <= 1e-9
    assert abs(s["NA"] - g["NA"]) <= 1e-10
    assert summary_from_moments(np.zeros(24))["n_rays"] == 0


def test_fourier_defect_generator_reproduces_reference_maps():
    """ModuleDefects.Fourrier with the same numpy seed builds the maps the reference built
    (tests/golden/par_fourier_ign.npz holds the reference's deformation / DerivX / DerivY)."""
    import attosecondraytracing_b200.ModuleDefects as mdef
    import attosecondraytracing_b200.ModuleSupport as msupp
    g = Golden("par_fourier_ign")
    d = g.spec["optics"][0]["defects"][0]
    f = mdef.Fourrier(msupp.SupportRectangle(40, 40), d["rms"], slope=d["slope"], smallest=d["smallest"], seed=d["seed"])
    for mine, key in ((f._h, "map0_h"), (f._dx, "map0_dx"), (f._dy, "map0_dy")):
        ref = g[key]
        assert mine.shape == ref.shape
        assert np.max(np.abs(mine - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
    dd = g.spec["derived_optics"][0]["defects"][0]
    assert f._extent == (dd["x0"], dd["x1"], dd["y0"], dd["y1"])
    assert abs(f.RMS() - d["rms"]) < 1e-15
    # host single-point evaluation == oracle bilinear interpolation
    od = g.oracle_elements()[0]["optic"]["defects"][0]
    pts = np.array([[3.3, -7.1, 0.0], [-19.9, 19.9, 0.0], [0.0, 0.0, 0.0]])
    for p in pts:
        assert abs(f.get_offset(p) - orc.gridmap_offset(od, p[None, :])[0]) <= 1e-18 + 1e-12 * abs(f.get_offset(p))
        assert np.max(np.abs(f.get_normal(p) - orc.gridmap_normal(od, p[None, :])[0])) <= 1e-14


def test_measured_map_follows_the_reference_conventions():
    """ModuleDefects.MeasuredMap keeps the reference's layout (ART/ModuleDefects.py:34-61): slopes by np.gradient
    with one spacing per axis, interpolation grids paired with the TRANSPOSED arrays (square maps only).  The
    numbers themselves are pinned by tests/golden/par_measured_*.npz (the reference run through the documented
    numpy-2 compatibility patch)."""
    import attosecondraytracing_b200.ModuleDefects as mdef
    import attosecondraytracing_b200.ModuleSupport as msupp
    sup = msupp.SupportRectangle(40, 20)
    i = np.arange(50)[:, None]
    j = np.arange(50)[None, :]
    Map = 1e-4 * (0.02 * i + 0.05 * j)  # a tilted plane: constant slopes along both array axes
    m = mdef.MeasuredMap(sup, Map)
    assert np.allclose(m.DerivX, 1e-4 * 0.02 / (40 / 50)) and np.allclose(m.DerivY, 1e-4 * 0.05 / (20 / 50))
    assert m._extent == (-40.0, 40.0, -20.0, 20.0) and m._h.shape == (50, 50)
    assert np.array_equal(m._h, Map.T)           # value at (X[a], Y[b]) is Map[b, a]
    n = m.get_normal(np.array([1.0, 2.0, 0.0]))
    assert abs(n[0] / n[2] - 1e-4 * 0.02 / (40 / 50)) < 1e-15
    with pytest.raises(ValueError, match="square"):
        mdef.MeasuredMap(sup, np.zeros((50, 30)))
    # the fixture's arrays are what this class builds from the same map
    import scenes as sc
    g = Golden("par_measured_def")
    d = g.spec["optics"][0]["defects"][0]
    mm = mdef.MeasuredMap(msupp.SupportRectangle(40, 40), sc.measured_map(d["nx"], d["ny"], d["amplitude"]))
    dd = g.spec["derived_optics"][0]["defects"][0]
    assert np.array_equal(mm._h, g[dd["arrays"] + "_h"])
    assert np.max(np.abs(mm._dx - g[dd["arrays"] + "_dx"])) <= 1e-18
    assert np.max(np.abs(mm._dy - g[dd["arrays"] + "_dy"])) <= 1e-18
    assert mm._extent == (dd["x0"], dd["x1"], dd["y0"], dd["y1"])


def test_save_and_load_compressed_round_trip(tmp_path, capsys):
    import attosecondraytracing_b200.ModuleProcessing as mp
    from attosecondraytracing_b200.ModuleOpticalRay import RayBundle
    import torch
    g = Golden("cfg3_2tor")
    b = RayBundle.from_numpy(g["src_P"], g["src_U"], intensity=g["src_I"])
    b.alive = torch.ones(b.n, dtype=torch.uint8)
    b.alive[100:] = 0
    b.invalidate()
    name = mp.save_compressed({"rays": b, "SpotSizeSD": 1.25}, str(tmp_path / "kept"))
    assert name.endswith("kept_0")
    assert mp.save_compressed({"x": 1}, str(tmp_path / "kept")).endswith("kept_1")  # never overwrites
    back = mp.load_compressed(name)
    assert back["SpotSizeSD"] == 1.25 and len(back["rays"]) == 100
    assert np.array_equal(back["rays"].to_numpy()["P"], g["src_P"][:100])
    assert "Saved results" in capsys.readouterr().out


def test_header_macros_match_the_python_constants():
    """ART_HIST_LEN / ART_PEER_BUFFER_BYTES / ART_HIST_FIXED_ONE of the header evaluated by the C compiler
    agree with _cabi.hist_len / peer_buffer_bytes / HIST_FIXED_ONE (sizes of caller-owned buffers)."""
    import subprocess
    import tempfile
    from attosecondraytracing_b200 import _cabi
    src = r'''
#include <stdio.h>
#include "art_b200.h"
int main(void) {
  printf("%lld %lld %lld %lld %.1f %d %d\n", (long long)ART_HIST_LEN(64, 48, 128), (long long)ART_HIST_LEN(1, 1, 1),
         (long long)ART_PEER_BUFFER_BYTES(2), (long long)ART_PEER_BUFFER_BYTES(8), (double)ART_HIST_FIXED_ONE,
         ART_PEER_MAX_RANKS, ART_PEER_MAX_VARIANTS);
  return 0;
}
'''
    with tempfile.TemporaryDirectory() as td:
        cfile, exe = os.path.join(td, "m.c"), os.path.join(td, "m")
        open(cfile, "w").write(src)
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", exe, cfile], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert [int(out[0]), int(out[1])] == [_cabi.hist_len(64, 48, 128), _cabi.hist_len(1, 1, 1)]
    assert [int(out[2]), int(out[3])] == [_cabi.peer_buffer_bytes(2), _cabi.peer_buffer_bytes(8)]
    assert float(out[4]) == _cabi.HIST_FIXED_ONE
    assert [int(out[5]), int(out[6])] == [_cabi.PEER_MAX_RANKS, _cabi.PEER_MAX_VARIANTS]


def test_split_histogram_layout_and_units():
    """engine.split_histogram: the int64 vector of art_detector_histogram -> counts, fixed-point sums back to
    floating point, edges centred on the bounding-box midpoint, delays in fs relative to the mean path."""
    from attosecondraytracing_b200 import _cabi
    from attosecondraytracing_b200.engine import split_histogram
    nx, ny, nt = 3, 2, 4
    m = np.zeros(_cabi.MOMENTS_LEN)
    m[_cabi.M_N], m[_cabi.M_SD] = 4.0, 4.0 * 0.25          # mean d = 0.25 mm
    m[_cabi.M_XMIN], m[_cabi.M_XMAX], m[_cabi.M_YMIN], m[_cabi.M_YMAX] = -1.0, 2.0, 0.0, 4.0
    m[_cabi.M_DMIN], m[_cabi.M_DMAX] = 0.0, 1.0
    h = np.zeros(_cabi.hist_len(nx, ny, nt), dtype=np.int64)
    one = int(_cabi.HIST_FIXED_ONE)
    h[1 * ny + 1] = 4                       # all four rays in spot bin (1, 1)
    h[nx * ny + 1 * ny + 1] = 2 * one       # summed intensity 2.0
    h[2 * nx * ny + 1 * ny + 1] = one       # summed normalised delay 1.0 -> mean d = 0.25
    h[3 * nx * ny + 0] = 4
    h[3 * nx * ny + nt + 0] = 2 * one
    s = split_histogram(h, m, bins=(nx, ny), delay_bins=nt)
    assert s["spot_count"].shape == (nx, ny) and s["spot_count"][1, 1] == 4 and s["spot_count"].sum() == 4
    assert s["spot_intensity"][1, 1] == 2.0
    assert abs(s["spot_delay"][1, 1]) < 1e-12 and np.isnan(s["spot_delay"][0, 0])   # mean delay of the bin = mean path
    assert np.allclose(s["x_edges"], np.linspace(-1.5, 1.5, nx + 1)) and np.allclose(s["y_edges"], np.linspace(-2, 2, ny + 1))
    to_fs = 1e15 / 299792458000.0
    assert np.allclose(s["delay_edges"], (np.linspace(0, 1, nt + 1) - 0.25) * to_fs)
    assert s["delay_count"].tolist() == [4, 0, 0, 0] and s["delay_intensity"][0] == 2.0


def test_geometry_helpers_match_the_reference():
    """The small helpers of ModuleGeometry under the reference's names (point lists, ray lists as RayBundle and
    as list[Ray], root filters) against tests/golden/geometry.npz written by the unmodified reference."""
    import attosecondraytracing_b200.ModuleGeometry as mg
    from attosecondraytracing_b200.ModuleOpticalRay import RayBundle
    from golden_util import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "geometry.npz"))
    A, u, P, n, I1, I2, T, pts2, pts3 = (z[k] for k in ("A", "u", "P", "n", "I1", "I2", "T", "pts2", "pts3"))
    tol = dict(rtol=0, atol=1e-14)
    assert np.allclose(mg.IntersectionLinePlane(A, u, P, n), z["ilp"], **tol)
    assert np.allclose(np.sort(mg.SolverQuadratic(2.0, -3.0, -7.0)), z["quad"], **tol)
    assert np.allclose(np.sort(mg.SolverQuartic(1.0, 0.5, -5.0, 0.25, 3.0)), z["quart"], rtol=0, atol=1e-12)
    assert mg.SolverQuadratic(1.0, 0.0, 1.0) == []
    assert mg.KeepPositiveSolution([-1.0, 1e-13, 2.0]) == [2.0] and mg.KeepNegativeSolution([-1.0, -1e-13, 2.0]) == [-1.0]
    assert np.array_equal(mg.ClosestPoint(A, I1, I2), z["closest"]) and np.array_equal(mg.FarestPoint(A, I1, I2), z["farest"])
    assert mg.DiameterPointList(pts2) == float(z["diam2"]) and mg.DiameterPointList(pts3) == float(z["diam3"])
    assert mg.DiameterPointList([]) is None
    assert np.allclose(mg.CentrePointList(pts2), z["centre2"], **tol)
    assert np.allclose(mg.SymmetricalVector(u, n), z["symm"], **tol)
    assert np.allclose(mg.RotationPointList(pts3, u, n), z["rotpl"], **tol)
    assert np.allclose(mg.TranslationPointList(pts3, T), z["trpl"], **tol)
    bundle = RayBundle.from_numpy(z["ray_P"], z["ray_U"])
    for name, args in (("TranslationRayList", (T,)), ("RotationRayList", (u, n)), ("RotationAroundAxisRayList", (n, 0.7))):
        got = getattr(mg, name)(bundle, *args)
        assert isinstance(got, RayBundle) and got is not bundle
        d = got.to_numpy()
        assert np.allclose(d["P"], z[name + "_P"], **tol) and np.allclose(d["U"], z[name + "_U"], **tol), name
        as_list = getattr(mg, name)(list(bundle)[:20], *args)
        assert np.allclose(np.array([r.point for r in as_list]), z[name + "_P"][:20], **tol), name
        assert np.allclose(np.array([r.vector for r in as_list]), z[name + "_U"][:20], **tol), name
    assert np.array_equal(bundle.to_numpy()["P"], z["ray_P"])  # the input bundle is left untouched


def test_plane_wave_square_builds_the_intended_grid():
    """ModuleSource.PlaneWaveSquare: the reference's function raises (ART/ModuleSource.py:202 compares arrays);
    here it builds what that code evidently intends -- a central ray plus the off-axis grid points."""
    import attosecondraytracing_b200.ModuleSource as msrc
    centre, axis = np.array([1.0, -2.0, 3.0]), np.array([1.0, 0.0, 0.0])
    b = msrc.PlaneWaveSquare(centre, axis, 10.0, 50)          # 7 x 7 grid, its middle row / column on the axes
    d = b.to_numpy()
    assert len(b) == 1 + 6 * 6 and np.array_equal(d["number"], np.arange(37))
    assert np.allclose(d["U"], axis) and np.allclose(d["P"][0], centre)
    rel = d["P"] - centre
    assert np.allclose(rel[:, 0], 0.0, atol=1e-14)             # the grid lies in the plane normal to the axis
    assert np.isclose(np.abs(rel[1:, 1:]).max(), 5.0) and np.abs(rel[1:, 1:]).min() > 1e-4
    assert len(msrc.PlaneWaveSquare(centre, axis, 10.0, 16)) == 17   # 4 x 4 grid, no point on the axes


def test_driver_defaults_and_error_behaviour():
    """attosecondraytracing_b200.ARTmain (the reference's ARTmain.py without plotting): option completion and
    the errors of setup_detector / main, which need no GPU."""
    from attosecondraytracing_b200 import ARTmain
    from attosecondraytracing_b200.DefaultOptions import DefaultDetectorOptions, DefaultSourceProperties
    src, det, ana = ARTmain.complete_defaults({"NumberRays": 7}, {"DistanceDetector": 50.0}, {"verbose": False})
    assert src["NumberRays"] == 7 and src["Wavelength"] == 50e-6 and det["ReflectionNumber"] == -1
    assert det["DistanceDetector"] == 50.0 and ana["verbose"] is False and ana["save_results"] is True
    assert DefaultSourceProperties["NumberRays"] == 1000 and DefaultDetectorOptions["DistanceDetector"] is None

    class _Chain:  # setup_detector only looks at the element positions
        class _E:
            position = np.zeros(3)
        optical_elements = [_E()]

    with pytest.raises(RuntimeError, match="DetectorCentre"):
        ARTmain.setup_detector(_Chain(), dict(det, ManualDetector=True))
    with pytest.raises(RuntimeError, match="DetectorNormal"):
        ARTmain.setup_detector(_Chain(), dict(det, ManualDetector=True, DetectorCentre=np.ones(3)))
    with pytest.raises(RuntimeError, match="DistanceDetector"):
        ARTmain.setup_detector(_Chain(), dict(det, DistanceDetector=None))
    with pytest.raises(RuntimeError, match="RayList"):
        ARTmain.setup_detector(_Chain(), det, None)
    manual = ARTmain.setup_detector(_Chain(), dict(det, ManualDetector=True, DetectorCentre=np.array([0.0, 0.0, 5.0]),
                                                   DetectorNormal=np.array([0.0, 0.0, -1.0])))
    assert abs(manual.get_distance() - 5.0) < 1e-12
    with pytest.raises(ValueError, match="neither an OpticalChain"):
        ARTmain.main("not a chain", {}, {}, {})
