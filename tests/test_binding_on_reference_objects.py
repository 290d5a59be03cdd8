"""The reference-side binding (integration/art_b200_binding.py) on the reference's OWN objects.

The binding duck-types ART's classes; here it is fed real ones: the unmodified reference is imported (from
/root/reference in the build container, else from the travelling copy oracle/_ref), its OEPlacement builds the
chain, and
  * `lower()` must produce, field by field, the element / Zernike descriptors the package lowers for the same scene
    (the descriptors are all the library ever sees of a scene), and
  * `make_RayTracingCalculation(Ray)` must turn the library's columns back into the list[list[Ray]] the reference
    returns -- checked with the numpy oracle standing in for libart_b200.so (no GPU here), against the fixtures the
    reference itself produced.
Skipped where no copy of the reference is present."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integration"))

import art_oracle as orc  # noqa: E402
import ref_runner  # noqa: E402
import scenes as sc  # noqa: E402
from golden_util import Golden, golden_optical_elements  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="no copy of the reference on this machine")

SCENES = ["cfg3_2tor", "cfg1_par", "cfg5_tele", "sph_zern2", "mask_rrh_plane", "ell_offaxis", "cyl_cx", "sph_recthole"]


def _fixture_name(scene):
    return scene + ("_def" if scene == "sph_zern2" else "")


@pytest.mark.parametrize("scene", SCENES)
def test_lowering_of_reference_objects_equals_the_package_lowering(scene):
    import art_b200_binding as b200
    from attosecondraytracing_b200._lowering import LoweredChain
    chain = ref_runner.build_chain(sc.resolve(scene))
    els, zd, nz, keep = b200.lower(chain.optical_elements)
    g = Golden(_fixture_name(scene))
    low = LoweredChain([golden_optical_elements(g)])
    assert len(els) == low.n_elements and nz == low.n_defects
    for k in range(low.n_elements):
        a, b = els[k], low.elements[k]
        assert (a.surface, a.support, a.n_defects, a.first_defect) == (b.surface, b.support, b.n_defects, b.first_defect)
        assert a.n_gridmaps == 0 and b.n_gridmaps == 0
        for field, tol in (("surface_params", 1e-12), ("support_params", 0.0), ("centre", 1e-12), ("position", 0.0),
                           ("normal", 0.0), ("majoraxis", 0.0)):
            va, vb = np.array(getattr(a, field)[:]), np.array(getattr(b, field)[:])
            assert np.max(np.abs(va - vb)) <= tol * max(1.0, np.max(np.abs(vb))), (scene, k, field, va, vb)
    for i in range(nz):
        a, b = zd[i], low.defects[i]
        assert a.radius == b.radius and a.n_coefficients == b.n_coefficients
        ca = {(a.n[j], a.m[j]): a.c[j] for j in range(a.n_coefficients)}
        cb = {(b.n[j], b.m[j]): b.c[j] for j in range(b.n_coefficients)}
        assert ca == cb


@pytest.mark.parametrize("scene", ["cfg3_2tor", "cfg1_par", "sph_zern2"])
def test_ray_lists_are_rebuilt_as_the_reference_returns_them(scene, monkeypatch):
    import art_b200_binding as b200
    R = ref_runner.ref()
    g = Golden(_fixture_name(scene))
    chain = ref_runner.build_chain(sc.resolve(scene))
    oracle_els = g.oracle_elements()

    def fake_trace_columns(P, U, optical_elements, IgnoreDefects=True, path=None):
        # the numpy oracle in the role of art_trace_host: same column layout the binding gets from the library
        assert len(optical_elements) == len(oracle_els)
        traced = orc.trace_chain(P, U, oracle_els, ignore_defects=IgnoreDefects)
        n = P.shape[0]
        res = []
        for t in traced:
            alive = np.zeros(n, bool)
            alive[t["index"]] = True
            full = {"alive": alive}
            for key, width in (("P", 3), ("U", 3)):
                arr = np.full((n, width), np.nan)
                arr[t["index"]] = t[key]
                full[key] = arr
            for key, src in (("path", "path"), ("incidence", "incidence")):
                arr = np.full(n, np.nan)
                arr[t["index"]] = t[src] + (path[t["index"]] if (key == "path" and path is not None) else 0.0)
                full[key] = arr
            res.append(full)
        return res

    monkeypatch.setattr(b200, "trace_columns", fake_trace_columns)
    rtc = b200.make_RayTracingCalculation(R.mray.Ray)
    out = rtc(chain.source_rays, chain.optical_elements, IgnoreDefects=g.ignore_defects)
    assert isinstance(out, list) and len(out) == g.n_elements
    for k, rays in enumerate(out):
        ref = g.out(k)
        assert all(type(r) is R.mray.Ray for r in rays)
        assert [r.number for r in rays] == list(ref["num"])
        if rays:
            P = np.array([r.point for r in rays])
            assert np.max(np.abs(P - ref["P"])) <= 1e-9
            assert np.max(np.abs(np.array([np.sum(r.path) for r in rays]) - ref["path"])) <= 2e-9
            assert np.max(np.abs(np.array([r.incidence for r in rays]) - ref["inc"])) <= 1e-9
            assert all(r.wavelength == chain.source_rays[0].wavelength for r in rays)
    # the reference's own consumers work on the rebuilt lists: Detector.autoplace + GetResultSummary
    final = out[-1]
    det = R.mdet.Detector(chain.optical_elements[-1].position)
    det.autoplace(final, g.spec["detector_distance"])
    import load_reference as lr
    with lr.quiet():
        sd, dur = R.mplots.GetResultSummary(det, final)
    assert abs(sd - g["SpotSizeSD"]) <= 1e-9 and abs(dur - g["DurationSD"]) <= 1e-5
