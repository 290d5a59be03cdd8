"""TEST INFRASTRUCTURE ONLY -- the small helpers of the UNMODIFIED reference's ART/ModuleGeometry.py (point-list
and ray-list translations / rotations, root filters, ...) on fixed inputs, written to tests/golden/geometry.npz.
    python oracle/gen_golden_geometry.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402


def main():
    R = gg.ref()
    g = R.mgeo
    rng = np.random.default_rng(20261018)
    A, u, P, n, I1, I2, T = rng.normal(size=(7, 3))
    pts2, pts3 = rng.normal(size=(50, 2)), rng.normal(size=(40, 3))
    out = dict(A=A, u=u, P=P, n=n, I1=I1, I2=I2, T=T, pts2=pts2, pts3=pts3)
    out["ilp"] = g.IntersectionLinePlane(A, u, P, n)
    out["quad"] = np.sort(g.SolverQuadratic(2.0, -3.0, -7.0))
    out["quart"] = np.sort(g.SolverQuartic(1.0, 0.5, -5.0, 0.25, 3.0))
    out["closest"] = g.ClosestPoint(A, I1, I2)
    out["farest"] = g.FarestPoint(A, I1, I2)
    out["diam2"] = np.array(g.DiameterPointList(list(pts2)))
    out["diam3"] = np.array(g.DiameterPointList(list(pts3)))
    out["centre2"] = np.array(g.CentrePointList(list(pts2)))
    out["symm"] = g.SymmetricalVector(u, n)
    out["rotpl"] = np.array(g.RotationPointList(list(pts3), u, n))
    out["trpl"] = np.array(g.TranslationPointList(list(pts3), T))
    rays = R.msource.PointSource(np.array([1.0, 2.0, 3.0]), np.array([0.2, 0.5, 1.0]), 0.05, 200, Wavelength=800e-6)
    b = gg.bundle_arrays(rays)
    out["ray_P"], out["ray_U"] = b["P"], b["U"]
    for name, args in (("TranslationRayList", (T,)), ("RotationRayList", (u, n)), ("RotationAroundAxisRayList", (n, 0.7))):
        res = gg.bundle_arrays(getattr(g, name)(rays, *args))
        out[name + "_P"], out[name + "_U"] = res["P"], res["U"]
    # MirrorProjection (ART/ModuleAnalysisAndPlots.py:470-483): impact points of the bundle after an element in
    # that element's support frame, from the reference's own ray-list transforms, on the cfg3 scene
    import scenes as sc
    import load_reference as lr
    chain = gg.build_chain(sc.resolve("cfg3_2tor"))
    with lr.quiet():
        outs = chain.get_output_rays()
    for k in (0, 1, 2):
        oe = chain.optical_elements[k]
        rl = g.TranslationRayList(outs[k], -oe.position)
        rl = g.RotationRayList(rl, oe.normal, np.array([0, 0, 1]))
        mp_ = g.RotationPoint(oe.majoraxis, oe.normal, np.array([0, 0, 1]))
        rl = g.RotationRayList(rl, mp_, np.array([1, 0, 0]))
        out[f"mproj{k}_xy"] = np.array([[r.point[0], r.point[1]] for r in rl])
        out[f"mproj{k}_incdeg"] = np.array([np.rad2deg(r.incidence) for r in outs[k]])
        out[f"mproj{k}_intensity"] = np.array([r.intensity for r in outs[k]])
    np.savez_compressed(os.path.join(gg.GOLDEN_DIR, "geometry.npz"), **out)
    print("written", len(out), "arrays")


if __name__ == "__main__":
    main()
