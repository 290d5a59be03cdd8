"""TEST INFRASTRUCTURE ONLY -- ReflectionMirrorRayList (ART/ModuleMirror.py:912-939, both IgnoreDefects values)
and TransmitMaskRayList (ART/ModuleMask.py:112-136) of the UNMODIFIED reference on rays given in the optic's own
frame, written to tests/golden/raylist.npz.   python oracle/gen_golden_raylist.py"""
import os
import sys

import json

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402
import load_reference as lr  # noqa: E402

# optic spec (oracle/scenes.py vocabulary), source point in the optic's frame, half-angle of the cone aimed
# at the optic's centre, number of rays
CASES = {
    "toroid": dict(optic={"kind": "toroidal", "majorradius": 5585.122305476701, "minorradius": 173.64817766693042,
                          "support": ("rect", 300, 50)}, S=[984.8, 0.0, -5585.1], div=0.02, n=600),
    "sphere_cx": dict(optic={"kind": "spherical", "radius_signed": -1500.0, "support": ("round", 25)},
                      S=[40.0, 10.0, -1100.0], div=0.06, n=500),
    "parabola_hole": dict(optic={"kind": "parabolic", "feff": 100.0, "offaxisangle_deg": 90.0,
                                 "support": ("roundhole", 30, 5, 10, 5)}, S=[95.0, 3.0, 400.0], div=0.12, n=700),
    "sphere_zernike": dict(optic={"kind": "spherical", "radius_signed": 800.0, "support": ("round", 20),
                                  "defects": [{"kind": "zernike", "coefficients": [[2, 0, 2e-4], [3, 1, -1e-4],
                                                                                   [4, 2, 5e-5], [6, 3, 3e-5]]}]},
                           S=[5.0, -3.0, -300.0], div=0.05, n=500),
    "mask": dict(optic={"kind": "mask", "support": ("roundhole", 20, 7, 0, 0)}, S=[0.5, -0.3, -400.0], div=0.07, n=800),
}


def main():
    R = gg.ref()
    out = {}
    for key, c in CASES.items():
        optic = gg.build_optic(c["optic"])
        S = np.array(c["S"], dtype=np.float64)
        axis = np.asarray(optic.get_centre(), dtype=np.float64) - S
        rays = R.msource.PointSource(S, axis / np.linalg.norm(axis), c["div"], c["n"], Wavelength=800e-6)
        b = gg.bundle_arrays(rays)
        out[f"{key}_src_P"], out[f"{key}_src_U"], out[f"{key}_src_num"] = b["P"], b["U"], b["num"]
        with lr.quiet():
            if c["optic"]["kind"] == "mask":
                res = {"out": R.mmask.TransmitMaskRayList(optic, rays)}
            else:
                res = {"out": R.mmirror.ReflectionMirrorRayList(optic, rays)}  # default IgnoreDefects=False
                if c["optic"].get("defects"):
                    res["outign"] = R.mmirror.ReflectionMirrorRayList(optic, rays, IgnoreDefects=True)
        for tag, lst in res.items():
            o = gg.bundle_arrays(lst)
            for k in ("num", "P", "U", "path", "inc"):
                out[f"{key}_{tag}_{k}"] = o[k]
            print(key, tag, len(rays), "->", len(lst))
    out["cases"] = np.array(json.dumps({k: dict(optic=c["optic"], S=c["S"], div=c["div"], n=c["n"]) for k, c in CASES.items()}))
    np.savez_compressed(os.path.join(gg.GOLDEN_DIR, "raylist.npz"), **out)


if __name__ == "__main__":
    main()
