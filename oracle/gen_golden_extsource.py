"""TEST INFRASTRUCTURE ONLY -- ExtendedSource of the UNMODIFIED reference (ART/ModuleSource.py:85-131) with
its Gaussian intensities, written to tests/golden/extsource.npz.   python oracle/gen_golden_extsource.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402

CASES = {"a": dict(Diameter=0.2, Divergence=0.02, NbRays=20000), "b": dict(Diameter=0.05, Divergence=0.004, NbRays=9100)}


def main():
    R = gg.ref()
    out = {}
    for key, c in CASES.items():
        rays = R.msource.ExtendedSource(np.array([0, 0, 0]), np.array([1, 0, 0]), c["Diameter"], c["Divergence"],
                                        c["NbRays"], Wavelength=800e-6)
        rays = R.msource.ApplyGaussianIntensityToRayList(rays, 1 / np.e**2)
        b = gg.bundle_arrays(rays)
        for k in ("num", "P", "U", "I"):
            out[f"{key}_{k}"] = b[k]
        out[f"{key}_params"] = np.array([c["Diameter"], c["Divergence"], c["NbRays"]])
        print(key, len(rays))
    np.savez_compressed(os.path.join(gg.GOLDEN_DIR, "extsource.npz"), **out)


if __name__ == "__main__":
    main()
