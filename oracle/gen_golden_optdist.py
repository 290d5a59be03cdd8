"""
TEST INFRASTRUCTURE ONLY -- golden values of the reference's detector-distance optimiser
(ART/ModuleProcessing.py:369-460 FindOptimalDistance), produced by running the UNMODIFIED reference on
the final bundles of some scenes of oracle/scenes.py.  Writes tests/golden/optdist.npz.

    python oracle/gen_golden_optdist.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, HERE)

import load_reference as lr  # noqa: E402
import scenes as sc  # noqa: E402
import gen_golden as gg  # noqa: E402

SCENES = ["cfg1_par", "cfg2_tor2f", "cfg3_2tor", "cfg5_tele", "ell_ab", "tele_pitch"]
CASES = [("intensity", False), ("duration", False), ("intensity", True)]


def main():
    R = gg.ref()
    out = {}
    meta = {}
    for name in SCENES:
        scene = sc.resolve(name)
        chain = gg.build_chain(scene)
        with lr.quiet():
            rays = R.mp.RayTracingCalculation(chain.source_rays, chain.optical_elements)[-1]
        det = R.mdet.Detector(chain.optical_elements[-1].position)
        det.autoplace(rays, scene["detector_distance"])
        for opt_for, weighted in CASES:
            with lr.quiet():
                d2, spot, dur = R.mp.FindOptimalDistance(det, rays, OptFor=opt_for, Amplitude=None, Precision=3,
                                                         IntensityWeighted=weighted)
            key = f"{name}__{opt_for}__{'w' if weighted else 'u'}"
            out[key] = np.array([d2.get_distance(), spot, dur], dtype=np.float64)
            out[key + "__centre"] = np.asarray(d2.centre, dtype=np.float64)
            print(key, out[key])
        meta[name] = {"first_distance": float(det.get_distance())}
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(gg.GOLDEN_DIR, "optdist.npz"), **out)


if __name__ == "__main__":
    main()
