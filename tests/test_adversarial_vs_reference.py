"""Adversarial rays against the LIVE reference (not a fixture): seeded random origins inside and outside each solid,
random directions, every surface class + a mask, through the reference's own `ReflectionMirrorRayList` /
`TransmitMaskRayList` (rays in the optic's own frame) and through the device code compiled for the host -- survivor
sets must be identical, points within 1e-9 mm.  This is the root-selection logic under stress: rays that start
inside a sphere / toroid, see two admissible roots, graze the surface or run away from it.
Runs where a copy of the reference is present (build container: /root/reference; elsewhere oracle/_ref)."""
import numpy as np
import pytest

import hostcheck_util
import ref_runner
from attosecondraytracing_b200 import _cabi
from attosecondraytracing_b200._lowering import LoweredChain
from golden_util import build_optic

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="no copy of the reference on this machine")

TOR = (5585.122305476701, 173.64817766693042)
CASES = {
    # optic spec, box the origins are drawn from (element frame, mm), around which point directions are aimed
    "toroid": ({"kind": "toroidal", "majorradius": TOR[0], "minorradius": TOR[1], "support": ("rect", 300, 50)},
               [(-400, 400), (-120, 120), (-6100, -5300)], (0.0, 0.0, -TOR[0] - TOR[1])),
    "sphere_cc": ({"kind": "spherical", "radius_signed": 800.0, "support": ("round", 60)},
                  [(-300, 300), (-300, 300), (-1100, 300)], (0.0, 0.0, -800.0)),
    "sphere_cx": ({"kind": "spherical", "radius_signed": -800.0, "support": ("recthole", 120, 90, 10, 5, -8)},
                  [(-300, 300), (-300, 300), (-1100, 300)], (0.0, 0.0, -800.0)),
    "parabola": ({"kind": "parabolic", "feff": 100.0, "offaxisangle_deg": 60.0, "support": ("roundhole", 30, 5, 10, 5)},
                 [(-150, 250), (-150, 150), (-100, 400)], None),
    "ellipsoid": ({"kind": "ellipsoidal", "SemiMajorAxis": 1000.0, "SemiMinorAxis": 173.64817766693042,
                   "support": ("rect", 200, 40)}, [(-1100, 1100), (-200, 200), (-300, 100)], None),
    "cylinder": ({"kind": "cylindrical", "radius_signed": 500.0, "support": ("rectrecthole", 80, 60, 20, 10, 4, -3)},
                 [(-100, 100), (-300, 300), (-700, 200)], (0.0, 0.0, -500.0)),
    "plane": ({"kind": "plane", "support": ("round", 40)}, [(-100, 100), (-100, 100), (-200, 200)], (0.0, 0.0, 0.0)),
    "mask": ({"kind": "mask", "support": ("roundhole", 20, 7, 1, -2)}, [(-60, 60), (-60, 60), (-300, 300)], (0.0, 0.0, 0.0)),
}
N = 700


@pytest.mark.parametrize("name", sorted(CASES))
def test_random_rays_in_the_optics_own_frame(name):
    R = ref_runner.ref()
    import load_reference as lr
    spec, box, aim = CASES[name]
    rng = np.random.default_rng(abs(hash(name)) % 2**32 if False else sum(map(ord, name)))
    ref_optic = ref_runner.build_optic(spec)
    centre = np.asarray(ref_optic.get_centre(), dtype=np.float64) if aim is None else np.asarray(aim, dtype=np.float64)
    P = np.column_stack([rng.uniform(lo, hi, N) for lo, hi in box])
    # half of the rays aim near the optic's centre (hits, grazing hits), half fly in random directions
    target = centre + rng.normal(0.0, 25.0, (N, 3))
    U = target - P
    rnd = rng.normal(size=(N, 3))
    pick = rng.random(N) < 0.5
    U[pick] = rnd[pick]
    U /= np.linalg.norm(U, axis=1)[:, None]
    rays = [R.mray.Ray(P[i].copy(), U[i].copy(), Number=i, Wavelength=800e-6) for i in range(N)]
    with lr.quiet():
        if spec["kind"] == "mask":
            out = R.mmask.TransmitMaskRayList(ref_optic, rays)
        else:
            out = R.mmirror.ReflectionMirrorRayList(ref_optic, rays)
    ref_num = np.array([r.number for r in out], dtype=np.int64)
    ref_P = np.array([r.point for r in out]).reshape(-1, 3)
    ref_U = np.array([r.vector for r in out]).reshape(-1, 3)
    # the same optic in an element whose frame is the lab frame, through the device code
    import attosecondraytracing_b200.ModuleOpticalElement as moe
    optic = build_optic(dict(spec, support=tuple(spec["support"])))
    oe = moe.OpticalElement(optic, np.asarray(optic.get_centre(), dtype=np.float64), np.array([0.0, 0.0, 1.0]),
                            np.array([1.0, 0.0, 0.0]))
    res = hostcheck_util.trace(LoweredChain([[oe]]), P, U, 0)[0]
    got = np.nonzero(res["alive"])[0]
    assert np.array_equal(got, ref_num), (name, np.setxor1d(got, ref_num)[:10])
    assert 0 < got.size < N, f"{name}: the case should mix hits and misses ({got.size}/{N})"
    assert np.max(np.abs(res["P"][got] - ref_P)) <= 1e-9
    assert np.max(np.abs(res["U"][got] - ref_U)) <= 1e-10
