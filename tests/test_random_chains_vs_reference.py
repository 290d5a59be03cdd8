"""Seeded random CHAINS against the LIVE reference: two to four random optics (every surface class, a mask, hole
supports) placed by the reference's own OEPlacement at random distances / incidences / plane angles and then
misaligned at random; a point-source bundle is traced by the reference's RayTracingCalculation and by the device
code compiled for the host (the kernel's element loop with its fused element-to-element hand-over) on the poses the
reference produced.  After every element the survivor sets are identical; the extended-precision arbiter
(oracle/art_oracle_ld.py) decides the numbers: the device code stays within 1e-10 mm (points, paths) of it on every
scene, the reference within 1e-8 mm (its np.roots toroid quartic and lab-frame round trips reach a few 1e-9 mm on
these 2 m chains).  Complements the fixed scenes of tests/golden (arbitrary relative poses stress the composed
hand-over map) and tests/test_adversarial_vs_reference.py (single optics, arbitrary rays).
Runs where a copy of the reference is present (build container: /root/reference; elsewhere oracle/_ref)."""
import numpy as np
import pytest

import art_oracle_ld as ld
import gen_golden
import hostcheck_util
import ref_runner
from attosecondraytracing_b200 import _cabi
from attosecondraytracing_b200._lowering import LoweredChain
from golden_util import build_optic

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="no copy of the reference on this machine")

TOR = (5585.122305476701, 173.64817766693042)   # ReturnOptimalToroidalRadii(500, 80 deg)
POOL = [
    # optic spec, incidence range in degrees
    ({"kind": "toroidal", "majorradius": TOR[0], "minorradius": TOR[1], "support": ["rect", 300, 60]}, (70, 82)),
    ({"kind": "spherical", "radius_signed": 2500.0, "support": ["round", 40]}, (2, 25)),
    ({"kind": "spherical", "radius_signed": -1500.0, "support": ["roundhole", 45, 3, 20, -15]}, (2, 25)),
    ({"kind": "parabolic", "feff": 400.0, "offaxisangle_deg": 30.0, "support": ["rect", 90, 90]}, (0, 0)),
    ({"kind": "ellipsoidal", "SemiMajorAxis": 1000.0, "SemiMinorAxis": 173.64817766693042,
      "support": ["recthole", 260, 60, 2, 70, 12]}, (75, 82)),
    ({"kind": "cylindrical", "radius_signed": 3000.0, "support": ["rectrecthole", 120, 90, 6, 4, 35, -25]}, (5, 40)),
    ({"kind": "plane", "support": ["round", 50]}, (5, 60)),
    ({"kind": "mask", "support": ["roundhole", 30, 12, 0, 0]}, (0, 0)),
]
MISALIGN = ["rotate_pitch_by", "rotate_roll_by", "rotate_yaw_by", "shift_along_normal", "shift_along_major",
            "shift_along_cross"]


def _scene(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 5))
    picks = [POOL[i] for i in rng.choice(len(POOL), n, replace=True)]
    if all(p[0]["kind"] == "mask" for p in picks):
        picks[-1] = POOL[1]
    optics = [dict(p[0]) for p in picks]
    incid = [float(rng.uniform(*p[1])) * (1 if rng.random() < 0.5 else -1) for p in picks]
    post = []
    for k in range(n):
        for _ in range(int(rng.integers(0, 3))):
            op = MISALIGN[int(rng.integers(len(MISALIGN)))]
            post.append({"element": k, "op": op, "value": float(rng.normal(0.0, 0.02 if op.startswith("rotate") else 0.3))})
    return {
        "name": f"random{seed}",
        "source": {"Divergence": float(rng.uniform(5e-3, 70e-3)), "SourceSize": 0, "Wavelength": 800e-6, "NumberRays": 400},
        "optics": optics,
        "distances": [float(rng.uniform(150, 700)) for _ in range(n)],
        "incidences": incid,
        "plane_angles": [float(rng.uniform(0, 360)) for _ in range(n)],
        "post": post,
    }


@pytest.mark.parametrize("seed", range(20))
def test_random_chain_matches_the_live_reference(seed):
    R = ref_runner.ref()
    import load_reference as lr
    import attosecondraytracing_b200.ModuleOpticalElement as moe
    scene = _scene(seed)
    chain = ref_runner.build_chain(scene)
    n = scene["source"]["NumberRays"]
    rays = ref_runner.subset_source_rays(scene, n, np.arange(n))
    src_P = np.array([r.point for r in rays], dtype=np.float64)
    src_U = np.array([r.vector for r in rays], dtype=np.float64)
    with lr.quiet():
        out = R.mp.RayTracingCalculation(rays, chain.optical_elements, IgnoreDefects=True)
    oes = []
    for spec, roe in zip(scene["optics"], chain.optical_elements):
        optic = build_optic(dict(spec, support=tuple(spec["support"])))
        oes.append(moe.OpticalElement(optic, np.asarray(roe.position, dtype=np.float64),
                                      np.asarray(roe.normal, dtype=np.float64), np.asarray(roe.majoraxis, dtype=np.float64)))
    dev = hostcheck_util.trace(LoweredChain([oes]), src_P, src_U, _cabi.TRACE_IGNORE_DEFECTS)
    kinds = "+".join(s["kind"] for s in scene["optics"])
    arb = None
    if ld.available():
        els = []
        for spec, roe in zip(scene["optics"], chain.optical_elements):
            optic = gen_golden.derived_optic(spec, roe.type)
            optic["support"] = tuple(optic["support"])
            els.append({"optic": optic, "position": np.asarray(roe.position, dtype=np.float64),
                        "normal": np.asarray(roe.normal, dtype=np.float64),
                        "majoraxis": np.asarray(roe.majoraxis, dtype=np.float64)})
        arb = ld.trace_chain(src_P, src_U, els, ignore_defects=True)
    for k, ref_list in enumerate(out):
        ref_num = np.array([r.number for r in ref_list], dtype=np.int64)
        got = np.nonzero(dev[k]["alive"])[0]
        assert np.array_equal(got, ref_num), (seed, kinds, k, np.setxor1d(got, ref_num)[:8])
        if ref_num.size == 0:
            continue
        ref_P = np.array([r.point for r in ref_list]).reshape(-1, 3)
        ref_U = np.array([r.vector for r in ref_list]).reshape(-1, 3)
        ref_L = np.array([np.sum(r.path) for r in ref_list], dtype=np.float64)
        # the reference against the device code: its own noise is the limit
        assert np.max(np.abs(dev[k]["P"][got] - ref_P)) <= 1e-8, (seed, kinds, k)
        assert np.max(np.abs(dev[k]["U"][got] - ref_U)) <= 1e-10, (seed, kinds, k)
        assert np.max(np.abs(dev[k]["path"][got] - ref_L)) <= 1e-8, (seed, kinds, k)
        if arb is not None:
            assert np.array_equal(arb[k]["number"], ref_num), (seed, kinds, k)
            d_dev = float(np.max(np.abs(dev[k]["P"][got] - arb[k]["P"])))
            d_path = float(np.max(np.abs(dev[k]["path"][got] - arb[k]["path"])))
            assert d_dev <= 1e-10 and d_path <= 2e-10, (seed, kinds, k, d_dev, d_path)
            assert float(np.max(np.abs(ref_P - arb[k]["P"]))) <= 1e-8, (seed, kinds, k)


DEFORMED = [
    ({"kind": "parabolic", "feff": 400.0, "offaxisangle_deg": 30.0, "support": ["rect", 90, 90]}, (0, 0)),
    ({"kind": "spherical", "radius_signed": 2500.0, "support": ["round", 40]}, (2, 25)),
    ({"kind": "plane", "support": ["round", 50]}, (5, 60)),
    ({"kind": "toroidal", "majorradius": TOR[0], "minorradius": TOR[1], "support": ["rect", 300, 60]}, (70, 82)),
]


def _deformed_scene(seed):
    """As _scene, with a Zernike-deformed mirror (random order 3..8, random coefficients) somewhere in the chain."""
    import scenes as sc
    rng = np.random.default_rng(1000 + seed)
    scene = _scene(1000 + seed)
    k = int(rng.integers(len(scene["optics"])))
    base, inc = DEFORMED[int(rng.integers(len(DEFORMED)))]
    optic = dict(base)
    optic["defects"] = [{"kind": "zernike",
                         "coefficients": sc.zernike_table(int(rng.integers(3, 9)), seed=seed, scale=float(rng.uniform(5e-5, 3e-4)))}]
    if rng.random() < 0.5:   # two stacked defects
        optic["defects"].append({"kind": "zernike", "coefficients": [[2, 1, 4e-5], [3, 0, -6e-5], [5, 4, 3e-5]]})
    scene["optics"][k] = optic
    scene["incidences"][k] = float(rng.uniform(*inc)) * (1 if rng.random() < 0.5 else -1)
    scene["source"]["Divergence"] = float(rng.uniform(5e-3, 30e-3))
    return scene


@pytest.mark.parametrize("ignore", [True, False])
@pytest.mark.parametrize("seed", range(8))
def test_random_chain_with_zernike_defects(seed, ignore):
    """Both IgnoreDefects modes of the reference (ART/ModuleMirror.py:945-980) on random chains with a deformed mirror."""
    R = ref_runner.ref()
    import load_reference as lr
    import attosecondraytracing_b200.ModuleOpticalElement as moe
    scene = _deformed_scene(seed)
    chain = ref_runner.build_chain(scene)
    n = scene["source"]["NumberRays"]
    rays = ref_runner.subset_source_rays(scene, n, np.arange(n))
    src_P = np.array([r.point for r in rays], dtype=np.float64)
    src_U = np.array([r.vector for r in rays], dtype=np.float64)
    with lr.quiet():
        out = R.mp.RayTracingCalculation(rays, chain.optical_elements, IgnoreDefects=ignore)
    oes, els = [], []
    for spec, roe in zip(scene["optics"], chain.optical_elements):
        pose = [np.asarray(getattr(roe, a), dtype=np.float64) for a in ("position", "normal", "majoraxis")]
        oes.append(moe.OpticalElement(build_optic(dict(spec, support=tuple(spec["support"]))), *pose))
        optic = gen_golden.derived_optic(spec, roe.type)
        optic["support"] = tuple(optic["support"])
        if optic.get("defects"):
            for dd in optic["defects"]:
                dd["coefficients"] = {(int(a), int(b)): c for a, b, c in dd["coefficients"]}
        els.append({"optic": optic, "position": pose[0], "normal": pose[1], "majoraxis": pose[2]})
    dev = hostcheck_util.trace(LoweredChain([oes]), src_P, src_U, _cabi.TRACE_IGNORE_DEFECTS if ignore else 0)
    arb = ld.trace_chain(src_P, src_U, els, ignore_defects=ignore) if ld.available() else None
    kinds = "+".join(s["kind"] + ("*" if s.get("defects") else "") for s in scene["optics"])
    for k, ref_list in enumerate(out):
        ref_num = np.array([r.number for r in ref_list], dtype=np.int64)
        got = np.nonzero(dev[k]["alive"])[0]
        assert np.array_equal(got, ref_num), (seed, kinds, k, np.setxor1d(got, ref_num)[:8])
        if ref_num.size == 0:
            continue
        ref_P = np.array([r.point for r in ref_list]).reshape(-1, 3)
        ref_U = np.array([r.vector for r in ref_list]).reshape(-1, 3)
        assert np.max(np.abs(dev[k]["P"][got] - ref_P)) <= 1e-8, (seed, kinds, k)
        assert np.max(np.abs(dev[k]["U"][got] - ref_U)) <= 1e-10, (seed, kinds, k)
        if arb is not None:
            assert np.array_equal(arb[k]["number"], ref_num), (seed, kinds, k)
            assert float(np.max(np.abs(dev[k]["P"][got] - arb[k]["P"]))) <= 1e-10, (seed, kinds, k)
