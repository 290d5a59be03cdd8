"""
TEST / BENCH INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference (ART v0.93, imported through
oracle/refshim from /root/reference or from the travelling copy oracle/_ref) on the scenes of
oracle/scenes.py, using nothing but the reference's own classes:

    scene spec -> reference optics (ModuleMirror / ModuleMask / ModuleSupport / ModuleDefects)
               -> mp.OEPlacement (ART/ModuleProcessing.py:133)        the aligned OpticalChain
               -> mp.RayTracingCalculation (ART/ModuleProcessing.py:250)
               -> Detector.autoplace + mplots.GetResultSummary (ARTmain.py:248-290 run_ART)

Used by oracle/gen_golden*.py (fixtures) and by bench.py's CPU legs (`--impl reference`,
`cpu_baseline.kind = "reference"`).  Imports nothing from attosecondraytracing_b200 or tests/.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(HERE, "refshim"), HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import load_reference as lr  # noqa: E402
import scenes as sc  # noqa: E402

_REF = None


def ref(root=None):
    """The reference's modules (loaded once per process)."""
    global _REF
    if _REF is None:
        _REF = lr.load(root)
    return _REF


def available():
    return lr.available()


# ------------------------------------------------------------------------------------------
# scene spec -> reference objects
# ------------------------------------------------------------------------------------------
def build_support(spec):
    ms = ref().msupp
    kind, p = spec[0], spec[1:]
    return {
        "round": ms.SupportRound, "roundhole": ms.SupportRoundHole, "rect": ms.SupportRectangle,
        "recthole": ms.SupportRectangleHole, "rectrecthole": ms.SupportRectangleRectHole,
    }[kind](*p)


def build_optic(spec):
    R = ref()
    mm = R.mmirror
    sup = build_support(spec["support"])
    k = spec["kind"]
    if k == "mask":
        return R.mmask.Mask(sup)
    if k == "plane":
        m = mm.MirrorPlane(sup)
    elif k == "spherical":
        m = mm.MirrorSpherical(spec["radius_signed"], sup)
    elif k == "cylindrical":
        m = mm.MirrorCylindrical(spec["radius_signed"], sup)
    elif k == "parabolic":
        m = mm.MirrorParabolic(spec["feff"], spec["offaxisangle_deg"], sup)
    elif k == "toroidal":
        m = mm.MirrorToroidal(spec["majorradius"], spec["minorradius"], sup)
    elif k == "ellipsoidal":
        kw = {a: spec[a] for a in ("SemiMajorAxis", "SemiMinorAxis", "OffAxisAngle", "f_object", "f_image")
              if a in spec}
        m = mm.MirrorEllipsoidal(sup, **kw)
    else:
        raise ValueError(k)
    if spec.get("defects"):
        dl = []
        for d in spec["defects"]:
            if d["kind"] == "zernike":
                coeffs = {(int(n), int(mm_)): c for n, mm_, c in d["coefficients"]}
                dl.append(R.mdef.Zernike(sup, coeffs))
            elif d["kind"] == "measuredmap":
                dl.append(R.mdef.MeasuredMap(sup, sc.measured_map(d["nx"], d["ny"], d["amplitude"])))
            elif d["kind"] == "fourier":
                np.random.seed(d["seed"])  # the reference draws its phases from the global numpy RNG
                dl.append(R.mdef.Fourrier(sup, d["rms"], slope=d["slope"], smallest=d["smallest"]))
            else:
                raise ValueError(d["kind"])
        m = mm.DeformedMirror(m, dl)
    return m


def build_chain(scene):
    """The reference's OpticalChain for a scene: its own OEPlacement, then the scene's misalignments."""
    R = ref()
    optics = [build_optic(s) for s in scene["optics"]]
    with lr.quiet():
        chain = R.mp.OEPlacement(dict(scene["source"]), optics, list(scene["distances"]),
                                 list(scene["incidences"]), list(scene["plane_angles"]), scene["name"])
    for op in scene.get("post", []):
        getattr(chain.optical_elements[op["element"]], op["op"])(op["value"])
    return chain


# ------------------------------------------------------------------------------------------
# rows of the full-size synthetic bundle as reference Ray objects
# ------------------------------------------------------------------------------------------
def spiral_rows(n_total, radius, k):
    """Rows k of SpiralVogel(n_total, radius), ART/ModuleGeometry.py:61-76 (same arithmetic, selected rows)."""
    golden_angle = np.pi * (3 - np.sqrt(5))
    k = np.asarray(k, dtype=np.float64)
    theta = golden_angle * k
    r = np.sqrt(k / n_total) * radius
    return np.stack([r * np.cos(theta), r * np.sin(theta)], axis=1)


def subset_source_rays(scene, n_full, idx, intensities=None):
    """Reference Ray objects for rays `idx` of the n_full-ray synthetic bundle of the scene's source.

    Built with the reference's own constructors / rotation (ART/ModuleSource.py:23-81, 135-169); only the
    Vogel-spiral row is evaluated per index instead of for all n_full rays.  `intensities`: per-ray values to
    attach (they take their normalisation from the FULL bundle, which the caller knows), else 1."""
    R = ref()
    sp = dict(scene["source"])
    ez = np.array([0, 0, 1])
    axis = np.array([1, 0, 0])
    rays = []
    if sp["Divergence"] == 0:
        xy = spiral_rows(n_full, sp["SourceSize"] / 2, idx)
        for (x, y), k in zip(xy, idx):
            rays.append(R.mray.Ray(np.array([x, y, 0]), np.array([0, 0, 1]), Number=int(k),
                                   Wavelength=sp["Wavelength"]))
    else:
        xy = spiral_rows(n_full, 1 * np.tan(sp["Divergence"]), idx)
        for (x, y), k in zip(xy, idx):
            rays.append(R.mray.Ray(np.array([0, 0, 0]), np.array([x, y, 1]), Number=int(k),
                                   Wavelength=sp["Wavelength"]))
    rays = R.mgeo.RotationRayList(rays, ez, axis)
    rays = R.mgeo.TranslationRayList(rays, np.array([0, 0, 0]))
    for j, r in enumerate(rays):
        r.intensity = np.float64(1.0 if intensities is None else intensities[j])
    return rays


# ------------------------------------------------------------------------------------------
# the timed unit of the CPU legs: what run_ART does for one chain (ARTmain.py:248-290)
# ------------------------------------------------------------------------------------------
def run_path(optical_elements, rays, detector_distance):
    """RayTracingCalculation + Detector.autoplace + GetResultSummary on `rays`.
    Returns (interactions, seconds, (SpotSizeSD, DurationSD) or None)."""
    R = ref()
    t0 = time.perf_counter()
    with lr.quiet():
        out = R.mp.RayTracingCalculation(rays, optical_elements, IgnoreDefects=True)
        final = out[-1]
        summary = None
        if len(final) > 1:
            det = R.mdet.Detector(optical_elements[-1].position)
            det.autoplace(final, detector_distance)
            summary = R.mplots.GetResultSummary(det, final)
    dt = time.perf_counter() - t0
    inter, entering = 0, len(rays)
    for o in out:
        inter += entering
        entering = len(o)
    return inter, dt, summary


_WORKER = {}


def _worker_init(scene_name, n_full, root):
    ref(root)
    scene = sc.resolve(scene_name)
    _WORKER["scene"] = scene
    _WORKER["chain"] = build_chain(scene)
    _WORKER["n_full"] = n_full


def _worker_run(idx):
    scene = _WORKER["scene"]
    rays = subset_source_rays(scene, _WORKER["n_full"], idx)
    inter, dt, _ = run_path(_WORKER["chain"].optical_elements, rays, scene["detector_distance"])
    return inter, dt


class ReferencePool:
    """`workers` processes, each holding the reference's aligned chain of one scene; `step(sample)` traces
    `sample` rays of the n_full-ray synthetic bundle (evenly spaced spiral indices, split over the workers) and
    returns (interactions, seconds of the slowest worker).  The reference is single-threaded and rays are
    independent, so independent worker processes are its honest multi-core use."""

    def __init__(self, scene_name, n_full, workers, root=None):
        import multiprocessing as mpc
        self.workers = workers
        self.scene = sc.resolve(scene_name)
        self.n_src = n_full - 1 if self.scene["source"]["Divergence"] == 0 else n_full
        self.pool = mpc.get_context("fork").Pool(workers, initializer=_worker_init, initargs=(scene_name, n_full, root))

    def step(self, sample):
        idx = np.linspace(0, self.n_src - 1, max(sample, self.workers)).astype(np.int64)
        chunks = [c for c in np.array_split(idx, self.workers) if c.size]
        res = self.pool.map(_worker_run, chunks, chunksize=1)
        return sum(r[0] for r in res), max(r[1] for r in res)

    def close(self):
        self.pool.close()
        self.pool.join()
