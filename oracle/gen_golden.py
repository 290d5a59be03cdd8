"""
TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/ART through oracle/refshim) on the scenes of oracle/scenes.py.

    python oracle/gen_golden.py            # all scenes + scale subsets
    python oracle/gen_golden.py cfg3_2tor  # selected ones

Runs only in the build container (needs /root/reference).  The fixtures it writes are committed;
nothing at test/bench time on the GPU box reads the reference.

Each fixture holds: the scene spec (JSON), the element poses the reference's OEPlacement +
misalignment methods produced, derived optic parameters, the source bundle, the bundle after
EVERY element (number, P, U, sum(path), incidence), the autoplace'd detector and the
statistics (GetResultSummary, getETransmission, weighted SDs, per-ray centred 2-D points and
delays).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, HERE)

import load_reference as lr  # noqa: E402
import scenes as sc  # noqa: E402
import art_oracle as orc  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = None
EXTRA_ARRAYS = {}  # arrays of gridded defects collected by derived_optic for the fixture being written


def ref():
    import ref_runner
    return ref_runner.ref()


# scene spec -> reference objects: shared with bench.py's CPU legs
from ref_runner import build_chain, build_optic, build_support  # noqa: E402,F401


def derived_optic(spec, obj):
    """Oracle-side description of a reference optic object (derived parameters included)."""
    k = spec["kind"]
    base = obj.Mirror if type(obj).__name__ == "DeformedMirror" else obj
    d = {"kind": k, "support": list(spec["support"]), "type": obj.type}
    if k in ("spherical", "cylindrical"):
        d["radius"] = float(base.radius)
    elif k == "parabolic":
        d.update(feff=float(base.feff), offaxisangle=float(base.offaxisangle), p=float(base.p))
    elif k == "toroidal":
        d.update(majorradius=float(base.majorradius), minorradius=float(base.minorradius))
    elif k == "ellipsoidal":
        d.update(a=float(base.a), b=float(base.b), offaxisangle=float(base._offaxisangle))
    if spec.get("defects"):
        d["defects"] = []
        for i, (ds, dobj) in enumerate(zip(spec["defects"], obj.DeformationList)):
            if ds["kind"] == "zernike":
                d["defects"].append({"kind": "zernike", "R": float(dobj.R), "max_order": int(dobj.max_order),
                                     "coefficients": ds["coefficients"]})
            else:
                # gridded defect: the arrays the reference built go into the fixture (EXTRA_ARRAYS); the
                # interpolation grid is what its RegularGridInterpolators were given
                rect = dobj.Support._CircumRect() if hasattr(dobj, "Support") else base.support._CircumRect()
                h = np.asarray(dobj.deformation, dtype=np.float64)
                if ds["kind"] == "measuredmap":
                    X = np.linspace(-rect[0], rect[0], num=h.shape[0])
                    Y = np.linspace(-rect[1], rect[1], num=h.shape[1])
                else:
                    X = np.linspace(-rect[0] / 2, rect[0] / 2, num=h.shape[1])
                    Y = np.linspace(-rect[1] / 2, rect[1] / 2, num=h.shape[0])
                key = f"map{len(EXTRA_ARRAYS) // 3}"
                EXTRA_ARRAYS[key + "_h"] = np.transpose(h)
                EXTRA_ARRAYS[key + "_dx"] = np.transpose(np.asarray(dobj.DerivX, dtype=np.float64))
                EXTRA_ARRAYS[key + "_dy"] = np.transpose(np.asarray(dobj.DerivY, dtype=np.float64))
                d["defects"].append({"kind": "gridmap", "source": ds["kind"], "arrays": key,
                                     "x0": float(X[0]), "x1": float(X[-1]), "y0": float(Y[0]), "y1": float(Y[-1])})
    d["centre"] = [float(v) for v in obj.get_centre()]
    return d


def bundle_arrays(rays):
    return {
        "num": np.array([r.number for r in rays], dtype=np.int64),
        "P": np.array([r.point for r in rays], dtype=np.float64).reshape(-1, 3),
        "U": np.array([r.vector for r in rays], dtype=np.float64).reshape(-1, 3),
        "path": np.array([np.sum(r.path) for r in rays], dtype=np.float64),
        "inc": np.array([np.nan if r.incidence is None else r.incidence for r in rays], dtype=np.float64),
        "I": np.array([np.nan if r.intensity is None else r.intensity for r in rays], dtype=np.float64),
    }


def run_scene(scene, ignore_defects=True, source_rays=None, extra=None):
    R = ref()
    EXTRA_ARRAYS.clear()
    chain = build_chain(scene)
    if source_rays is not None:
        chain.source_rays = source_rays
    t0 = time.time()
    with lr.quiet():
        out = R.mp.RayTracingCalculation(chain.source_rays, chain.optical_elements, IgnoreDefects=ignore_defects)
    dt = time.time() - t0
    data = {}
    spec = dict(scene)
    spec["ignore_defects"] = bool(ignore_defects)
    spec["derived_optics"] = [derived_optic(s, oe.type) for s, oe in zip(scene["optics"], chain.optical_elements)]
    src = bundle_arrays(chain.source_rays)
    data["src_num"], data["src_P"], data["src_U"], data["src_I"] = src["num"], src["P"], src["U"], src["I"]
    for k, oe in enumerate(chain.optical_elements):
        data[f"el{k}_position"] = np.asarray(oe.position, dtype=np.float64)
        data[f"el{k}_normal"] = np.asarray(oe.normal, dtype=np.float64)
        data[f"el{k}_majoraxis"] = np.asarray(oe.majoraxis, dtype=np.float64)
        b = bundle_arrays(out[k])
        for key in ("num", "P", "U", "path", "inc"):
            data[f"out{k}_{key}"] = b[key]
    final = out[-1]
    ninter, entering = 0, len(chain.source_rays)
    for k in range(len(out)):
        ninter += entering
        entering = len(out[k])
    spec["interactions"] = ninter
    spec["reference_trace_seconds"] = dt
    if len(final) > 1:
        det = R.mdet.Detector(chain.optical_elements[-1].position)
        det.autoplace(final, scene["detector_distance"])
        with lr.quiet():
            sd, dur = R.mplots.GetResultSummary(det, final)
        xy = np.array(det.get_PointList2DCentre(final))
        delays = np.array(det.get_Delays(final))
        w = [r.intensity for r in final]
        data.update(
            det_centre=det.centre, det_normal=det.normal, det_refpoint=det.refpoint,
            det_xy_centre=xy, det_delays=delays,
            SpotSizeSD=np.float64(sd), DurationSD=np.float64(dur),
            ETransmission=np.float64(R.mplots.getETransmission(chain.source_rays, final)),
            SpotSizeSD_w=np.float64(R.mp.WeightedStandardDeviation(xy, w)),
            DurationSD_w=np.float64(R.mp.WeightedStandardDeviation(delays, w)),
            NA=np.float64(R.mp.ReturnNumericalAperture(final, 1)),
            Diameter=np.float64(R.mgeo.DiameterPointList(list(xy))),
            det_distance=np.float64(det.get_distance()),
        )
    if extra:
        spec.update(extra)
    data.update(EXTRA_ARRAYS)
    data["spec"] = np.array(json.dumps(spec))
    return data, [len(o) for o in out], dt


def save(name, data):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **data)
    return path


# ------------------------------------------------------------------------------------------
def subset_source_rays(scene, n_full, idx):
    """Reference Ray objects for rays `idx` of the n_full-ray synthetic bundle (ref_runner.subset_source_rays);
    intensities take their normalisation from the full bundle (oracle.source_for)."""
    import ref_runner
    sp = dict(scene["source"])
    sp["NumberRays"] = n_full
    _, _, _, inten = orc.source_for(sp, first_support=None, k=idx)
    return ref_runner.subset_source_rays(scene, n_full, idx, intensities=inten)


def gen_subsets(which):
    # M = 10^4 seeded indices per full-size bundle (SURVEY.md 8(d) "parity check at scale");
    # of the three cfg5 sweep variants the middle one carries 10^4 rays, the end points 1000 each (fixture size)
    todo = {
        "cfg2_sub": ("cfg2", 10000, [True]),
        "cfg3_sub": ("cfg3", 10000, [True]),
        "cfg4_sub": ("cfg4", 10000, [True, False]),
        "cfg5_sub": ("cfg5", 10000, [True]),
    }
    small_variant_rays = 1000
    for name, (wl, m, modes) in todo.items():
        if which and name not in which:
            continue
        w = sc.WORKLOADS[wl]
        scene = sc.resolve(w["scene"])
        n_full = w["rays"]
        n_src = n_full - 1 if scene["source"]["Divergence"] == 0 else n_full  # PlaneWaveDisk off-by-one
        variants = [None]
        if "sweep" in w:
            sw = w["sweep"]
            vals = np.linspace(sw["lo"], sw["hi"], sw["n"])
            variants = [(i, float(vals[i])) for i in (0, 300, 1023)]
        for var in variants:
            for ign in modes:
                scn = dict(scene)
                tag = name
                extra = {"subset_of": n_full, "workload": wl}
                if var is not None:
                    sw = w["sweep"]
                    scn["post"] = list(scn.get("post", [])) + [
                        {"op": f"rotate_{sw['axis']}_by", "element": sw["element"], "value": var[1]}]
                    tag += f"_v{var[0]}"
                    extra["variant_index"] = var[0]
                if len(modes) > 1:
                    tag += "_ign" if ign else "_def"
                m_here = small_variant_rays if (var is not None and var[0] != 300) else m
                idx = np.sort(np.random.default_rng(1234).choice(n_src, m_here, replace=False))
                rays = subset_source_rays(scn, n_full, idx)
                data, counts, dt = run_scene(scn, ignore_defects=ign, source_rays=rays, extra=extra)
                p = save(tag, data)
                print(f"{tag:28s} survivors {counts}  ref trace {dt:.2f}s -> {os.path.relpath(p)}")


def main(argv):
    which = set(argv)
    for name in sc.SCENES:
        if which and name not in which:
            continue
        if name == "par_measured":
            continue  # needs the numpy-2 compatibility copy of the reference: oracle/gen_golden_gridmap.py
        scene = sc.resolve(name)
        has_def = any(o.get("defects") for o in scene["optics"])
        # gridded defects: the reference's get_normal raises under numpy >= 2 (inhomogeneous list in
        # np.linalg.norm, ART/ModuleDefects.py:126), so only its default IgnoreDefects=True path can be run
        gridded = any(d["kind"] != "zernike" for o in scene["optics"] for d in o.get("defects", []))
        for ign in ([True] if (gridded or not has_def) else [True, False]):
            tag = name + (("_ign" if ign else "_def") if has_def else "")
            data, counts, dt = run_scene(scene, ignore_defects=ign)
            p = save(tag, data)
            print(f"{tag:28s} survivors {counts}  ref trace {dt:.2f}s -> {os.path.relpath(p)}")
    gen_subsets(which)


if __name__ == "__main__":
    main(sys.argv[1:])
