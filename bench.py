#!/usr/bin/env python
"""Benchmark of the ray-bundle hot path (BASELINE.json metric: ray-element interactions/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg4|cfg4def|cfg5]
                    [--sub cfg2,cfg4,cfg4def,cfg5|none] [--impl reference]

One "step" = one pass of the hot path over one synthetic bundle: trace through every element of the
chain (fused kernel), all-reduce + autoplace of the detector, detector response + moments.

The headline workload is BASELINE config 3 -- the north star's 2-toroid chain
(examples/CONFIG_2toroidals_f-x-f.py: mask + two toroids), 12.5 M rays per GPU, i.e. BASELINE's 100 M rays
on 8 GPUs; weak scaling: every rank traces its own round-robin share of an N x 12.5 M-ray bundle and the
only exchanges are the central sums and the moments rows.  The other BASELINE configs are measured in the
same run, device-timed the same way, and attached as `workloads`: cfg2 (one toroid, 10 M rays / GPU), cfg4
(Zernike order 20, 50 M rays / GPU, both IgnoreDefects modes) and cfg5 (1024 telescope variants x 1 M rays,
the VARIANT axis sharded over the GPUs: strong scaling).

Prints ONE JSON line (rank 0).  `--impl reference` times the UNMODIFIED reference (oracle/_ref, a
byte-for-byte copy of ART v0.93 made by oracle/make_ref.py, driven through its own OEPlacement /
RayTracingCalculation / Detector.autoplace / GetResultSummary) on all host cores on bounded samples of the
same workload; that arm imports nothing from attosecondraytracing_b200.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "ray-element interactions/s"
UNIT = "interactions/s"
DEFAULT_SUBS = "cfg2,cfg4,cfg4def,cfg5"


# ----------------------------------------------------------------------------------------------
# workloads (scene catalogue shared with the tests: oracle/scenes.py is plain data)
# ----------------------------------------------------------------------------------------------
def load_workload(name):
    import scenes as sc
    base = "cfg4" if name == "cfg4def" else name
    w = dict(sc.WORKLOADS[base])
    w["name"] = name
    w["defect_normals"] = name == "cfg4def"   # IgnoreDefects=False: cfg4's second mode
    w["scene_spec"] = sc.resolve(w["scene"])
    if base == "cfg3":
        w["rays"] = w["rays"] // 8            # 100 M rays over 8 GPUs -> 12.5 M per GPU
    return w


def workload_config(w, n):
    """The `config` object: a function of the workload alone, so that both arms print the same one."""
    s = w["scene_spec"]
    cfg = {"workload": f"{w['name']}: {w['scene']} ({', '.join(o['kind'] for o in s['optics'])}), "
                       f"{n} rays per GPU, detector autoplace at {s['detector_distance']} mm",
           "rays_per_gpu": int(n), "elements": len(s["optics"]),
           "l2": "inputs larger than L2 (source columns + stored final bundle of one step exceed 126 MB)",
           "ignore_defects": not w["defect_normals"]}
    if w.get("sweep"):
        sw = w["sweep"]
        cfg.update({"variants_total": sw["n"], "sweep": f"{sw['axis']} of element {sw['element']} over "
                    f"[{sw['lo']}, {sw['hi']}] deg",
                    "l2": "source bundle (32 MB) re-read per variant from L2 by design; per-variant outputs are 34 doubles"})
    return cfg


def source_properties(w, n_total):
    sp = dict(w["scene_spec"]["source"])
    sp["NumberRays"] = int(n_total)
    return sp


def algorithmic_bytes(n_src, n_surv, want_inc=True, in_columns=7):
    """read the source columns (P, U, intensity = 7 doubles; 4 for a point source whose
    origin is shared by all rays); write P,U,path(,incidence)+alive per survivor."""
    return 8 * in_columns * n_src + (57 + (8 if want_inc else 0)) * n_surv


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: one streaming
    `nvidia-smi -lms 20` process started right before the region and stopped right after it."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            # wait until nvidia-smi is streaming (its start-up takes 0.1 s alone and over a second when
            # eight ranks start theirs together), so that the 20 ms samples fall INSIDE the timed region
            import select
            self.first = None
            if select.select([self.proc.stdout], [], [], 8.0)[0]:
                self.first = self.proc.stdout.readline()
        except Exception:
            self.proc = None

    def summary(self):
        rows = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                lines = out.splitlines()
                if not lines and self.first:  # a region shorter than one sampling period: the sample taken
                    lines = [self.first]      # right before it
                for line in lines:
                    parts = [p.strip() for p in line.split(",")]
                    if len(parts) >= 6:
                        rows.append(parts)
            except Exception:
                pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows)}


# ----------------------------------------------------------------------------------------------
# CPU legs.  kind "reference": the literal reference from oracle/_ref in P worker processes;
# kind "port": the numpy restatement oracle/art_oracle.py (labelled, never the headline).
# ----------------------------------------------------------------------------------------------
def reference_rate(w, n_total, workers, seconds, steps=1, warmup=0):
    """Interactions/s of the unmodified reference (ref_runner.ReferencePool) on bounded samples of the
    n_total-ray bundle: `steps` timed steps of about `seconds` each.  Returns (rate, sample, t_total, steps)."""
    import ref_runner
    pool = ref_runner.ReferencePool(w["scene"], n_total, workers)
    try:
        probe = 40 * workers
        pool.step(probe)                                   # page-in
        inter, wall = pool.step(probe)
        per_ray = wall / (probe / workers)                 # seconds per ray per worker
        sample = int(max(workers * 20, workers * seconds / per_ray))
        for _ in range(warmup):
            pool.step(sample)
        inter_total, t_total = 0, 0.0
        for _ in range(steps):
            inter, wall = pool.step(sample)
            inter_total += inter
            t_total += wall
    finally:
        pool.close()
    return inter_total / t_total, sample, t_total, steps


def ref_root():
    """Where the reference was imported from, relative to the repository when it is the travelling copy."""
    import ref_runner
    root = ref_runner.lr.REFERENCE_ROOT
    return os.path.relpath(root, ROOT) if root.startswith(ROOT) else root


def run_reference(args, w):
    """--impl reference: the unmodified reference on all host cores; imports only oracle/."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ref_runner
    if not ref_runner.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref missing: run python oracle/make_ref.py "
                          "where /root/reference exists"}), flush=True)
        return
    cores = os.cpu_count() or 1
    n = int(args.rays or w["rays"])
    n_total = n * args.gpus
    # each step is a bounded sample: ~args.ref_seconds of work per core, so that the driver's
    # --steps 20 --warmup 5 run ends within a few minutes
    value, sample, t_total, steps = reference_rate(w, n_total, cores, args.ref_seconds, steps=args.steps,
                                                   warmup=max(args.warmup - 1, 0) if args.ref_warm else 0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(w, n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{sample} rays (evenly spaced spiral indices) of the {n_total}-ray bundle per step: "
                                   f"ART v0.93 OEPlacement + RayTracingCalculation + Detector.autoplace + "
                                   f"GetResultSummary, unmodified ({ref_root()}), in {cores} processes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _oracle_elements(oes):
    """Oracle element dicts from the package's OpticalElement list (poses + optic parameters)."""
    els = []
    for oe in oes:
        o = oe.type
        base = getattr(o, "Mirror", o)
        kind = {"Plane Mirror": "plane", "SphericalCC Mirror": "spherical", "SphericalCX Mirror": "spherical",
                "Parabolic Mirror": "parabolic", "Toroidal Mirror": "toroidal", "Ellipsoidal Mirror": "ellipsoidal",
                "CylindricalCC Mirror": "cylindrical", "CylindricalCX Mirror": "cylindrical", "Mask": "mask"}[o.type]
        sk, sp = o.support._lower()
        sup = (["round", "roundhole", "rect", "recthole", "rectrecthole"][sk],) + tuple(sp[: [1, 4, 2, 5, 6][sk]])
        d = {"kind": kind, "support": sup}
        if kind in ("spherical", "cylindrical"):
            d["radius"] = base.radius
        elif kind == "parabolic":
            d.update(feff=base.feff, offaxisangle=base.offaxisangle, p=base.p)
        elif kind == "toroidal":
            d.update(majorradius=base.majorradius, minorradius=base.minorradius)
        elif kind == "ellipsoidal":
            d.update(a=base.a, b=base.b, offaxisangle=base._offaxisangle)
        if hasattr(o, "DeformationList"):
            d["defects"] = [{"kind": "zernike", "R": z.R, "max_order": z.max_order, "coefficients": dict(z.coefficients)}
                            for z in o.DeformationList]
        els.append({"optic": d, "position": oe.position, "normal": oe.normal, "majoraxis": oe.majoraxis})
    return els


def port_rate(w, oes, n_total, seconds):
    """Interactions/s of the numpy oracle (the PORT) on one core, ~`seconds` of work."""
    import art_oracle as orc
    sp = source_properties(w, n_total)
    els = _oracle_elements(oes)
    n_src = n_total - 1 if sp["Divergence"] == 0 else n_total
    dist = w["scene_spec"]["detector_distance"]

    def once(sample):
        idx = np.linspace(0, n_src - 1, sample).astype(np.int64)
        if sp["Divergence"] == 0:
            P, U, _ = orc.plane_wave_disk(np.zeros(3), orc.EX, sp["SourceSize"] / 2, sp["NumberRays"], idx)
        else:
            P, U, _ = orc.point_source(np.zeros(3), orc.EX, sp["Divergence"], sp["NumberRays"], idx)
        t0 = time.perf_counter()
        traced = orc.trace_chain(P, U, els, ignore_defects=not w["defect_normals"])
        last = traced[-1]
        if last["index"].size > 1:
            det = orc.detector_autoplace(last["P"], last["U"], dist)
            orc.result_summary(det, last["P"], last["U"], last["path"])
        return orc.count_interactions(P.shape[0], traced), time.perf_counter() - t0

    once(4000)
    inter, wall = once(4000)
    sample = int(min(4_000_000, max(2000, 4000 * seconds / wall)))
    inter, wall = once(sample)
    return inter / wall, sample, wall


# ----------------------------------------------------------------------------------------------
# scene spec -> objects of the package
# ----------------------------------------------------------------------------------------------
def build_optic(spec):
    import attosecondraytracing_b200.ModuleDefects as mdef
    import attosecondraytracing_b200.ModuleMask as mmask
    import attosecondraytracing_b200.ModuleMirror as mmirror
    import attosecondraytracing_b200.ModuleSupport as msupp
    kind, p = spec["support"][0], spec["support"][1:]
    sup = {"round": msupp.SupportRound, "roundhole": msupp.SupportRoundHole, "rect": msupp.SupportRectangle,
           "recthole": msupp.SupportRectangleHole, "rectrecthole": msupp.SupportRectangleRectHole}[kind](*p)
    k = spec["kind"]
    if k == "mask":
        return mmask.Mask(sup)
    if k == "plane":
        m = mmirror.MirrorPlane(sup)
    elif k == "spherical":
        m = mmirror.MirrorSpherical(spec["radius_signed"], sup)
    elif k == "cylindrical":
        m = mmirror.MirrorCylindrical(spec["radius_signed"], sup)
    elif k == "parabolic":
        m = mmirror.MirrorParabolic(spec["feff"], spec["offaxisangle_deg"], sup)
    elif k == "toroidal":
        m = mmirror.MirrorToroidal(spec["majorradius"], spec["minorradius"], sup)
    elif k == "ellipsoidal":
        kw = {a: spec[a] for a in ("SemiMajorAxis", "SemiMinorAxis", "OffAxisAngle", "f_object", "f_image") if a in spec}
        m = mmirror.MirrorEllipsoidal(sup, **kw)
    else:
        raise ValueError(k)
    if spec.get("defects"):
        dl = []
        for d in spec["defects"]:
            if d["kind"] != "zernike":
                raise ValueError("bench workloads carry Zernike defects only")
            dl.append(mdef.Zernike(sup, {(int(n), int(mm_)): c for n, mm_, c in d["coefficients"]}))
        m = mmirror.DeformedMirror(m, dl)
    return m


def build_chain_elements(w):
    """OpticalElement list of the workload's scene, aligned by the package's OEPlacement restatement."""
    import attosecondraytracing_b200.ModuleProcessing as mp
    s = w["scene_spec"]
    optics = [build_optic(o) for o in s["optics"]]
    oes = mp.place_optical_elements(optics, s["distances"], s["incidences"], s["plane_angles"])
    for op in s.get("post", []):
        getattr(oes[op["element"]], op["op"])(op["value"])
    return oes


# ----------------------------------------------------------------------------------------------
class Ctx:
    pass


def _events(n):
    import torch
    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


def _graph_or_eager(step, allow):
    """The step as a CUDA graph (its launches replayed without host work in between) when `allow`."""
    import torch
    if not allow:
        return step, False
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            captured = step()
        graph.replay()
        torch.cuda.synchronize()
        return (lambda: (graph.replay(), captured)[1]), True
    except Exception as exc:  # keep the eager step if capture is not possible
        print(f"[bench] CUDA graph capture failed ({exc}); running eagerly", file=sys.stderr)
        torch.cuda.synchronize()
        return step, False


def _roofline(ctx, kernel_name, k_ms, abytes, flops_kernel, traffic_key=None, note=None):
    hbm_peak = ctx.hbm_peak
    achieved = abytes / (k_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": None, "kernel": kernel_name, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": abytes,
            "peak_source": ctx.peak_src,
            "fp64": {"model_flops_per_launch": flops_kernel, "achieved_tflops": flops_kernel / (k_ms * 1e-3) / 1e12,
                     "peak_tflops_measured_dfma": None if ctx.fp64_peak is None else ctx.fp64_peak / 1e12,
                     "frac": None if ctx.fp64_peak is None else flops_kernel / (k_ms * 1e-3) / ctx.fp64_peak,
                     "model": "SURVEY.md 8(d) canonical FLOP count of the reference's algorithm (bench_flops.py); the "
                              "hardware pipe utilisation is in profiles/ (sm__pipe_fp64_cycles_active)"}}
    if note:
        roof["note"] = note
    t = ctx.traffic.get(traffic_key) if traffic_key else None
    if t:
        roof["traffic"] = t.get("traffic")
        roof["traffic_source"] = (f"profiles/{ctx.traffic_file}: dram__bytes_read.sum + dram__bytes_write.sum of one "
                                  f"ncu --set full capture of this kernel at commit {ctx.traffic.get('head')}")
    return roof


def measure_bundle(ctx, name, steps, warmup, rays=0, main=False, histograms=False):
    """One ray-sharded workload (cfg2 / cfg3 / cfg4 / cfg4def): weak scaling, rays dealt round-robin."""
    import torch
    import torch.distributed as dist
    from attosecondraytracing_b200 import _cabi, engine
    from attosecondraytracing_b200 import distributed as ad
    import attosecondraytracing_b200.ModuleSource as msrc
    from bench_flops import chain_flops  # canonical FLOP model of SURVEY.md 8(d)

    w = load_workload(name)
    n = int(rays or w["rays"])
    oes = build_chain_elements(w)
    rank, world, dev, lib, peer = ctx.rank, ctx.world, ctx.dev, ctx.lib, ctx.peer
    distance = w["scene_spec"]["detector_distance"]
    ign = not w["defect_normals"]  # the reference's default through get_output_rays is IgnoreDefects=True

    n_total = n * world                     # the bundle all ranks share (weak scaling)
    sp = source_properties(w, n_total)
    n_src_total = n_total - 1 if sp["Divergence"] == 0 else n_total
    # rays are dealt round-robin over the ranks (rank, rank + world, ...): every rank sees the whole
    # aperture, so masks that block a contiguous range of spiral indices do not unbalance the ranks
    first, count, stride = ad.shard_strided(n_src_total, rank, world)
    src = msrc.synthetic_source(sp, device=dev, first=first, count=count, stride=stride,
                                group=True if world > 1 else None)
    chain = engine.DeviceChain(oes, device=dev)
    # buffers of one step, allocated once and reused (no allocator traffic inside the timed region)
    out = chain.new_output(src, want_incidence=True)
    central_b = torch.empty((1, _cabi.CENTRAL_LEN), dtype=torch.float64, device=dev)
    det_b = torch.empty((1, _cabi.DETECTOR_DOUBLES), dtype=torch.float64, device=dev)
    mom_b = torch.empty((1, _cabi.MOMENTS_LEN), dtype=torch.float64, device=dev)
    inten = src.col("intensity")
    gather_b = torch.empty((world, 1, _cabi.MOMENTS_LEN), dtype=torch.float64, device=dev)
    hist_b = torch.empty((_cabi.hist_len(64, 64, 128),), dtype=torch.int64, device=dev) if histograms else None

    def step():
        if peer is not None:
            # the per-block rows of the trace / detector kernel are folded INSIDE the exchange kernels:
            # four launches per step (trace, exchange + autoplace, detector, exchange)
            chain.trace(src, ignore_defects=ign, history=False, want_incidence=True, out=out, central=central_b, fold=False)
            peer.all_reduce_central(central_b, distance, det_b, chain=chain)
            chain.moments(out, det_b, intensity=inten, fold=False)
            peer.all_reduce_moments(mom_b, chain=chain)
        else:
            chain.trace(src, ignore_defects=ign, history=False, want_incidence=True, out=out, central=central_b)
            ad.all_reduce_central(central_b)
            chain.autoplace(central_b, distance, det=det_b)
            chain.moments(out, det_b, intensity=inten, out=mom_b)
            ad.all_reduce_moments(mom_b, gather_buffer=gather_b)
        if hist_b is not None:  # SpotDiagram / DelayGraph bins over the merged extents, exact int64 SUM
            chain.histogram(out, det_b, mom_b, bins=(64, 64), delay_bins=128, intensity=inten, out=hist_b)
            ad.all_reduce_histogram(hist_b)
        return out, central_b, det_b, mom_b

    e, sv = chain.count_entering(src, ignore_defects=ign)
    entering = [int(x) for x in e[0].cpu()]
    n_surv = int(sv[0])
    interactions_rank = int(sum(entering))
    abytes = algorithmic_bytes(count, n_surv, in_columns=len(src._names))

    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    # Multi-GPU steps with NCCL calls in them stay eager (capturing the collectives changed nothing at N=2
    # and left the communicator unable to shut down cleanly); the peer-memory exchange has no NCCL call.
    nccl_in_step = world > 1 and (peer is None or histograms)
    run_step, graphed = _graph_or_eager(step, not ctx.args.no_graph and not nccl_in_step)

    sampler = ClockSampler(ctx.local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.art_launch_count()
    step()  # one eager step to count this step's launches
    launches_per_step = lib.art_launch_count() - launches0
    ctx.barrier()
    if peer is not None:
        peer.stats(reset=True)
    ev0.record()
    for _ in range(steps):
        final, central, det, mom = run_step()
    ev1.record()
    ctx.barrier()
    ms_total = ev0.elapsed_time(ev1)
    breakdown = None
    if peer is not None:
        if peer.status() != 0:
            raise RuntimeError("peer-memory exchange timed out (a rank did not arrive)")
        # what a timeline of the step would show: per exchange, how long each rank waited for the others' flags
        # (skew between the ranks + the NVLink round trip) and how long the exchange kernel ran
        st = peer.stats()
        mine = torch.tensor([st["wait_us"], st["kernel_us"]], dtype=torch.float64, device=dev)
        allr = torch.empty((world, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.cpu().numpy()
        breakdown = {"exchanges_per_step": st["exchanges"] / steps,
                     "wait_for_peers_us": {"min_rank": float(allr[:, 0].min()), "mean": float(allr[:, 0].mean()),
                                           "max_rank": float(allr[:, 0].max())},
                     "exchange_kernel_us": {"min_rank": float(allr[:, 1].min()), "mean": float(allr[:, 1].mean()),
                                            "max_rank": float(allr[:, 1].max())},
                     "note": "per exchange, %globaltimer inside peer_exchange_kernel; the rank that arrives last waits "
                             "only the NVLink round trip (min_rank), the others also the skew between the ranks"}

    # the dominant kernel alone, CUDA events on the launching stream.  want_central=False only skips the
    # separate fold launch: trace_kernel itself always reduces the central sums, so this IS the step's K1.
    kev = _events(max(3, min(steps, 50)))
    torch.cuda.synchronize()
    for e0, e1 in kev:
        e0.record()
        chain.trace(src, ignore_defects=ign, history=False, want_incidence=True, want_central=False, out=out)
        e1.record()
    torch.cuda.synchronize()
    k_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in kev]))
    tms = torch.tensor([ms_total, k_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(interactions_rank)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, k_ms = (float(x) for x in tms.cpu())
    interactions_all = float(tot.cpu()[0])
    # the timed region of a 0.4 ms step is shorter than nvidia-smi's sampling period: keep the SAME steps running
    # (untimed; the same count on every rank, from the all-reduced step time) for ~0.25 s so that the clocks line
    # describes this load
    for _ in range(max(0, min(2000, int(250.0 / max(ms_total / steps, 1e-3)) - steps))):
        run_step()
    torch.cuda.synchronize()
    clocks = sampler.summary()
    clocks["sampled"] = "timed region + ~0.25 s continuation of the same steps (20 ms period)"
    s = engine.summary_from_moments(mom.cpu().numpy()[0], central.cpu().numpy()[0])
    flops_kernel = chain_flops(oes, entering, n_surv, True) - n_surv * 60.0  # the detector is another kernel
    res = {
        "config": workload_config(w, n), "value": interactions_all * steps / (ms_total * 1e-3), "unit": UNIT,
        "scaling": "weak", "steps": steps, "ms_per_step": ms_total / steps,
        "roofline": _roofline(ctx, "trace_kernel<WANT_INC=1,WITH_DET=0>", k_ms, abytes, flops_kernel, traffic_key=name),
        "gpu_launches": int(launches_per_step) * steps, "cuda_graph": graphed, "clocks": clocks,
        "interactions_per_step": interactions_all, "survivors_rank0": n_surv,
        "result": {k: s[k] for k in ("SpotSizeSD", "DurationSD", "ETransmission") if k in s},
    }
    if histograms:
        res["histograms"] = "64x64 spot + 128 delay bins per step, int64 all-reduce"
    if breakdown is not None:
        res["multi_gpu_breakdown"] = breakdown

    if main:
        # ---- end to end, the reference's real host input: the source DESCRIPTION (SourceProperties) in,
        #      statistics out, through one C-ABI call; the bundle is generated on the device inside the call
        desc = msrc.source_descriptor(sp, first=first, count=count, stride=stride)
        import ctypes
        for _ in range(2):
            m_s, c_s, d_s = chain.run_source(desc, distance, ignore_defects=ign, peer=peer)
        ctx.barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, min(steps, 20))
        for _ in range(e2e_steps):
            m_s, c_s, d_s = chain.run_source(desc, distance, ignore_defects=ign, peer=peer)
        t_src = ctx.max_over_ranks(time.perf_counter() - t0)
        s_src = engine.summary_from_moments(m_s, c_s)
        same = all(abs(s_src[k] - s[k]) <= 1e-9 * max(1.0, abs(s[k])) for k in ("SpotSizeSD", "DurationSD", "ETransmission"))
        res["e2e"] = {"value": interactions_all * e2e_steps / t_src, "unit": UNIT,
                      "h2d_bytes_per_step": ctypes.sizeof(desc) + 24,
                      "d2h_bytes_per_step": 8 * (_cabi.MOMENTS_LEN + _cabi.CENTRAL_LEN + _cabi.DETECTOR_DOUBLES) + 16,
                      "steps": e2e_steps, "ms_per_step": 1e3 * t_src / e2e_steps,
                      "statistics_equal_device_step": bool(same),
                      "api": "art_run_source_host (ctypes): SourceProperties descriptor in; bundle generated + weighted "
                             "on the device (K0), trace, autoplace, moments; moments/central/detector out"
                             + ("" if peer is None else "; ranks combined over peer memory inside the call")}
        # ---- the same with HOST ray columns (a caller that already holds rays): PCIe-bound
        host = src.to("cpu").pin_memory()
        h2d = 8 * len(src._names) * count + (24 if src.origin is not None else 0)
        for _ in range(2):
            chain.run_host(host, distance, ignore_defects=ign, peer=peer)
        ctx.barrier()
        t0 = time.perf_counter()
        hc_steps = max(3, min(steps, 10))
        for _ in range(hc_steps):
            chain.run_host(host, distance, ignore_defects=ign, peer=peer)
        t_hc = ctx.max_over_ranks(time.perf_counter() - t0)
        res["e2e_host_columns"] = {
            "value": interactions_all * hc_steps / t_hc, "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": 8 * (_cabi.MOMENTS_LEN + _cabi.CENTRAL_LEN + _cabi.DETECTOR_DOUBLES),
            "steps": hc_steps, "ms_per_step": 1e3 * t_hc / hc_steps,
            "h2d_gbs_per_gpu": h2d * hc_steps / t_hc / 1e9,
            "numa": (f"each rank bound to the {ctx.numa_cpus} CPUs local to its GPU before pinning its staging"
                     if ctx.numa_cpus else "no CPU binding"),
            "api": "art_run_host" + ("" if peer is None else "_sharded") + " (ctypes): pinned host ray columns in "
                   "(8 PCIe chunks overlapped with the trace), moments/central/detector out"}
        del host
    ctx.oes_main = oes if main else getattr(ctx, "oes_main", None)
    chain.close()
    del src, out, chain
    torch.cuda.empty_cache()
    return res


def measure_sweep(ctx, name, steps, warmup, rays=0, variants=0):
    """cfg5: 1024 misaligned variants of the telescope x 10^6 rays.  STRONG scaling: the variant axis is
    sharded (1024 / N variants per rank, every rank holds the whole source bundle); the per-variant result
    rows are all-gathered, there is no per-ray traffic."""
    import copy
    import torch
    import torch.distributed as dist
    from attosecondraytracing_b200 import _cabi, engine
    import attosecondraytracing_b200.ModuleSource as msrc
    from bench_flops import chain_flops

    w = load_workload(name)
    sweep = w["sweep"]
    n = int(rays or w["rays"])
    oes = build_chain_elements(w)
    rank, world, dev, lib = ctx.rank, ctx.world, ctx.dev, ctx.lib
    distance = w["scene_spec"]["detector_distance"]
    nv_total = int(variants or sweep["n"])
    nv_rank = nv_total // world
    if nv_rank * world != nv_total:
        raise ValueError("the number of variants must be divisible by the number of GPUs")
    vals = np.linspace(sweep["lo"], sweep["hi"], nv_total)
    variants_ = []
    for x in vals[rank * nv_rank:(rank + 1) * nv_rank]:
        v = copy.deepcopy(oes)
        getattr(v[sweep["element"]], "rotate_%s_by" % sweep["axis"])(float(x))
        variants_.append(v)
    sp = source_properties(w, n)
    src = msrc.synthetic_source(sp, device=dev)
    count = src.n
    chain = engine.DeviceChain(variants_, device=dev)
    bufs = (torch.empty((nv_rank, _cabi.MOMENTS_LEN), dtype=torch.float64, device=dev),
            torch.empty((nv_rank, _cabi.CENTRAL_LEN), dtype=torch.float64, device=dev),
            torch.empty((nv_rank, _cabi.DETECTOR_DOUBLES), dtype=torch.float64, device=dev))
    gathered = torch.empty((world, nv_rank, _cabi.MOMENTS_LEN), dtype=torch.float64, device=dev)

    def step():
        mom, central, det = chain.sweep(src, distance, ignore_defects=True, out=bufs)
        if world > 1:
            dist.all_gather_into_tensor(gathered, mom)  # per-variant result rows, no per-ray traffic
        return mom, central, det

    entering = torch.zeros(chain.n_elements, dtype=torch.int64, device=dev)
    n_surv = 0
    for v0 in range(0, nv_rank, 16):
        e, sv = chain.count_entering(src, variant_first=v0, n_variants=min(16, nv_rank - v0))
        entering += e.sum(dim=0)
        n_surv += int(sv.sum())
    entering = [int(x) for x in entering.cpu()]
    interactions_rank = int(sum(entering))

    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    run_step, graphed = _graph_or_eager(step, not ctx.args.no_graph and world == 1)
    sampler = ClockSampler(ctx.local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.art_launch_count()
    step()
    launches_per_step = lib.art_launch_count() - launches0
    ctx.barrier()
    ev0.record()
    for _ in range(steps):
        mom, central, det = run_step()
    ev1.record()
    ctx.barrier()
    ms_total = ev0.elapsed_time(ev1)
    # second pass of art_sweep alone = fused trace + detector over all variants of this rank
    kev = _events(max(3, min(steps, 10)))
    torch.cuda.synchronize()
    for e0, e1 in kev:
        e0.record()
        chain.trace_detect(src, bufs[2], ignore_defects=True)
        e1.record()
    torch.cuda.synchronize()
    k_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in kev]))
    clocks = sampler.summary()
    tms = torch.tensor([ms_total, k_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(interactions_rank)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, k_ms = (float(x) for x in tms.cpu())
    interactions_all = float(tot.cpu()[0])

    # end to end: the host moves the bundle in (pinned) and the per-variant result rows out every step
    host = src.to("cpu").pin_memory()
    h2d = 8 * len(src._names) * count + (24 if src.origin is not None else 0)
    d2h = 8 * nv_rank * (_cabi.MOMENTS_LEN + _cabi.CENTRAL_LEN)
    dsrc = engine.RayBundle(count, device=dev, columns=host._names)
    dsrc.origin = src.origin
    ctx.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(steps, 3))
    for _ in range(e2e_steps):
        dsrc._storage.copy_(host._storage, non_blocking=True)
        m_, c_, d_ = chain.sweep(dsrc, distance, ignore_defects=True, out=bufs)
        rows = torch.cat([m_, c_], dim=1).cpu()  # noqa: F841
    torch.cuda.synchronize()
    t_e2e = ctx.max_over_ranks(time.perf_counter() - t0)

    # two traces per variant (central pass + detector pass); per-ray outputs are never stored
    flops = 2.0 * nv_rank * chain_flops(oes, [e_ / nv_rank for e_ in entering], n_surv / nv_rank, True)
    abytes = 8 * len(src._names) * count * nv_rank  # the source bundle re-read per variant (L2-resident)
    s = engine.summary_from_moments(mom.cpu().numpy()[0], central.cpu().numpy()[0])
    cfg = workload_config(w, n)
    cfg["variants_per_gpu"] = nv_rank
    res = {
        "config": cfg, "value": interactions_all * steps / (ms_total * 1e-3), "unit": UNIT, "scaling": "strong",
        "steps": steps, "ms_per_step": ms_total / steps,
        "roofline": _roofline(ctx, "trace_kernel<WANT_INC=0,WITH_DET=1> (2nd pass of art_sweep)", k_ms, abytes, flops / 2.0,
                              note="FP64-bound path: the source bundle stays in L2; the binding ceiling is the fp64 entry"),
        "e2e": {"value": interactions_all * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "api": "DeviceChain.sweep (art_sweep): pinned host bundle in, per-variant result rows out"},
        "gpu_launches": int(launches_per_step) * steps, "cuda_graph": graphed, "clocks": clocks,
        "interactions_per_step": interactions_all,
        "result_variant0_rank0": {k: s[k] for k in ("SpotSizeSD", "DurationSD", "ETransmission") if k in s},
    }
    chain.close()
    del src, dsrc, host, chain
    torch.cuda.empty_cache()
    return res


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    # CPU baseline first (rank 0, N = 1 only), BEFORE this process touches CUDA: the worker pool forks
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        w = load_workload(args.workload)
        n_cpu = int(args.rays or w["rays"])
        try:
            import ref_runner
            if not ref_runner.available():
                raise RuntimeError("oracle/_ref missing")
            cores = os.cpu_count() or 1
            rate, sample, wall, _ = reference_rate(w, n_cpu, cores, args.cpu_seconds)
            cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference",
                            "sample": f"{sample} rays (evenly spaced spiral indices) of the same {n_cpu}-ray bundle: "
                                      f"ART v0.93 OEPlacement + RayTracingCalculation + Detector.autoplace + "
                                      f"GetResultSummary, unmodified ({ref_root()}), in {cores} processes, {wall:.1f} s"}
        except Exception as exc:
            cpu_baseline = {"unavailable": f"literal reference could not run here: {exc}"}

    import ctypes
    import torch
    import torch.distributed as dist
    from attosecondraytracing_b200 import _cabi
    from attosecondraytracing_b200 import distributed as ad

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # host staging (the pinned ray columns of e2e_host_columns) should live on the GPU's own NUMA node
    numa_cpus = ad.bind_to_gpu_numa(local) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Ctx()
    ctx.numa_cpus = numa_cpus
    ctx.args, ctx.rank, ctx.world, ctx.local, ctx.dev = args, rank, world, local, dev
    ctx.lib = _cabi.lib()
    # multi-GPU: the two exchanges run inside one kernel each over peer memory (NVLink); NCCL when
    # symmetric memory is unavailable or --nccl asks for it
    ctx.peer = None if (world == 1 or args.nccl) else ad.PeerExchange.create(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        torch.cuda.synchronize()
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.cpu()[0])

    ctx.barrier, ctx.max_over_ranks = barrier, max_over_ranks
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    ctx.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ctx.peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    fp64 = ctypes.c_double()
    ctx.fp64_peak = fp64.value if ctx.lib.art_probe_fp64(fp64) == 0 else None
    ctx.traffic, ctx.traffic_file = {}, "r02_traffic.json"
    try:
        ctx.traffic = json.load(open(os.path.join(ROOT, "profiles", ctx.traffic_file)))
    except Exception:
        pass

    def measure(name, steps, warmup, main=False):
        if load_workload(name).get("sweep"):
            return measure_sweep(ctx, name, max(2, min(steps, 5)) if not main else steps, warmup,
                                 rays=args.rays if main else 0, variants=args.variants)
        return measure_bundle(ctx, name, steps, warmup, rays=args.rays if main else 0, main=main,
                              histograms=args.histograms and main)

    main_res = measure(args.workload, args.steps, args.warmup, main=True)
    subs = {}
    names = [] if args.sub in ("none", "") else [s for s in args.sub.split(",") if s and s != args.workload]
    for name in names:
        try:
            subs[name] = measure(name, max(3, min(args.steps, 10)), 3)
        except Exception as exc:  # a sub-workload must not take the headline down with it
            subs[name] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": main_res["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": main_res["config"], "roofline": main_res["roofline"], "e2e": main_res.get("e2e"),
            "gpu_launches": main_res["gpu_launches"], "cuda_graph": main_res["cuda_graph"],
            "clocks": main_res["clocks"], "interactions_per_step": main_res["interactions_per_step"],
        }
        for k in ("e2e_host_columns", "survivors_rank0", "result", "result_variant0_rank0", "histograms",
                  "multi_gpu_breakdown"):
            if k in main_res:
                line[k] = main_res[k]
        if world > 1:
            line["multi_gpu"] = ("central sums and moments exchanged inside one kernel each over peer memory (NVLink)"
                                 if ctx.peer is not None else "NCCL all-reduce + all-gather")
        if subs:
            line["workloads"] = subs
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
            if not args.no_port_baseline and getattr(ctx, "oes_main", None) is not None:
                try:
                    w = load_workload(args.workload)
                    n_cpu = int(args.rays or w["rays"])
                    rate, sample, wall = port_rate(w, ctx.oes_main, n_cpu, 4.0)
                    line["cpu_baseline_port"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                                                 "sample": f"{sample} rays of the same bundle, numpy restatement "
                                                           f"(oracle/art_oracle.py) on one core, {wall:.1f} s; NOT the "
                                                           "reference -- for orientation only"}
                except Exception as exc:
                    line["cpu_baseline_port"] = {"unavailable": str(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="cfg3", choices=["cfg2", "cfg3", "cfg4", "cfg4def", "cfg5"])
    ap.add_argument("--sub", default=DEFAULT_SUBS,
                    help="comma list of further workloads measured in the same run and attached as `workloads` (or none)")
    ap.add_argument("--rays", type=int, default=0, help="rays per GPU of the main workload (default: the workload's)")
    ap.add_argument("--variants", type=int, default=0, help="sweep workloads: total chain variants (default 1024)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work per core of the cpu_baseline sample")
    ap.add_argument("--ref-seconds", type=float, default=2.0, help="--impl reference: CPU work per core and step")
    ap.add_argument("--ref-warm", action="store_true", help="--impl reference: also run the warm-up steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-port-baseline", action="store_true")
    ap.add_argument("--nccl", action="store_true",
                    help="multi-GPU: exchange the central sums / moments with NCCL instead of the peer-memory kernel")
    ap.add_argument("--histograms", action="store_true",
                    help="also bin the detector response (64x64 spot + 128 delay bins) and all-reduce the int64 bins")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels eagerly instead of as a CUDA graph")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args, load_workload(args.workload))
        return
    run_b200(args)


if __name__ == "__main__":
    main()
