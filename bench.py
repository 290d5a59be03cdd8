#!/usr/bin/env python
"""Benchmark of the ray-bundle hot path (BASELINE.json metric: ray-element interactions/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5] [--impl reference]

One "step" = one pass of the hot path over one synthetic bundle: trace through every element of the
chain (fused kernel), all-reduce + autoplace of the detector, detector response + moments.  The
default workload is BASELINE config 2 (examples/CONFIG_toroidal2f-2f.py, 10M rays per GPU).
Scaling is weak: every rank traces its own 10M-ray slice of an N x 10M-ray bundle; the only
collectives are the all-reduces of the central sums and of the moments.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference
(oracle/, all host cores) on a bounded sample of the same workload.
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


# ----------------------------------------------------------------------------------------------
# workloads (scene catalogue shared with the tests: oracle/scenes.py is plain data)
# ----------------------------------------------------------------------------------------------
def load_workload(name):
    import scenes as sc
    w = dict(sc.WORKLOADS[name])
    w["name"] = name
    w["scene_spec"] = sc.resolve(w["scene"])
    return w


def build_chain_elements(w):
    """OpticalElement list of the workload's scene, aligned by the package's OEPlacement restatement."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import build_optic
    import attosecondraytracing_b200.ModuleProcessing as mp
    s = w["scene_spec"]
    optics = [build_optic(o) for o in s["optics"]]
    oes = mp.place_optical_elements(optics, s["distances"], s["incidences"], s["plane_angles"])
    for op in s.get("post", []):
        getattr(oes[op["element"]], op["op"])(op["value"])
    return oes


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
# CPU legs (oracle): the cpu_baseline object and --impl reference
# ----------------------------------------------------------------------------------------------
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


def _cpu_source(sp, idx):
    """Rows `idx` of the synthetic bundle for the CPU legs (oracle generators; the Gaussian weights are
    normalised with the bundle axis = +x and the edge ray, which is what the full-bundle normalisation
    amounts to for a Vogel spiral)."""
    import art_oracle as orc
    n = sp["NumberRays"]
    if sp["Divergence"] == 0:
        radius = sp["SourceSize"] / 2
        P, U, num = orc.plane_wave_disk(np.zeros(3), orc.EX, radius, n, idx)
        inten = np.exp(-2 * (orc.norm(P) / radius) ** 2)
    else:
        P, U, num = orc.point_source(np.zeros(3), orc.EX, sp["Divergence"], n, idx)
        ang = orc.angle_between(np.broadcast_to(orc.EX, U.shape), U)
        inten = np.exp(-2 * (np.tan(ang) / sp["Divergence"]) ** 2)
    return P, U, inten


def _cpu_worker(args):
    import art_oracle as orc
    sp, els, idx, dist = args
    P, U, inten = _cpu_source(sp, idx)
    t0 = time.perf_counter()  # timed: the path itself (trace + detector + statistics), not the source generation
    traced = orc.trace_chain(P, U, els, ignore_defects=True)
    last = traced[-1]
    if last["index"].size > 1:
        det = orc.detector_autoplace(last["P"], last["U"], dist)
        orc.result_summary(det, last["P"], last["U"], last["path"])
    return orc.count_interactions(P.shape[0], traced), time.perf_counter() - t0


def cpu_oracle_rate(w, oes, n_total, sample, workers):
    """Interactions/s of the numpy oracle on `sample` rays of the workload split over `workers`
    processes that run concurrently: interactions / slowest worker's compute time."""
    import multiprocessing as mpc
    sp = source_properties(w, n_total)
    els = _oracle_elements(oes)
    n_src = n_total - 1 if sp["Divergence"] == 0 else n_total
    idx = np.linspace(0, n_src - 1, sample).astype(np.int64)
    chunks = [c for c in np.array_split(idx, workers) if c.size]
    dist = w["scene_spec"]["detector_distance"]
    if workers == 1:
        res = [_cpu_worker((sp, els, chunks[0], dist))]
    else:
        with mpc.get_context("fork").Pool(workers) as pool:
            res = pool.map(_cpu_worker, [(sp, els, c, dist) for c in chunks])
    wall = max(r[1] for r in res)
    inter = sum(r[0] for r in res)
    return inter / wall, inter, wall


def auto_sample(w, oes, n_total, workers, seconds, lo=2000, hi=4_000_000):
    """Sample size (rays) whose CPU pass takes about `seconds`, from a small probe of the same workload."""
    probe = 4000 * workers
    cpu_oracle_rate(w, oes, n_total, probe, workers)  # imports, page-in
    rate, inter, wall = cpu_oracle_rate(w, oes, n_total, probe, workers)
    per_ray = wall / (probe / workers)  # seconds per ray per worker
    return int(min(hi, max(lo, workers * seconds / per_ray)))


def reference_literal_rate(w):
    """Speed of the UNMODIFIED reference on this workload's seeded subset, as measured in the build
    container while generating the golden fixture (tests/golden/<cfg>_sub*.npz metadata; the reference
    itself cannot travel to the GPU box)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from golden_util import Golden, golden_names
        names = [n for n in golden_names() if n.startswith(w["name"] + "_sub")]
        g = Golden(names[0])
        return {"value": g.spec["interactions"] / g.spec["reference_trace_seconds"], "unit": UNIT, "cores": 1,
                "where": "build container, ART v0.93 RayTracingCalculation on fixture " + names[0]}
    except Exception:
        return None


def run_reference(args, w, oes):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_total = w["rays"]
    # each step is a bounded sample of the workload: ~0.3 s of CPU work per step so that the default
    # --steps 200 --warmup 10 run ends within a few minutes
    sample = args.cpu_sample or auto_sample(w, oes, n_total, cores, 0.3)
    for _ in range(max(args.warmup - 1, 0)):
        cpu_oracle_rate(w, oes, n_total, sample, cores)
    rates, inter_total, t_total = [], 0, 0.0
    for _ in range(args.steps):
        r, inter, wall = cpu_oracle_rate(w, oes, n_total, sample, cores)
        rates.append(r)
        inter_total += inter
        t_total += wall
    value = inter_total / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(w, args, n_total),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} rays of the {n_total}-ray bundle per step, numpy oracle "
                                   f"(oracle/art_oracle.py) in {cores} processes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(w, args, n_total):
    s = w["scene_spec"]
    return {"workload": f"{w['name']}: {w['scene']} ({', '.join(o['kind'] for o in s['optics'])}), "
                        f"{n_total} rays per GPU, detector autoplace at {s['detector_distance']} mm",
            "rays_per_gpu": int(n_total), "elements": len(s["optics"]),
            "l2": "inputs larger than L2 (>= 320 MB of source columns + 650 MB of outputs per step)",
            "ignore_defects": not getattr(args, "defect_normals", False)}


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--rays", type=int, default=0, help="rays per GPU (default: the workload's)")
    ap.add_argument("--variants", type=int, default=0, help="sweep workloads: chain variants per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl", action="store_true",
                    help="multi-GPU: exchange the central sums / moments with NCCL instead of the peer-memory kernel")
    ap.add_argument("--defect-normals", action="store_true",
                    help="IgnoreDefects=False: surface defects also tilt the normals (SURVEY.md 8d, cfg4's second mode)")
    ap.add_argument("--histograms", action="store_true",
                    help="also bin the detector response (64x64 spot + 128 delay bins) and all-reduce the int64 bins")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels eagerly instead of as a CUDA graph")
    args = ap.parse_args()

    w = load_workload(args.workload)
    if args.rays:
        w["rays"] = args.rays
    if args.workload == "cfg3" and not args.rays:
        w["rays"] = w["rays"] // 8  # 100M rays over 8 GPUs -> 12.5M per GPU
    oes = build_chain_elements(w)
    if args.impl == "reference":
        run_reference(args, w, oes)
        return

    run_b200(args, w, oes)


def run_b200(args, w, oes):
    import copy
    import ctypes
    import torch
    import torch.distributed as dist
    from attosecondraytracing_b200 import _cabi, engine
    from attosecondraytracing_b200 import distributed as ad
    import attosecondraytracing_b200.ModuleSource as msrc
    from bench_flops import chain_flops  # canonical FLOP model of SURVEY.md 8(d)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()
    distance = w["scene_spec"]["detector_distance"]
    sweep = w.get("sweep")
    n = int(w["rays"])
    ign = not args.defect_normals  # the reference's default through get_output_rays is IgnoreDefects=True

    if sweep:
        # cfg5: every rank holds the full n-ray bundle and its share of the (weak-scaled) variant axis
        nv_rank = int(args.variants or sweep["n"])
        nv_total = nv_rank * world
        vals = np.linspace(sweep["lo"], sweep["hi"], nv_total)
        variants = []
        for x in vals[rank * nv_rank:(rank + 1) * nv_rank]:
            v = copy.deepcopy(oes)
            getattr(v[sweep["element"]], "rotate_%s_by" % sweep["axis"])(float(x))
            variants.append(v)
        sp = source_properties(w, n)
        src = msrc.synthetic_source(sp, device=dev)
        count = src.n
        chain = engine.DeviceChain(variants, device=dev)
        bufs = (torch.empty((nv_rank, _cabi.MOMENTS_LEN), dtype=torch.float64, device=dev),
                torch.empty((nv_rank, _cabi.CENTRAL_LEN), dtype=torch.float64, device=dev),
                torch.empty((nv_rank, _cabi.DETECTOR_DOUBLES), dtype=torch.float64, device=dev))
        gathered = torch.empty((world, nv_rank, _cabi.MOMENTS_LEN), dtype=torch.float64, device=dev)

        def step():
            mom, central, det = chain.sweep(src, distance, ignore_defects=True, out=bufs)
            if world > 1:
                dist.all_gather_into_tensor(gathered, mom)  # per-variant result rows, no per-ray traffic
            return None, central, det, mom

        entering = torch.zeros(chain.n_elements, dtype=torch.int64, device=dev)
        n_surv = 0
        for v0 in range(0, nv_rank, 16):
            e, sv = chain.count_entering(src, variant_first=v0, n_variants=min(16, nv_rank - v0))
            entering += e.sum(dim=0)
            n_surv += int(sv.sum())
        entering = [int(x) for x in entering.cpu()]
        kernel_name = "trace_kernel<WANT_INC=0,WITH_DET=1> (2nd pass of art_sweep)"
        abytes = None
    else:
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
        hist_b = torch.empty((_cabi.hist_len(64, 64, 128),), dtype=torch.int64, device=dev) if args.histograms else None
        # multi-GPU: the two exchanges run inside one kernel each over peer memory (NVLink); NCCL when
        # symmetric memory is unavailable or --nccl asks for it
        peer = None if (world == 1 or args.nccl) else ad.PeerExchange.create(dev)

        def step():
            chain.trace(src, ignore_defects=ign, history=False, want_incidence=True, out=out, central=central_b)
            if peer is not None:
                peer.all_reduce_central(central_b, distance, det_b)   # sum over ranks + autoplace
            else:
                ad.all_reduce_central(central_b)
                chain.autoplace(central_b, distance, det=det_b)
            chain.moments(out, det_b, intensity=inten, out=mom_b)
            if peer is not None:
                peer.all_reduce_moments(mom_b)
            else:
                ad.all_reduce_moments(mom_b, gather_buffer=gather_b)
            if hist_b is not None:  # SpotDiagram / DelayGraph bins over the merged extents, exact int64 SUM
                chain.histogram(out, det_b, mom_b, bins=(64, 64), delay_bins=128, intensity=inten, out=hist_b)
                ad.all_reduce_histogram(hist_b)
            return out, central_b, det_b, mom_b

        e, sv = chain.count_entering(src, ignore_defects=ign)
        entering = [int(x) for x in e[0].cpu()]
        n_surv = int(sv[0])
        kernel_name = "trace_kernel<WANT_INC=1,WITH_DET=0>"
        abytes = algorithmic_bytes(count, n_surv, in_columns=len(src._names))
    interactions_rank = int(sum(entering))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        final, central, det, mom = step()
    torch.cuda.synchronize()

    # the step as a CUDA graph: its kernel launches replayed without host work in between.
    # Multi-GPU steps stay eager: capturing the two NCCL collectives in the graph was measured at N=2
    # (0.442 vs 0.446 ms per step -- the collectives' latency, not launch overhead, is what the step waits
    # for) and the captured communicator did not shut down cleanly, so the option was removed.
    run_step = step
    graphed = False
    peer_graph = world > 1 and not sweep and peer is not None and not args.histograms  # no NCCL call in the step
    if not args.no_graph and (world == 1 or peer_graph):
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
            run_step = lambda: (graph.replay(), captured)[1]  # noqa: E731
            graphed = True
        except Exception as exc:  # keep the eager step if capture is not possible
            print(f"[bench] CUDA graph capture failed ({exc}); running eagerly", file=sys.stderr)
            torch.cuda.synchronize()

    # ---- timed region: K steps, device-timed, max over ranks --------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.art_launch_count()
    step()  # one eager step to count this step's launches
    launches_per_step = lib.art_launch_count() - launches0
    barrier()
    ev0.record()
    for _ in range(args.steps):
        final, central, det, mom = run_step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1 and not sweep and peer is not None and peer.status() != 0:
        raise RuntimeError("peer-memory exchange timed out (a rank did not arrive)")

    # ---- the dominant kernel alone, CUDA events on the launching stream ----------------------------
    ksteps = max(3, min(args.steps, 50))
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ksteps)]
    torch.cuda.synchronize()
    if sweep:
        # second pass of art_sweep = fused trace + detector over all variants of this rank
        for e0, e1 in kev:
            e0.record()
            chain.trace_detect(src, bufs[2], ignore_defects=True)
            e1.record()
    else:
        for e0, e1 in kev:
            e0.record()
            chain.trace(src, ignore_defects=ign, history=False, want_incidence=True, want_central=False, out=out)
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
    value = interactions_all * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (H2D + D2H inside the timed region) ---------
    e2e = None
    if not sweep:
        host = src.to("cpu").pin_memory()
        h2d = 8 * len(src._names) * count + (24 if src.origin is not None else 0)
        d2h = 8 * (_cabi.MOMENTS_LEN + _cabi.CENTRAL_LEN) + 8 * _cabi.DETECTOR_DOUBLES
        for _ in range(2):
            chain.run_host(host, distance, ignore_defects=ign, peer=peer)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(e2e_steps):
            mom_h, cen_h, det_h = chain.run_host(host, distance, ignore_defects=ign, peer=peer)
        torch.cuda.synchronize()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = {"value": interactions_all * e2e_steps / float(t_e2e.cpu()[0]), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "api": "art_run_host (ctypes, pinned host columns in, moments/central/detector out)" if peer is None else
                      "art_run_host_sharded (ctypes, pinned host columns of this rank's shard in, whole-bundle "
                      "moments/central/detector out; ranks combined over peer memory inside the call)"}
    else:
        # the sweep's per-step host traffic is the pose table in (built once) and the result rows out
        host = src.to("cpu").pin_memory()
        h2d = 8 * len(src._names) * count + (24 if src.origin is not None else 0)
        d2h = 8 * nv_rank * (_cabi.MOMENTS_LEN + _cabi.CENTRAL_LEN)
        dsrc = engine.RayBundle(count, device=dev, columns=host._names)
        dsrc.origin = src.origin
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(e2e_steps):
            dsrc._storage.copy_(host._storage, non_blocking=True)
            m_, c_, d_ = chain.sweep(dsrc, distance, ignore_defects=True, out=bufs)
            rows = torch.cat([m_, c_], dim=1).cpu()
        torch.cuda.synchronize()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = {"value": interactions_all * e2e_steps / float(t_e2e.cpu()[0]), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "api": "DeviceChain.sweep (art_sweep): pinned host bundle in, per-variant result rows out"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        fp64 = ctypes.c_double()
        fl = fp64.value if lib.art_probe_fp64(fp64) == 0 else None
        if sweep:
            # two traces per variant (central pass + detector pass); per-ray outputs are never stored
            flops = 2.0 * nv_rank * chain_flops(oes, [e_ / nv_rank for e_ in entering], n_surv / nv_rank, True)
            flops_kernel = flops / 2.0
            abytes = 8 * len(src._names) * count * nv_rank  # the source bundle re-read per variant (L2-resident)
            roof = {"bound": "hbm", "achieved": abytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": abytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                    "note": "FP64-bound path: the source bundle stays in L2; the binding ceiling is the fp64 entry"}
        else:
            flops_kernel = chain_flops(oes, entering, n_surv, True) - n_surv * 60.0  # detector is another kernel
            achieved = abytes / (k_ms * 1e-3) / 1e9
            traffic = None
            try:  # DRAM bytes of this kernel from the committed ncu --set full capture of the same command
                tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
                if w["name"] in tj and not args.rays:
                    traffic = tj[w["name"]]["traffic"]
            except Exception:
                pass
            roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic}
        roof.update({"kernel": kernel_name, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": abytes,
                     "peak_source": peak_src,
                     "fp64": {"model_flops_per_launch": flops_kernel,
                              "achieved_tflops": flops_kernel / (k_ms * 1e-3) / 1e12,
                              "peak_tflops_measured_dfma": None if fl is None else fl / 1e12,
                              "frac": None if fl is None else flops_kernel / (k_ms * 1e-3) / fl}})
        s = engine.summary_from_moments(mom.cpu().numpy()[0], central.cpu().numpy()[0])
        cfg = workload_config(w, args, n)
        if args.histograms and not sweep:
            cfg["histograms"] = "64x64 spot + 128 delay bins per step, int64 all-reduce"
        if world > 1 and not sweep:
            cfg["exchange"] = ("central sums and moments exchanged inside one kernel each over peer memory (NVLink)"
                               if peer is not None else "NCCL all-reduce + all-gather")
        if sweep:
            cfg.update({"variants_per_gpu": nv_rank, "sweep": f"{sweep['axis']} of element {sweep['element']} over "
                        f"[{sweep['lo']}, {sweep['hi']}] deg", "l2": "source bundle (56 MB) re-read per variant from L2 "
                        "by design; per-variant outputs are 34 doubles"})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "roofline": roof, "e2e": e2e, "gpu_launches": int(launches_per_step) * args.steps, "cuda_graph": graphed,
            "clocks": clocks, "interactions_per_step": interactions_all, "survivors_rank0": int(n_surv),
            "result": {k: s[k] for k in ("SpotSizeSD", "DurationSD", "ETransmission") if k in s},
        }
        if not args.no_cpu_baseline and world == 1:
            n_cpu = n if sweep else n * world
            sample = args.cpu_sample or auto_sample(w, oes, n_cpu, 1, 15.0)  # ~15 s on one core
            rate, inter, wall = cpu_oracle_rate(w, oes, n_cpu, sample, 1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{sample} rays of the same {n_cpu}-ray bundle, numpy oracle "
                                              f"(oracle/art_oracle.py), {wall:.1f} s",
                                    "reference_literal": reference_literal_rate(w)}
        print(json.dumps(line), flush=True)
    chain.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def C_double():
    import ctypes
    return ctypes.c_double()


if __name__ == "__main__":
    main()
