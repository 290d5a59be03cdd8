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
import threading
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


def algorithmic_bytes(n_src, n_surv, want_inc=True, with_intensity=True):
    """read P,U (+intensity) per source ray; write P,U,path(,incidence)+alive per survivor."""
    return (48 + (8 if with_intensity else 0)) * n_src + (57 + (8 if want_inc else 0)) * n_surv


# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [p.strip() for p in out.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


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


def _cpu_worker(args):
    import art_oracle as orc
    sp, els, idx, dist = args
    P, U, num, inten = orc.source_for(sp, k=idx)
    t0 = time.perf_counter()
    traced = orc.trace_chain(P, U, els, ignore_defects=True)
    last = traced[-1]
    if last["index"].size > 1:
        det = orc.detector_autoplace(last["P"], last["U"], dist)
        orc.result_summary(det, last["P"], last["U"], last["path"])
    return orc.count_interactions(P.shape[0], traced), time.perf_counter() - t0


def cpu_oracle_rate(w, oes, n_total, sample, workers):
    """Interactions/s of the numpy oracle on `sample` rays of the workload split over `workers` processes."""
    import multiprocessing as mpc
    sp = source_properties(w, n_total)
    els = _oracle_elements(oes)
    n_src = n_total - 1 if sp["Divergence"] == 0 else n_total
    idx = np.linspace(0, n_src - 1, sample).astype(np.int64)
    chunks = [c for c in np.array_split(idx, workers) if c.size]
    dist = w["scene_spec"]["detector_distance"]
    t0 = time.perf_counter()
    if workers == 1:
        res = [_cpu_worker((sp, els, chunks[0], dist))]
    else:
        with mpc.get_context("fork").Pool(workers) as pool:
            res = pool.map(_cpu_worker, [(sp, els, c, dist) for c in chunks])
    wall = time.perf_counter() - t0
    inter = sum(r[0] for r in res)
    return inter / wall, inter, wall


def run_reference(args, w, oes):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_total = w["rays"]
    sample = args.cpu_sample or 400_000
    cpu_oracle_rate(w, oes, n_total, min(sample, 20000), cores)  # warm-up (imports, page-in)
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
            "l2": "inputs larger than L2 (>= 480 MB of ray columns per step)", "ignore_defects": True}


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--rays", type=int, default=0, help="rays per GPU (default: the workload's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    import torch
    import torch.distributed as dist
    from attosecondraytracing_b200 import _cabi, engine
    import attosecondraytracing_b200.ModuleSource as msrc

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()

    n = int(w["rays"])                      # rays per GPU
    n_total = n * world                     # the bundle all ranks share
    sp = source_properties(w, n_total)
    n_src_total = n_total - 1 if sp["Divergence"] == 0 else n_total
    first = rank * n
    count = min(n, n_src_total - first)
    src = msrc.synthetic_source(sp, device=dev, first=first, count=count, group=True if world > 1 else None)
    chain = engine.DeviceChain(oes, device=dev)
    distance = w["scene_spec"]["detector_distance"]
    K = chain.n_elements

    # buffers of one step, allocated once and reused (no allocator traffic inside the timed region)
    out = chain.new_output(src, want_incidence=True)
    central = torch.empty((1, _cabi.CENTRAL_LEN), dtype=torch.float64, device=dev)
    det = torch.empty((1, _cabi.DETECTOR_DOUBLES), dtype=torch.float64, device=dev)
    mom_local = torch.empty((1, _cabi.MOMENTS_LEN), dtype=torch.float64, device=dev)
    inten = src.col("intensity")

    def step():
        chain.trace(src, ignore_defects=True, history=False, want_incidence=True, out=out, central=central)
        if world > 1:
            dist.all_reduce(central, op=dist.ReduceOp.SUM)
        chain.autoplace(central, distance, det=det)
        mom, _, _, _ = chain.moments(out, det, intensity=inten, out=mom_local)
        if world > 1:
            sums = mom[:, :14].contiguous()
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            mx = torch.cat([-mom[:, [14, 16, 18]], mom[:, [15, 17, 19, 20]]], dim=1).contiguous()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            mom = torch.cat([sums, torch.stack([-mx[:, 0], mx[:, 3], -mx[:, 1], mx[:, 4], -mx[:, 2], mx[:, 5], mx[:, 6]],
                                               dim=1), mom[:, 21:]], dim=1)
        return out, central, det, mom

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (also gives the per-element interaction counts)
    for _ in range(max(args.warmup, 3)):
        final, central, det, mom = step()
    torch.cuda.synchronize()
    hist, _ = chain.trace(src, history=True, want_central=False)
    entering = [count] + [len(h) for h in hist[:-1]]
    interactions_rank = int(sum(entering))
    n_surv = len(hist[-1])
    del hist
    torch.cuda.empty_cache()

    # the step as a CUDA graph: five kernel launches (trace, fold, autoplace, detector, fold) replayed
    # without host work in between.  Multi-GPU steps keep their NCCL all-reduces eager.
    run_step = step
    graphed = False
    if world == 1 and not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            graph.replay()
            torch.cuda.synchronize()
            run_step = lambda: (graph.replay(), (out, central, det, mom_local))[1]  # noqa: E731
            graphed = True
        except Exception as exc:  # keep the eager step if capture is not possible
            print(f"[bench] CUDA graph capture failed ({exc}); running eagerly", file=sys.stderr)
            torch.cuda.synchronize()

    # ---- timed region: K steps, device-timed, max over ranks --------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.art_launch_count()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        final, central, det, mom = run_step()
    ev1.record()
    barrier()
    launches = lib.art_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.summary()

    # ---- the dominant kernel alone (trace_kernel), CUDA events on the launching stream -------------
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for e0, e1 in kev:
        e0.record()
        chain.trace(src, ignore_defects=True, history=False, want_incidence=True, want_central=False, out=out)
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
    value = interactions_all * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (H2D + D2H inside the timed region) ---------
    e2e = None
    host = src.to("cpu").pin_memory()
    h2d = (7 * 8) * count
    d2h = 8 * (_cabi.MOMENTS_LEN + _cabi.CENTRAL_LEN) + 8 * _cabi.DETECTOR_DOUBLES
    for _ in range(2):
        chain.run_host(host, distance)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        mom_h, cen_h, det_h = chain.run_host(host, distance)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e = {"value": interactions_all * e2e_steps / float(t_e2e.cpu()[0]), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": e2e_steps,
           "api": "art_run_host (ctypes, pinned host columns in, moments/central/detector out)"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        abytes = algorithmic_bytes(count, n_surv)
        achieved = abytes / (k_ms * 1e-3) / 1e9
        fp64 = C_double()
        fl = None
        if lib.art_probe_fp64(fp64) == 0:
            fl = fp64.value
        from bench_flops import chain_flops  # canonical FLOP model of SURVEY.md 8(d)
        flops = chain_flops(oes, entering, n_surv, ignore_defects=True)
        s = engine.summary_from_moments(mom.cpu().numpy()[0], central.cpu().numpy()[0])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(w, args, n),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": None, "kernel": "trace_kernel<WANT_INC=1,WITH_DET=0>",
                         "kernel_ms": k_ms, "algorithmic_bytes_per_launch": abytes, "peak_source": peak_src,
                         "fp64": {"model_flops_per_launch": flops, "achieved_tflops": flops / (k_ms * 1e-3) / 1e12,
                                  "peak_tflops_measured_dfma": None if fl is None else fl / 1e12,
                                  "frac": None if fl is None else flops / (k_ms * 1e-3) / fl}},
            "e2e": e2e, "gpu_launches": int(launches) if not graphed else 5 * args.steps, "cuda_graph": graphed,
            "clocks": clocks,
            "interactions_per_step": interactions_all, "survivors_rank0": int(n_surv),
            "result": {k: s[k] for k in ("SpotSizeSD", "DurationSD", "ETransmission") if k in s},
        }
        if not args.no_cpu_baseline and world == 1:
            sample = args.cpu_sample or 200_000
            cpu_oracle_rate(w, oes, n_total, 10000, 1)
            rate, inter, wall = cpu_oracle_rate(w, oes, n_total, sample, 1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{sample} rays of the same {n_total}-ray bundle, numpy oracle "
                                              f"(oracle/art_oracle.py), {wall:.1f} s"}
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
