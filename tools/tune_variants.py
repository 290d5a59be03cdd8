#!/usr/bin/env python
"""Kernel tuning helper: times the trace kernel of every library build under build/variants/ (and the
in-tree one) on the BASELINE workloads.  Each build runs in its own process (ART_B200_LIB).

    python tools/tune_variants.py            # driver: one subprocess per library
    python tools/tune_variants.py --one      # worker
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def worker():
    import numpy as np
    import torch
    import bench
    from attosecondraytracing_b200 import engine
    import attosecondraytracing_b200.ModuleSource as msrc
    res = {}
    for wl, nrays in (("cfg2", 10_000_000), ("cfg3", 12_500_000), ("cfg5", 4_000_000), ("cfg4", 10_000_000)):
        w = bench.load_workload(wl)
        oes = bench.build_chain_elements(w)
        sp = bench.source_properties(w, nrays)
        src = msrc.synthetic_source(sp, device="cuda")
        chain = engine.DeviceChain(oes)
        for inc in (True, False):
            for _ in range(3):
                chain.trace(src, want_incidence=inc, want_central=True)
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                chain.trace(src, want_incidence=inc, want_central=True)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[f"{wl}_inc{int(inc)}_ms"] = round(float(np.median(ts)), 4)
        # detector kernel on the stored bundle
        outs, central = chain.trace(src, want_incidence=False)
        det = chain.autoplace(central, w["scene_spec"]["detector_distance"])
        ts = []
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            chain.moments(outs[0], det, intensity=src.col("intensity"))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[f"{wl}_det_ms"] = round(float(np.median(ts)), 4)
        # the statistics-only alternative: no stored bundle, two traces (central pass + fused detector pass)
        for _ in range(2):
            chain.sweep(src, w["scene_spec"]["detector_distance"])
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            chain.sweep(src, w["scene_spec"]["detector_distance"])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[f"{wl}_two_trace_stats_ms"] = round(float(np.median(ts)), 4)
        if wl == "cfg3":  # descriptor-driven end-to-end call (K0 + trace + detector), wall clock
            import time
            desc = msrc.source_descriptor(sp)
            for _ in range(3):
                chain.run_source(desc, w["scene_spec"]["detector_distance"])
            t0 = time.perf_counter()
            for _ in range(10):
                chain.run_source(desc, w["scene_spec"]["detector_distance"])
            res["cfg3_run_source_ms"] = round((time.perf_counter() - t0) * 100, 4)
        chain.close()
        del src, outs
        torch.cuda.empty_cache()
    # cfg5: the batched sweep (64 variants x 10^6 rays), both passes
    import copy
    w = bench.load_workload("cfg5")
    oes = bench.build_chain_elements(w)
    sw = w["sweep"]
    variants = []
    for x in np.linspace(sw["lo"], sw["hi"], 64):
        v = copy.deepcopy(oes)
        getattr(v[sw["element"]], "rotate_%s_by" % sw["axis"])(float(x))
        variants.append(v)
    src = msrc.synthetic_source(bench.source_properties(w, 1_000_000), device="cuda")
    chain = engine.DeviceChain(variants)
    for _ in range(2):
        chain.sweep(src, w["scene_spec"]["detector_distance"])
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        chain.sweep(src, w["scene_spec"]["detector_distance"])
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res["cfg5_sweep64_ms"] = round(float(np.median(ts)), 4)
    chain.close()
    print("RESULT " + json.dumps(res), flush=True)


def main():
    libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so")))
    libs.append(os.path.join(ROOT, "attosecondraytracing_b200", "libart_b200.so"))
    runs = [(lib, {}) for lib in libs]
    runs.append((libs[-1], {"ART_B200_DET_LEGACY": "1"}))  # A/B: the LDGSTS detector kernel instead of the bulk-copy one
    # process-to-process variation on a shared box is a few per cent: every build runs REPS times, interleaved with
    # the others, and the fastest time of each entry is kept
    reps = int(os.environ.get("TUNE_REPS", "2"))
    best = {}
    for _ in range(reps):
        for lib, extra in runs:
            key = os.path.basename(lib) + " " + " ".join(f"{k}={v}" for k, v in extra.items())
            env = dict(os.environ, ART_B200_LIB=lib, **extra)
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True,
                                     text=True, timeout=300)
                line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
                if not line:
                    best[key] = "FAILED " + out.stderr[-400:]
                    continue
                res = json.loads(line[0][7:])
                cur = best.get(key)
                best[key] = res if not isinstance(cur, dict) else {k: min(v, cur[k]) for k, v in res.items()}
            except subprocess.TimeoutExpired:
                best[key] = "TIMEOUT"
    for key, res in best.items():
        print(key, json.dumps(res) if isinstance(res, dict) else res, flush=True)


if __name__ == "__main__":
    worker() if "--one" in sys.argv else main()
