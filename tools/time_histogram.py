#!/usr/bin/env python
"""Times art_detector_histogram (64x64 + 128 bins) on the cfg2 / cfg3 bundles for the library in ART_B200_LIB."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import bench
from attosecondraytracing_b200 import engine
import attosecondraytracing_b200.ModuleSource as msrc

for wl, nrays in (("cfg2", 10_000_000), ("cfg3", 12_500_000)):
    w = bench.load_workload(wl)
    src = msrc.synthetic_source(bench.source_properties(w, nrays), device="cuda")
    chain = engine.DeviceChain(bench.build_chain_elements(w))
    outs, central = chain.trace(src)
    det = chain.autoplace(central, w["scene_spec"]["detector_distance"])
    mom, _, _, _ = chain.moments(outs[0], det, intensity=src.col("intensity"))
    for bins, nt in (((64, 64), 128), ((512, 512), 1024)):
        ts = []
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h = chain.histogram(outs[0], det, mom, bins=bins, delay_bins=nt, intensity=src.col("intensity"))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(os.path.basename(os.environ.get("ART_B200_LIB", "in-tree")), wl, bins, nt, "%.4f ms" % float(np.median(ts)),
              int(h[:bins[0] * bins[1]].sum()))
    chain.close()
