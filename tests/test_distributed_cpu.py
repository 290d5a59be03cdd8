"""world_size-2 gloo test of the multi-GPU plumbing (host logic only, CPU tensors): index-range
sharding plus the central / moments all-reduces reproduce the single-process statistics.
The per-shard rows are computed with the oracle -- what the GPU kernels produce per rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import art_oracle as orc
from golden_util import Golden


HIST_BINS, HIST_NT = (12, 10), 16


def oracle_central(P, U, path, w_out, w_in_sum):
    return np.array([*U.sum(axis=0), *P.sum(axis=0), path.sum(), P.shape[0], w_out.sum(), w_in_sum])


def oracle_moments(det, l0, P, U, path, w):
    """The moments row (include/art_b200.h ART_M_*) from oracle arithmetic."""
    xy = orc.detector_points2d(det, P, U)
    L = orc.detector_optical_paths(det, P, U, path)
    d = L - l0
    x, y = xy[:, 0], xy[:, 1]
    cv = -np.asarray(det["normal"])
    tan2 = np.sum((U - cv) ** 2, axis=1) / np.sum((U + cv) ** 2, axis=1)
    m = np.zeros(24)
    m[:14] = [P.shape[0], x.sum(), y.sum(), (x * x).sum(), (y * y).sum(), d.sum(), (d * d).sum(), w.sum(),
              (w * x).sum(), (w * y).sum(), (w * x * x).sum(), (w * y * y).sum(), (w * d).sum(), (w * d * d).sum()]
    if P.shape[0]:  # a shard without survivors keeps the reduction identities, as the kernel does
        m[14:21] = [x.min(), x.max(), y.min(), y.max(), d.min(), d.max(), tan2.max()]
    else:
        m[14:21] = [np.inf, -np.inf, np.inf, -np.inf, np.inf, -np.inf, -np.inf]
    return m


def oracle_hist_vector(det, l0, m, P, U, path, w, bins, nt):
    """One shard's int64 histogram in the layout of art_detector_histogram (include/art_b200.h), binned
    with the kernel's rule floor((v - lo) * nbins / (hi - lo)) against the MERGED extents of row m."""
    nx, ny = bins
    xy = orc.detector_points2d(det, P, U)
    d = orc.detector_optical_paths(det, P, U, path) - l0

    def bin_of(v, lo, hi, nb):
        k = np.floor((v - lo) * (nb / (hi - lo))).astype(np.int64) if hi > lo else np.zeros(len(v), np.int64)
        return np.clip(k, 0, nb - 1)

    one = 2.0 ** 26
    ix, iy = bin_of(xy[:, 0], m[14], m[15], nx), bin_of(xy[:, 1], m[16], m[17], ny)
    it = bin_of(d, m[18], m[19], nt)
    wq = np.rint(np.clip(w, 0.0, 1.0) * one).astype(np.int64)
    dq = np.rint(np.clip((d - m[18]) / (m[19] - m[18]), 0.0, 1.0) * one).astype(np.int64)
    h = np.zeros(3 * nx * ny + 2 * nt, dtype=np.int64)
    np.add.at(h, ix * ny + iy, 1)
    np.add.at(h, nx * ny + ix * ny + iy, wq)
    np.add.at(h, 2 * nx * ny + ix * ny + iy, dq)
    np.add.at(h, 3 * nx * ny + it, 1)
    np.add.at(h, 3 * nx * ny + nt + it, wq)
    return h


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from attosecondraytracing_b200 import distributed as ad
    from attosecondraytracing_b200.engine import summary_from_moments
    g = Golden(name)
    n = g["src_P"].shape[0]
    first, count = ad.shard_range(n, rank, world)
    sl = slice(first, first + count)
    traced = orc.trace_chain(g["src_P"][sl], g["src_U"][sl], g.oracle_elements(), ignore_defects=g.ignore_defects)
    last = traced[-1]
    w_src = g["src_I"][sl]
    w = w_src[last["index"]]
    central = torch.from_numpy(oracle_central(last["P"], last["U"], last["path"], w, w_src.sum())).reshape(1, -1)
    ad.all_reduce_central(central)
    c = central.numpy()[0]
    # every rank derives the same detector from the reduced sums (Detector.autoplace)
    N = c[7]
    cv = orc.normalize(c[0:3] / N)
    cp = c[3:6] / N
    dd = g.spec["detector_distance"]
    det = {"normal": -cv, "centre": cp + cv * dd, "refpoint": cp}
    l0 = c[6] / N + dd
    mom = torch.from_numpy(oracle_moments(det, l0, last["P"], last["U"], last["path"], w)).reshape(1, -1)
    ad.all_reduce_moments(mom)
    s = summary_from_moments(mom.numpy()[0], c)
    # binned detector response: every rank bins its shard against the merged extents, int64 SUM all-reduce
    hist = torch.from_numpy(oracle_hist_vector(det, l0, mom.numpy()[0], last["P"], last["U"], last["path"], w,
                                               HIST_BINS, HIST_NT))
    ad.all_reduce_histogram(hist)
    if rank == 0:
        q.put((s, det["centre"], det["normal"], hist.numpy(), mom.numpy()[0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["cfg3_2tor", "cfg1_par"])
def test_two_ranks_reproduce_the_reference_statistics(name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        s, centre, normal, hist, mrow = q.get(timeout=180)
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    g = Golden(name)
    assert np.max(np.abs(centre - g["det_centre"])) <= 1e-9
    assert np.max(np.abs(normal - g["det_normal"])) <= 1e-12
    assert abs(s["SpotSizeSD"] - g["SpotSizeSD"]) <= 1e-9
    assert abs(s["DurationSD"] - g["DurationSD"]) <= 1e-5
    assert abs(s["ETransmission"] - g["ETransmission"]) <= 1e-9
    assert abs(s["SpotSizeSD_w"] - g["SpotSizeSD_w"]) <= 1e-9
    assert abs(s["DurationSD_w"] - g["DurationSD_w"]) <= 1e-5
    assert abs(s["Diameter"] - g["Diameter"]) <= 1e-9
    assert abs(s["NA"] - g["NA"]) <= 1e-11
    # the all-reduced histogram is the histogram of the whole bundle (oracle, numpy.histogram2d)
    from attosecondraytracing_b200.engine import split_histogram
    h = split_histogram(hist, mrow, bins=HIST_BINS, delay_bins=HIST_NT)
    last = g.out(g.n_elements - 1)
    odet = {"centre": g["det_centre"], "normal": g["det_normal"], "refpoint": g["det_refpoint"]}
    w = g["src_I"][np.searchsorted(g["src_num"], last["num"])]
    o = orc.detector_histograms(odet, last["P"], last["U"], last["path"], intensity=w, bins=HIST_BINS,
                                delay_bins=HIST_NT)
    n = last["num"].size
    assert h["spot_count"].sum() == n and h["delay_count"].sum() == n
    assert np.abs(h["spot_count"] - o["spot_count"]).sum() <= 4
    assert np.abs(h["delay_count"] - o["delay_count"]).sum() <= 4
    assert np.allclose(h["x_edges"], o["x_edges"], rtol=0, atol=1e-9)
    assert np.allclose(h["delay_edges"], o["delay_edges"], rtol=0, atol=1e-5)
    same = h["spot_count"] == o["spot_count"]
    assert np.allclose(h["spot_intensity"][same], o["spot_intensity"][same], rtol=0, atol=n * 2.0 ** -26)


def test_shard_ranges_tile_the_bundle():
    from attosecondraytracing_b200 import distributed as ad
    for n in (0, 1, 7, 1000, 10**8 + 3):
        for world in (1, 2, 3, 8):
            got = [ad.shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == n
            for (f0, c0), (f1, _) in zip(got, got[1:]):
                assert f0 + c0 == f1


def test_strided_shards_partition_the_bundle():
    from attosecondraytracing_b200 import distributed as ad
    for n in (0, 1, 7, 1001, 10**7 + 1):
        for world in (1, 2, 3, 8):
            seen = 0
            for r in range(world):
                first, count, stride = ad.shard_strided(n, r, world)
                assert stride == world and first == r
                assert count == len(range(r, n, world))
                seen += count
            assert seen == n


def test_merge_moments_matches_all_reduce_semantics():
    from attosecondraytracing_b200 import distributed as ad
    rng = np.random.default_rng(5)
    rows = rng.normal(size=(3, 24))
    out = ad.merge_moments(list(rows)).numpy()
    assert np.allclose(out[:14], rows[:, :14].sum(axis=0))
    assert np.array_equal(out[[14, 16, 18]], rows[:, [14, 16, 18]].min(axis=0))
    assert np.array_equal(out[[15, 17, 19, 20]], rows[:, [15, 17, 19, 20]].max(axis=0))
