"""Run under torchrun on >= 2 GPUs of one node (tests/test_gpu_multi.py does so when they are there):
the peer-memory exchange kernel (art_peer_exchange) against the NCCL path it replaces -- sums in rank order
vs NCCL's all-reduce (equal to rounding), merged moments bit for bit, identical rows on every rank, detectors
placed in the same kernel, many back-to-back epochs, and a CUDA-graph-captured sequence."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from attosecondraytracing_b200 import _cabi, engine  # noqa: E402
from attosecondraytracing_b200 import distributed as ad  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    peer = ad.PeerExchange.create(dev)
    assert peer is not None, "symmetric memory rendezvous failed"
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    for nv in (1, 3, 64):
        for it in range(20):
            # central rows: unit-ish direction sums, points, path, counts
            c = (torch.rand((nv, _cabi.CENTRAL_LEN), generator=g, dtype=torch.float64) + 0.5).to(dev)
            c[:, 7] = 1000 + rank  # N
            ref = c.clone()
            dist.all_reduce(ref, op=dist.ReduceOp.SUM)
            det_ref = torch.empty((nv, _cabi.DETECTOR_DOUBLES), dtype=torch.float64, device=dev)
            _cabi.check(_cabi.lib().art_detector_autoplace(ref.data_ptr(), 123.0, nv, det_ref.data_ptr(),
                                                           torch.cuda.current_stream().cuda_stream))
            det = torch.empty_like(det_ref)
            peer.all_reduce_central(c, 123.0, det)
            assert torch.allclose(c, ref, rtol=1e-14, atol=0), (nv, it)
            assert torch.allclose(det, det_ref, rtol=1e-12, atol=1e-12), (nv, it)
            rows = [torch.empty_like(c) for _ in range(world)]
            dist.all_gather(rows, c)
            assert all(torch.equal(r, rows[0]) for r in rows), "ranks disagree on the reduced central rows"
            # moments rows
            m = torch.randn((nv, _cabi.MOMENTS_LEN), generator=g, dtype=torch.float64).to(dev)
            if rank == world - 1 and it % 5 == 0:  # a rank without survivors keeps the reduction identities
                m[:, 0:14] = 0.0
                m[:, list(_cabi.MOMENT_MIN)] = float("inf")
                m[:, list(_cabi.MOMENT_MAX)] = float("-inf")
            ref = m.clone()
            ad.all_reduce_moments(ref)  # NCCL all-gather + art_moments_merge
            peer.all_reduce_moments(m)
            assert torch.equal(m[:, :21], ref[:, :21]), (nv, it)
    # graph-captured sequence: the epoch lives in the buffer, so replays keep working
    c = torch.ones((1, _cabi.CENTRAL_LEN), dtype=torch.float64, device=dev)
    m = torch.ones((1, _cabi.MOMENTS_LEN), dtype=torch.float64, device=dev)
    det = torch.empty((1, _cabi.DETECTOR_DOUBLES), dtype=torch.float64, device=dev)
    cin, min_ = c.clone(), m.clone()

    def seq():
        c.copy_(cin)
        m.copy_(min_)
        peer.all_reduce_central(c, 10.0, det)
        peer.all_reduce_moments(m)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        seq()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        seq()
    for _ in range(50):
        graph.replay()
    torch.cuda.synchronize()
    assert float(c[0, 0]) == world and float(m[0, 0]) == world
    assert peer.status() == 0
    # end to end: every rank feeds its round-robin shard of the golden cfg3 bundle (host buffers) to
    # art_run_host_sharded; all ranks get the statistics of the WHOLE bundle, which the reference pinned
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from golden_util import Golden, golden_optical_elements
    from attosecondraytracing_b200.ModuleOpticalRay import RayBundle
    gd = Golden("cfg3_2tor")
    chain = engine.DeviceChain(golden_optical_elements(gd), device=dev)
    sel = np.arange(rank, gd["src_P"].shape[0], world)
    host = RayBundle.from_numpy(gd["src_P"][sel], gd["src_U"][sel], intensity=gd["src_I"][sel], device="cpu")
    mom, cen, det = chain.run_host(host, gd.spec["detector_distance"], ignore_defects=gd.ignore_defects, peer=peer)
    s = engine.summary_from_moments(mom, cen)
    assert abs(s["SpotSizeSD"] - gd["SpotSizeSD"]) <= 1e-9 and abs(s["DurationSD"] - gd["DurationSD"]) <= 1e-5, s
    assert abs(s["ETransmission"] - gd["ETransmission"]) <= 1e-9
    assert np.max(np.abs(np.array(det.centre[:]) - gd["det_centre"])) <= 1e-9
    rows = [torch.empty(_cabi.MOMENTS_LEN, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(rows, torch.from_numpy(mom).to(dev))
    assert all(torch.equal(r, rows[0]) for r in rows), "ranks disagree on the whole-bundle moments"
    # the exchange that folds the per-block rows itself == fold launch + plain exchange (another, equally fixed,
    # summation order: equal to rounding; extents exactly)
    from attosecondraytracing_b200.ModuleOpticalRay import RayBundle as _RB
    dsrc = _RB.from_numpy(gd["src_P"][sel], gd["src_U"][sel], intensity=gd["src_I"][sel], device=dev)
    dist_ = gd.spec["detector_distance"]
    outs_a, cen_a = chain.trace(dsrc, ignore_defects=gd.ignore_defects, history=False)
    det_a = torch.empty((1, _cabi.DETECTOR_DOUBLES), dtype=torch.float64, device=dev)
    peer.all_reduce_central(cen_a, dist_, det_a)
    mom_a, _, _, _ = chain.moments(outs_a[0], det_a, intensity=dsrc.col("intensity"))
    peer.all_reduce_moments(mom_a)
    cen_b = torch.empty_like(cen_a)
    det_b = torch.empty_like(det_a)
    mom_b = torch.empty_like(mom_a)
    outs_b, _ = chain.trace(dsrc, ignore_defects=gd.ignore_defects, history=False, central=cen_b, fold=False)
    peer.all_reduce_central(cen_b, dist_, det_b, chain=chain)
    chain.moments(outs_b[0], det_b, intensity=dsrc.col("intensity"), fold=False)
    peer.all_reduce_moments(mom_b, chain=chain)
    torch.cuda.synchronize()
    assert torch.allclose(cen_a, cen_b, rtol=1e-13, atol=0) and torch.allclose(det_a, det_b, rtol=1e-12, atol=1e-12)
    assert torch.allclose(mom_a[:, :14], mom_b[:, :14], rtol=1e-11, atol=1e-12) and torch.equal(mom_a[:, 14:21], mom_b[:, 14:21])
    rows_b = [torch.empty_like(mom_b) for _ in range(world)]
    dist.all_gather(rows_b, mom_b)
    assert all(torch.equal(r, rows_b[0]) for r in rows_b), "ranks disagree after the folding exchange"
    # the same from the source DESCRIPTION: every rank generates its own round-robin share on the device, the axis /
    # extent of the intensity profile are combined over the ranks inside the call (art_run_source_host)
    import attosecondraytracing_b200.ModuleSource as msrc
    n_src = gd["src_P"].shape[0]
    desc = msrc.source_descriptor(dict(gd.spec["source"]), first=rank, stride=world)
    mom2, cen2, det2 = chain.run_source(desc, gd.spec["detector_distance"], ignore_defects=gd.ignore_defects, peer=peer)
    s2 = engine.summary_from_moments(mom2, cen2)
    for key in ("SpotSizeSD", "SpotSizeSD_w", "ETransmission"):
        assert abs(s2[key] - gd[key]) <= 1e-9, (key, s2[key], float(gd[key]))
    assert abs(s2["DurationSD"] - gd["DurationSD"]) <= 1e-5 and abs(s2["DurationSD_w"] - gd["DurationSD_w"]) <= 1e-5
    assert s2["n_rays"] == gd.out(gd.n_elements - 1)["num"].size
    # a rank with an EMPTY shard still takes part in the exchanges (rank 0 holds everything, the others nothing)
    if rank == 0:
        host0 = RayBundle.from_numpy(gd["src_P"], gd["src_U"], intensity=gd["src_I"], device="cpu")
    else:
        host0 = RayBundle.from_numpy(gd["src_P"][:0], gd["src_U"][:0], intensity=gd["src_I"][:0], device="cpu")
    mom3, cen3, _ = chain.run_host(host0, gd.spec["detector_distance"], ignore_defects=gd.ignore_defects, peer=peer)
    s3 = engine.summary_from_moments(mom3, cen3)
    assert abs(s3["SpotSizeSD"] - gd["SpotSizeSD"]) <= 1e-9 and s3["n_rays"] == s2["n_rays"]
    st = peer.stats()
    assert st["exchanges"] > 0 and st["kernel_us"] > 0
    chain.close()
    dist.barrier()
    if rank == 0:
        print("peer exchange ok on %d ranks (exchange kernel %.1f us, of which %.1f us waiting for the peers, "
              "averaged over %d exchanges on rank 0)" % (world, st["kernel_us"], st["wait_us"], st["exchanges"]), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
