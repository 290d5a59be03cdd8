"""Full-size checks on the GPU (BASELINE.json sizes): the rows of the full 10^6..10^8-ray bundles at
the seeded subset indices against what the reference produced for exactly those rays
(tests/golden/*_sub*.npz), plus size-independent properties of the path: permutation and sharding
invariance, chaining two partial chains == one chain (to rounding), merged shard moments == full moments, and
the known answers of the geometry (parabola focus, ellipsoid focus-to-focus path 2a, plane mirror)."""
import numpy as np
import pytest
import torch

from golden_util import DELAY_TOL_FS, Golden, compare_bundle, golden_optical_elements, point_tol

pytestmark = pytest.mark.gpu


def _mods():
    from attosecondraytracing_b200 import engine
    import attosecondraytracing_b200.ModuleSource as msrc
    from attosecondraytracing_b200.ModuleOpticalRay import RayBundle
    return engine, msrc, RayBundle


def _gather(bundle, idx, RayBundle, intensity=None):
    """Small bundle holding rows `idx` of a big one (keeps alive flags)."""
    names = [n for n in bundle._names]
    small = RayBundle(idx.numel(), device=bundle.device, columns=names, with_alive=bundle.alive is not None)
    for n in names:
        small.col(n).copy_(bundle.col(n)[idx])
    if bundle.alive is not None:
        small.alive.copy_(bundle.alive[idx])
    small.origin = bundle.origin
    small.number = idx.clone()
    small.invalidate()
    return small


@pytest.mark.parametrize("name", ["cfg2_sub", "cfg3_sub", "cfg4_sub_ign", "cfg4_sub_def", "cfg5_sub_v0", "cfg5_sub_v300",
                                  "cfg5_sub_v1023"])
def test_full_size_bundle_rows_match_reference_subset(name):
    eng, msrc, RayBundle = _mods()
    g = Golden(name)
    n_full = int(g.spec["subset_of"])
    sp = dict(g.spec["source"])
    sp["NumberRays"] = n_full
    src = msrc.synthetic_source(sp, device="cuda")
    assert src.n == (n_full - 1 if sp["Divergence"] == 0 else n_full)
    chain = eng.DeviceChain(golden_optical_elements(g))
    outs, central = chain.trace(src, ignore_defects=g.ignore_defects, history=False)
    final = outs[0]
    idx = torch.from_numpy(g["src_num"]).cuda()
    # the device-generated source rows are the reference's source rays
    srows = _gather(src, idx, RayBundle).to_numpy()
    assert np.max(np.abs(srows["P"] - g["src_P"])) <= 1e-12 and np.max(np.abs(srows["U"] - g["src_U"])) <= 1e-14
    assert np.max(np.abs(srows["intensity"] - g["src_I"])) <= 1e-9
    sub = _gather(final, idx, RayBundle)
    d = sub.to_numpy()
    k = g.n_elements - 1
    compare_bundle(name, k, g.out(k), d["number"], d["P"], d["U"], d["path"], d["incidence"])
    # interaction count of the full bundle is consistent with the survivors
    entering, surv = chain.count_entering(src, ignore_defects=g.ignore_defects)
    assert int(surv[0]) == len(final) == int(central.cpu()[0, 7])
    assert int(entering[0, 0]) == src.n and all(int(entering[0, i]) >= int(entering[0, i + 1]) for i in range(k))
    # detector response of the subset (autoplace'd on the subset, as the reference did)
    if "det_delays" in g:
        inten = src.col("intensity")[idx].contiguous()
        c2 = torch.zeros((1, 10), dtype=torch.float64, device="cuda")
        alive = sub.alive.bool()
        for j, col in enumerate(("ux", "uy", "uz", "px", "py", "pz", "path")):
            c2[0, j] = sub.col(col)[alive].sum()
        c2[0, 7] = alive.sum()
        det = chain.autoplace(c2, g.spec["detector_distance"])
        mom, x, y, l = chain.moments(sub, det, intensity=inten, want_points=True)
        delays = chain.delays(l, sub.alive, det, mom)
        torch.cuda.synchronize()
        dl = delays[alive].cpu().numpy()
        assert np.max(np.abs(dl - g["det_delays"])) <= DELAY_TOL_FS
        s = eng.summary_from_moments(mom.cpu().numpy()[0])
        assert abs(s["SpotSizeSD"] - g["SpotSizeSD"]) <= point_tol(name)
        assert abs(s["DurationSD"] - g["DurationSD"]) <= DELAY_TOL_FS
    chain.close()


def test_permutation_sharding_and_chaining_invariance_10M():
    """cfg3 chain on 10^7 rays: (a) a permuted bundle gives the permuted result bit for bit; (b) two
    half-bundle shards give the rows of the full trace and their merged moments equal the full moments;
    (c) tracing element 0 and then elements 1-2 from the stored intermediate bundle equals one call."""
    eng, msrc, RayBundle = _mods()
    from attosecondraytracing_b200 import distributed as ad
    g = Golden("cfg3_2tor")
    n = 10_000_000
    sp = dict(g.spec["source"])
    sp["NumberRays"] = n
    src = msrc.synthetic_source(sp, device="cuda")
    oes = golden_optical_elements(g)
    chain = eng.DeviceChain(oes)
    outs, central = chain.trace(src, history=False)
    full = outs[0]
    det = chain.autoplace(central, g.spec["detector_distance"])
    mom, _, _, _ = chain.moments(full, det, intensity=src.col("intensity"))
    cols = ("px", "py", "pz", "ux", "uy", "uz", "path", "incidence")
    # (a) permutation
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
    psrc = _gather(src, perm, RayBundle)
    psrc.number = None
    pout, _ = chain.trace(psrc, history=False)
    assert torch.equal(pout[0].alive, full.alive[perm])
    live = pout[0].alive.bool()
    for c in cols:
        assert torch.equal(pout[0].col(c)[live], full.col(c)[perm][live]), c
    # (b) shards
    rows = []
    for r in range(2):
        first, count = ad.shard_range(n, r, 2)
        part = _gather(src, torch.arange(first, first + count, device="cuda"), RayBundle)
        po, pc = chain.trace(part, history=False)
        assert torch.equal(po[0].alive, full.alive[first:first + count])
        lv = po[0].alive.bool()
        for c in cols:
            assert torch.equal(po[0].col(c)[lv], full.col(c)[first:first + count][lv]), c
        pm, _, _, _ = chain.moments(po[0], det, intensity=part.col("intensity"))
        rows.append(pm[0].cpu())
    merged = ad.merge_moments(rows).numpy()
    m = mom[0].cpu().numpy()
    assert np.allclose(merged[:14], m[:14], rtol=1e-11, atol=1e-9)
    assert np.array_equal(merged[14:21], m[14:21])
    # (b') binned detector response at full size: every survivor lands in exactly one bin, the binned
    # intensity adds up to the summed intensity of the survivors, and the histograms of the two shards
    # (binned against the extents of the WHOLE bundle) add to the histogram of the whole bundle exactly
    n_surv = int(full.alive.sum())
    for bins, nt in (((64, 64), 128), ((300, 200), 4096)):  # shared-memory bins / global atomics
        hist = chain.histogram(full, det, mom, bins=bins, delay_bins=nt, intensity=src.col("intensity"))
        h = eng.split_histogram(hist.cpu().numpy(), m, bins=bins, delay_bins=nt)
        assert h["spot_count"].sum() == n_surv and h["delay_count"].sum() == n_surv
        w_out = float(src.col("intensity")[full.alive.bool()].sum())
        assert abs(h["spot_intensity"].sum() - w_out) <= n_surv * 2.0 ** -27
        assert abs(h["delay_intensity"].sum() - w_out) <= n_surv * 2.0 ** -27
        total = torch.zeros_like(hist)
        for r in range(2):
            first, count = ad.shard_range(n, r, 2)
            part = _gather(src, torch.arange(first, first + count, device="cuda"), RayBundle)
            po, _ = chain.trace(part, history=False)
            total += chain.histogram(po[0], det, mom, bins=bins, delay_bins=nt, intensity=part.col("intensity"))
        assert torch.equal(total, hist)
    # (c) chaining: element 0, then elements 1..2 starting from the stored bundle (path + alive carried)
    first_el = eng.DeviceChain(oes[:1])
    rest = eng.DeviceChain(oes[1:])
    o1, _ = first_el.trace(src, history=False)
    o2, c2 = rest.trace(o1[0], history=False)
    assert torch.equal(o2[0].alive, full.alive)
    lv = full.alive.bool()
    # one call hands the rays from element to element through ONE composed affine map (apply_element), the split
    # chain goes through the lab frame in between: equal to rounding (1e-13 mm per step, amplified ~100x by two
    # reflections at 80 degrees grazing incidence), not bit for bit
    for c in cols:
        assert float((o2[0].col(c)[lv] - full.col(c)[lv]).abs().max()) <= 1e-10, c
    assert float((c2[0, :8] - central[0, :8]).abs().max()) <= 1e-10 * float(central[0, 7])  # sums of N such rows
    assert c2[0, 7] == central[0, 7]  # the stored intermediate bundle carries no intensities
    for ch in (chain, first_el, rest):
        ch.close()


def test_known_answers_of_the_geometry():
    """Physics the surfaces must obey: an on-axis parabola focuses a collimated bundle
    to a point with equal paths; an ellipsoid images focus to focus with path 2a; a plane mirror adds
    paths and preserves angles."""
    eng, msrc, RayBundle = _mods()
    import attosecondraytracing_b200.ModuleMirror as mm
    import attosecondraytracing_b200.ModuleSupport as ms
    import attosecondraytracing_b200.ModuleOpticalElement as moe
    n = 200_001
    rng = np.random.default_rng(11)
    # parabola, on axis: rays along -z onto z = (x^2+y^2)/(2p), focus at (0,0,p/2)
    f = 50.0
    par = mm.MirrorParabolic(f, 0, ms.SupportRound(20))
    oe = moe.OpticalElement(par, np.zeros(3), np.array([0.0, 0, 1.0]), np.array([1.0, 0, 0]))
    xy = rng.uniform(-12, 12, size=(n, 2))
    P = np.column_stack([xy, np.full(n, 300.0)])
    U = np.tile([0.0, 0, -1.0], (n, 1))
    chain = eng.DeviceChain([oe])
    out, _ = chain.trace(RayBundle.from_numpy(P, U, device="cuda"), history=False)
    d = out[0].to_numpy()
    assert d["number"].size == n
    focus = np.array([0.0, 0.0, par.p / 2])
    t = (focus[2] - d["P"][:, 2]) / d["U"][:, 2]
    hit = d["P"] + d["U"] * t[:, None]
    assert np.max(np.abs(hit - focus)) <= 1e-11
    total = d["path"] + t
    assert np.max(np.abs(total - total[0])) <= 1e-11  # Fermat: equal optical paths to the focus
    chain.close()
    # ellipsoid: rays from focus F1 = (-c,0,0) reach F2 = (+c,0,0) after path 2a
    a, b = 400.0, 150.0
    ell = mm.MirrorEllipsoidal(ms.SupportRectangle(2000, 2000), a, b)
    ctr = ell.get_centre()
    oe = moe.OpticalElement(ell, ctr.copy(), np.array([0.0, 0, 1.0]), np.array([1.0, 0, 0]))  # element frame = lab frame
    c = np.sqrt(a * a - b * b)
    ang = rng.uniform(-0.3, 0.3, size=(n, 2))
    Uc = np.column_stack([np.sin(ang[:, 0]), np.sin(ang[:, 1]) * 0.2, -np.cos(ang[:, 0])])
    Pc = np.tile([-c, 0.0, 0.0], (n, 1))
    chain = eng.DeviceChain([oe])
    out, _ = chain.trace(RayBundle.from_numpy(Pc, Uc, device="cuda"), history=False)
    d = out[0].to_numpy()
    assert d["number"].size > n // 2
    F2 = np.array([c, 0.0, 0.0])
    to_f2 = F2 - d["P"]
    dist = np.linalg.norm(to_f2, axis=1)
    assert np.max(np.abs(np.cross(d["U"], to_f2 / dist[:, None]))) <= 1e-12  # heading for the second focus
    assert np.max(np.abs(d["path"] + dist - 2 * a)) <= 1e-10
    chain.close()
    # plane mirror at 45 degrees: angle of incidence preserved, image of the source point behind the mirror
    pl = mm.MirrorPlane(ms.SupportRound(1e4))
    nrm = np.array([-1.0, 0, 1.0]) / np.sqrt(2)
    oe = moe.OpticalElement(pl, np.array([100.0, 0, 0]), nrm, np.array([1.0, 0, 1.0]) / np.sqrt(2))
    Us = np.column_stack([np.ones(n), rng.uniform(-0.05, 0.05, n), rng.uniform(-0.05, 0.05, n)])
    Us /= np.linalg.norm(Us, axis=1)[:, None]
    chain = eng.DeviceChain([oe])
    out, _ = chain.trace(RayBundle.from_numpy(np.zeros((n, 3)), Us, device="cuda"), history=False)
    d = out[0].to_numpy()
    image = 2 * np.dot(np.array([100.0, 0, 0]), nrm) * nrm  # mirror image of the origin
    back = d["P"] - d["U"] * d["path"][:, None]
    assert np.max(np.abs(back - image)) <= 1e-10
    assert np.max(np.abs(d["incidence"] - np.arccos(np.clip(-(Us @ nrm), -1, 1)))) <= 1e-12
    chain.close()
