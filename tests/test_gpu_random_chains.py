"""`-m gpu` twin of tests/test_random_chains_vs_reference.py: the KERNEL (through the C ABI, per-element history and
the final-only trace) on the same seeded random chains, against the live reference (the copy in oracle/_ref travels
to the GPU box) for survivors and against the extended-precision arbiter for the numbers."""
import numpy as np
import pytest

import art_oracle_ld as ld
import gen_golden
import ref_runner
from golden_util import build_optic
from test_random_chains_vs_reference import _scene

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_runner.available(), reason="no copy of the reference on this machine")]


@pytest.mark.parametrize("seed", range(20))
def test_kernel_on_random_chains(seed):
    from attosecondraytracing_b200 import engine as eng
    from attosecondraytracing_b200.ModuleOpticalRay import RayBundle
    import attosecondraytracing_b200.ModuleOpticalElement as moe
    import load_reference as lr
    R = ref_runner.ref()
    scene = _scene(seed)
    chain_ref = ref_runner.build_chain(scene)
    n = scene["source"]["NumberRays"]
    rays = ref_runner.subset_source_rays(scene, n, np.arange(n))
    src_P = np.array([r.point for r in rays], dtype=np.float64)
    src_U = np.array([r.vector for r in rays], dtype=np.float64)
    with lr.quiet():
        out = R.mp.RayTracingCalculation(rays, chain_ref.optical_elements, IgnoreDefects=True)
    oes, els = [], []
    for spec, roe in zip(scene["optics"], chain_ref.optical_elements):
        pose = [np.asarray(getattr(roe, a), dtype=np.float64) for a in ("position", "normal", "majoraxis")]
        oes.append(moe.OpticalElement(build_optic(dict(spec, support=tuple(spec["support"]))), *pose))
        optic = gen_golden.derived_optic(spec, roe.type)
        optic["support"] = tuple(optic["support"])
        els.append({"optic": optic, "position": pose[0], "normal": pose[1], "majoraxis": pose[2]})
    arb = ld.trace_chain(src_P, src_U, els, ignore_defects=True) if ld.available() else None
    chain = eng.DeviceChain(oes)
    src = RayBundle.from_numpy(src_P, src_U, device="cuda")
    outs, _ = chain.trace(src, ignore_defects=True, history=True)
    final, _ = chain.trace(src, ignore_defects=True, history=False)
    kinds = "+".join(s["kind"] for s in scene["optics"])
    for k, ref_list in enumerate(out):
        ref_num = np.array([r.number for r in ref_list], dtype=np.int64)
        d = outs[k].to_numpy()
        assert np.array_equal(d["number"], ref_num), (seed, kinds, k, np.setxor1d(d["number"], ref_num)[:8])
        if ref_num.size == 0:
            continue
        ref_P = np.array([r.point for r in ref_list]).reshape(-1, 3)
        assert np.max(np.abs(d["P"] - ref_P)) <= 1e-8, (seed, kinds, k)
        if arb is not None:
            assert float(np.max(np.abs(d["P"] - arb[k]["P"]))) <= 1e-10, (seed, kinds, k)
            assert float(np.max(np.abs(d["path"] - arb[k]["path"]))) <= 2e-10, (seed, kinds, k)
    # the final-only trace stores the same bundle as the last entry of the history
    f, h = final[0].to_numpy(), outs[-1].to_numpy()
    assert np.array_equal(f["number"], h["number"]) and np.array_equal(f["P"], h["P"]) and np.array_equal(f["U"], h["U"])
    chain.close()
