"""The extended-precision arbiter (oracle/art_oracle_ld.py, np.longdouble = x87 80-bit) against
  (1) the reference's golden fixtures -- this measures the REFERENCE's own float64 noise,
  (2) the float64 oracle, and
  (3) the device code compiled for the host (tests/hostcheck),
scene by scene.  On the 5 m-arm telescope (BASELINE config 5) the reference is 1e-8 mm away from the
extended-precision evaluation -- ten times the north star's 1e-9 mm bar -- while the device code is within
3e-12 mm of it: the 3e-8 mm tolerance those fixtures need against the reference (tests/golden_util.py) is the
reference's noise, not the kernel's.  The `-m gpu` twin of (3) is
tests/test_gpu_parity.py::test_trace_matches_extended_precision_arbiter."""
import numpy as np
import pytest

import art_oracle as orc
import art_oracle_ld as ld
import hostcheck_util
from attosecondraytracing_b200 import _cabi
from attosecondraytracing_b200._lowering import LoweredChain
from golden_util import DELAY_TOL_FS, Golden, golden_names, golden_optical_elements

pytestmark = pytest.mark.skipif(not ld.available(), reason="np.longdouble is not wider than float64 here")


def _has_gridmap(g):
    return any(d.get("kind") == "gridmap" for o in g.spec["derived_optics"] for d in (o.get("defects") or []))


NAMES = [n for n in golden_names() if not _has_gridmap(Golden(n))]
DEVICE_TOL_MM = 1e-10   # device code vs extended precision: 10x tighter than the north-star bar, on EVERY scene


def _ld_trace(g):
    return ld.trace_chain(g["src_P"], g["src_U"], g.oracle_elements(), ignore_defects=g.ignore_defects,
                          numbers=g["src_num"])


@pytest.mark.parametrize("name", NAMES)
def test_device_code_and_oracle_sit_on_the_extended_precision_side(name):
    g = Golden(name)
    t = _ld_trace(g)
    o64 = orc.trace_chain(g["src_P"], g["src_U"], g.oracle_elements(), ignore_defects=g.ignore_defects,
                          numbers=g["src_num"])
    low = LoweredChain([golden_optical_elements(g)])
    dev = hostcheck_util.trace(low, g["src_P"], g["src_U"], _cabi.TRACE_IGNORE_DEFECTS if g.ignore_defects else 0)
    for k in range(g.n_elements):
        ref = g.out(k)
        assert np.array_equal(t[k]["number"], ref["num"]), f"{name}: extended precision keeps other rays after element {k}"
        if ref["num"].size == 0:
            continue
        a = dev[k]["alive"]
        d_dev = float(np.max(np.abs(dev[k]["P"][a] - t[k]["P"])))
        d_o64 = float(np.max(np.abs(o64[k]["P"] - t[k]["P"])))
        d_ref = float(np.max(np.abs(ref["P"] - t[k]["P"])))
        assert d_dev <= DEVICE_TOL_MM, (name, k, d_dev)
        assert d_o64 <= 1e-9, (name, k, d_o64)
        assert float(np.max(np.abs(dev[k]["path"][a] - t[k]["path"]))) <= 2 * DEVICE_TOL_MM
        # the reference itself: within the north-star bar except on the 5 m arms, where IT is the noisy side
        if name.startswith(("cfg5", "tele")):
            assert d_ref > 10 * d_dev, (name, k, d_ref, d_dev)
            assert d_ref <= 3e-8
        else:
            assert d_ref <= 1e-9, (name, k, d_ref)


@pytest.mark.parametrize("name", ["cfg5_tele", "cfg5_sub_v300", "tele_yaw_shift", "cfg3_2tor", "cfg1_par"])
def test_detector_response_against_extended_precision(name):
    """Delays and in-plane points of the float64 oracle (the checker of the GPU detector kernels) against the
    extended-precision evaluation: delays within 0.01 as, points within 1e-9 mm."""
    g = Golden(name)
    t = _ld_trace(g)[-1]
    o = orc.trace_chain(g["src_P"], g["src_U"], g.oracle_elements(), ignore_defects=g.ignore_defects)[-1]
    det_ld = ld.detector_autoplace(t["P"], t["U"], g.spec["detector_distance"])
    xy_ld, dl_ld = ld.detector_response(det_ld, t["P"], t["U"], t["path"])
    det = orc.detector_autoplace(o["P"], o["U"], g.spec["detector_distance"])
    xy = orc.detector_points2d_centre(det, o["P"], o["U"])
    dl = orc.detector_delays(det, o["P"], o["U"], o["path"])
    assert float(np.max(np.abs(dl - dl_ld))) <= DELAY_TOL_FS
    assert float(np.max(np.abs(xy - xy_ld))) <= 1e-9
    # and the reference's own delays agree with extended precision everywhere (delays are fine)
    assert float(np.max(np.abs(g["det_delays"] - dl_ld))) <= DELAY_TOL_FS
