"""The per-ray device code (csrc/art_device.cuh: element frame, intersections, Newton toroid,
Zernike recurrence, reflection) compiled for the HOST by tests/hostcheck and compared with the
reference's golden fixtures.  This is how kernel numerics are checked in the GPU-less build
container; the -m gpu tests repeat the comparison through libart_b200.so on the device."""
import numpy as np
import pytest

import hostcheck_util
from attosecondraytracing_b200 import _cabi
from attosecondraytracing_b200._lowering import LoweredChain
from golden_util import Golden, compare_bundle, golden_names, golden_optical_elements

NAMES = golden_names()


@pytest.mark.parametrize("name", NAMES)
def test_device_code_matches_reference(name):
    g = Golden(name)
    low = LoweredChain([golden_optical_elements(g)])
    flags = _cabi.TRACE_IGNORE_DEFECTS if g.ignore_defects else 0
    res = hostcheck_util.trace(low, g["src_P"], g["src_U"], flags)
    num = g["src_num"]
    for k, b in enumerate(res):
        a = b["alive"]
        compare_bundle(name, k, g.out(k), num[a], b["P"][a], b["U"][a], b["path"][a], b["inc"][a])


@pytest.mark.parametrize("key,tag,ignore", [("toroid", "out", False), ("sphere_cx", "out", False),
                                            ("parabola_hole", "out", False), ("mask", "out", True),
                                            ("sphere_zernike", "out", False), ("sphere_zernike", "outign", True)])
def test_element_frame_functions_match_reference(key, tag, ignore):
    """ReflectionMirrorRayList / TransmitMaskRayList of the reference (rays in the optic's own frame) = one
    element whose frame is the lab frame; the device code on the host against tests/golden/raylist.npz."""
    from golden_util import RayListGolden
    g = RayListGolden()
    low = LoweredChain([[g.identity_element(key)]])
    P, U, num = g.source(key)
    res = hostcheck_util.trace(low, P, U, _cabi.TRACE_IGNORE_DEFECTS if ignore else 0)[0]
    a = res["alive"]
    compare_bundle("raylist_" + key, 0, g.out(key, tag), num[a], res["P"][a], res["U"][a], res["path"][a], res["inc"][a])
