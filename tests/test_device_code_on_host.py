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
