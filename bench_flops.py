"""Canonical FLOP model of SURVEY.md 8(d) (add/sub/mul/div/sqrt = 1, FMA = 2, compares free,
incidence angle excluded) so that bench.py and a reader use the same number for the FP64 fraction."""

FRAME = 33          # lab <-> element frame change
ACT = {"Plane Mirror": 13, "Mask": 13, "SphericalCC Mirror": 72, "SphericalCX Mirror": 72, "Parabolic Mirror": 72,
       "Ellipsoidal Mirror": 77, "CylindricalCC Mirror": 67, "CylindricalCX Mirror": 67, "Toroidal Mirror": 163}
DETECTOR = 60       # per surviving ray


def element_flops(oe, ignore_defects=True):
    optic = oe.type
    f = FRAME + ACT[optic.type]
    for d in getattr(optic, "DeformationList", []):
        N = max(int(d.max_order), 2)
        T = (N + 1) * (N + 2) // 2
        f += (11 * T if ignore_defects else 11 * T + 24 * T) + 26
    return f


def chain_flops(oes, entering, n_surv, ignore_defects=True):
    """entering[k] = rays entering element k; n_surv = rays leaving the last element."""
    total = sum(n * element_flops(oe, ignore_defects) for oe, n in zip(oes, entering))
    return float(total + n_surv * (FRAME + DETECTOR))
