"""TEST INFRASTRUCTURE ONLY -- the scene of examples/toroidal_2f2f_byhand.py built with the UNMODIFIED reference's
modules (1000 rays) and analysed as its driver does (Detector.autoplace + GetResultSummary + getETransmission),
written to tests/golden/example_byhand.npz.   python oracle/gen_golden_example.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402
import load_reference as lr  # noqa: E402


def main():
    R = gg.ref()
    n, incidence, focal, div = 1000, 80.0, 300.0, 15e-3 / 2
    major, minor = R.mmirror.ReturnOptimalToroidalRadii(focal, incidence)
    mirror = R.mmirror.MirrorToroidal(major, minor, R.msupp.SupportRectangle(120, 30))
    element = R.moe.OpticalElement(mirror, np.zeros(3), np.array([0.0, 0.0, 1.0]), np.array([1.0, 0.0, 0.0]))
    a = np.deg2rad(incidence)
    S = 2 * focal * np.array([np.sin(a), 0.0, np.cos(a)])
    rays = R.msource.PointSource(S, -S, div, n)
    rays = R.msource.ApplyGaussianIntensityToRayList(rays, 1 / np.e**2)
    chain = R.moc.OpticalChain(rays, [element], "by hand")
    with lr.quiet():
        out = chain.get_output_rays()[-1]
        det = R.mdet.Detector(element.position)
        det.autoplace(out, 2 * focal)
        spot, dur = R.mplots.GetResultSummary(det, out, False)
        et = R.mplots.getETransmission(chain.source_rays, out)
    np.savez_compressed(os.path.join(gg.GOLDEN_DIR, "example_byhand.npz"), n=np.array(n), SpotSizeSD=np.array(spot),
                        DurationSD=np.array(dur), ETransmission=np.array(et), det_centre=np.array(det.centre),
                        det_normal=np.array(det.normal), survivors=np.array(len(out)))
    print(len(out), spot, dur, et)


if __name__ == "__main__":
    main()
