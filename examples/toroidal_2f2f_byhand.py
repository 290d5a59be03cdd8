"""A config script in the style of the reference's examples/ (one toroidal mirror imaging a point source 2f-2f,
the elements placed by hand), written against this package: the only change a user of the reference makes is the
import block.  Run:  python examples/toroidal_2f2f_byhand.py [number of rays]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import attosecondraytracing_b200.ModuleMirror as mmirror  # noqa: E402   (reference: import ART.ModuleMirror as mmirror)
import attosecondraytracing_b200.ModuleOpticalChain as moc  # noqa: E402
import attosecondraytracing_b200.ModuleOpticalElement as moe  # noqa: E402
import attosecondraytracing_b200.ModuleSource as msource  # noqa: E402
import attosecondraytracing_b200.ModuleSupport as msupp  # noqa: E402
from attosecondraytracing_b200.ARTmain import main  # noqa: E402   (reference: from ARTmain import main)


def build(number_rays=1000, incidence_deg=80.0, focal=300.0):
    source_properties = {"Divergence": 15e-3 / 2, "SourceSize": 0, "Wavelength": 50e-6, "DeltaFT": 0.5,
                         "NumberRays": int(number_rays)}
    major, minor = mmirror.ReturnOptimalToroidalRadii(focal, incidence_deg)
    mirror = mmirror.MirrorToroidal(major, minor, msupp.SupportRectangle(120, 30))
    element = moe.OpticalElement(mirror, np.zeros(3), np.array([0.0, 0.0, 1.0]), np.array([1.0, 0.0, 0.0]))
    a = np.deg2rad(incidence_deg)
    source_point = 2 * focal * np.array([np.sin(a), 0.0, np.cos(a)])
    rays = msource.PointSource(source_point, -source_point, source_properties["Divergence"], source_properties["NumberRays"])
    rays = msource.ApplyGaussianIntensityToRayList(rays, 1 / np.e**2)
    chain = moc.OpticalChain(rays, [element], "single toroidal mirror, 2f-2f, placed by hand")
    detector_options = {"ReflectionNumber": -1, "ManualDetector": False, "DistanceDetector": 2 * focal,
                        "AutoDetectorDistance": False, "OptFor": "intensity"}
    analysis_options = {"verbose": True, "save_results": False}
    return chain, source_properties, detector_options, analysis_options


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    kept = main(*build(n))
    print("SpotSizeSD %.4g um, DurationSD %.4g fs, ETransmission %.1f %%" % (
        kept["SpotSizeSD"][0] * 1e3, kept["DurationSD"][0], kept["ETransmission"][0]))
