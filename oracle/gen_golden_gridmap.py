"""
TEST INFRASTRUCTURE ONLY -- pins the gridded-defect paths the reference cannot run as shipped.

`MeasuredMap.__init__` (ART/ModuleDefects.py:44) and `Fourrier.get_normal` (:125) raise under the numpy of
this image; oracle/make_ref.py keeps a second copy of the reference package, oracle/_ref/ART_np2, with the
documented two-line compatibility patch applied (one call signature, one `.flatten()`).  This script runs
THAT copy -- everything else unmodified -- and writes the fixtures the stock reference cannot produce:

    par_fourier_def    Fourrier + Zernike defects, IgnoreDefects=False (slope path through get_normal)
    par_measured_ign   MeasuredMap defect, IgnoreDefects=True  (height path)
    par_measured_def   MeasuredMap defect, IgnoreDefects=False (height + slope path)

    python oracle/make_ref.py && python oracle/gen_golden_gridmap.py

The fixtures' spec records `reference_patch` so that nobody mistakes them for output of the stock tree; the
IgnoreDefects=True Fourier fixture (par_fourier_ign) stays the stock reference's (oracle/gen_golden.py).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
NP2_ROOT = os.path.join(HERE, "_ref", "ART_np2")
if not os.path.isdir(os.path.join(NP2_ROOT, "ART")):
    sys.exit("oracle/_ref/ART_np2 missing: run python oracle/make_ref.py first")
os.environ["ART_REFERENCE_ROOT"] = NP2_ROOT   # before load_reference is imported
sys.path.insert(0, HERE)

import gen_golden as gg  # noqa: E402
import scenes as sc  # noqa: E402

PATCH_NOTE = "numpy>=2 compatibility patch of ART/ModuleDefects.py (2 lines, oracle/make_ref.py)"


def main():
    todo = [("par_fourier", False, "par_fourier_def"), ("par_measured", True, "par_measured_ign"),
            ("par_measured", False, "par_measured_def")]
    for scene_name, ign, tag in todo:
        scene = sc.resolve(scene_name)
        data, counts, dt = gg.run_scene(scene, ignore_defects=ign, extra={"reference_patch": PATCH_NOTE})
        p = gg.save(tag, data)
        print(f"{tag:20s} survivors {counts}  ref trace {dt:.2f}s -> {os.path.relpath(p)}")


if __name__ == "__main__":
    main()
