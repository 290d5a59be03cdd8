"""
TEST / BENCH INFRASTRUCTURE ONLY -- builds oracle/_ref/, a travelling copy of the UNMODIFIED reference.

    python oracle/make_ref.py            # needs /root/reference (the build container)

The reference (mightymightys/AttosecondRaytracing, ART v0.93) is pure Python without setup.py /
pyproject.toml, so it cannot be pip-installed; its hot-path modules are importable as they lie once
numpy-quaternion and the plotting stack are shimmed (oracle/refshim).  /root/reference does not exist
on the GPU box, therefore this recipe copies the package directory, byte for byte, into the git-ignored
(but NOT gpurun-ignored) oracle/_ref/ART so that `bench.py --impl reference` and bench.py's
`cpu_baseline` leg can time the literal reference on the box's host cores.  Nothing is copied into
tracked paths: oracle/_ref/ is listed in .gitignore.

Layout written:
    oracle/_ref/ART/*.py            the reference package, unmodified (sha256 of every file in MANIFEST.json)
    oracle/_ref/ART_np2/ART/*.py    the same with the two-line numpy>=2 compatibility patch below applied to
                                    ModuleDefects.py -- used ONLY by oracle/gen_golden_gridmap.py to pin the
                                    gridded-defect slope path, never by the bench
    oracle/_ref/MANIFEST.json       source path, file hashes, the patch

The numpy >= 2 compatibility patch (ART/ModuleDefects.py), each replacing one line by the evident intent:
  1. MeasuredMap.__init__ (:42)  `np.gradient(self.deformation, rect/self.deformation.shape)`
     passes ONE array where np.gradient wants one spacing per axis (raises under numpy >= 1.13);
     ->                          `np.gradient(self.deformation, *(rect/self.deformation.shape))`
  2. Fourrier.get_normal (:124)  `dX, dY = self.DerivInterp(Point)` leaves two 1-element arrays, and
     `np.linalg.norm([dX, dY, 1])` then raises "inhomogeneous shape" under numpy >= 1.24;
     ->                          `dX, dY = self.DerivInterp(Point).flatten()`  (as MeasuredMap.get_normal :53-54 does)
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ART_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

PATCHES = [
    ("        self.DerivX, self.DerivY = np.gradient(self.deformation, rect/self.deformation.shape)\n",
     "        self.DerivX, self.DerivY = np.gradient(self.deformation, *(rect/self.deformation.shape))\n"),
    ("        dX, dY = self.DerivInterp(Point)\n",
     "        dX, dY = self.DerivInterp(Point).flatten()\n"),
]


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def build(verbose=True):
    src_pkg = os.path.join(SRC, "ART")
    if not os.path.isdir(src_pkg):
        raise RuntimeError(f"reference tree not found at {SRC}; oracle/_ref can only be built where it exists")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {"source": SRC, "files": {}, "numpy2_patch": [{"old": o.strip(), "new": n.strip()} for o, n in PATCHES]}
    try:
        manifest["source_commit"] = subprocess.run(["git", "-C", SRC, "rev-parse", "HEAD"], capture_output=True,
                                                   text=True, timeout=10).stdout.strip() or None
    except Exception:
        manifest["source_commit"] = None
    for variant in ("ART", os.path.join("ART_np2", "ART")):
        out = os.path.join(DST, variant)
        os.makedirs(out)
        for f in sorted(os.listdir(src_pkg)):
            if f.endswith(".py"):
                shutil.copyfile(os.path.join(src_pkg, f), os.path.join(out, f))
                if variant == "ART":
                    manifest["files"][f] = sha(os.path.join(out, f))
    lic = os.path.join(SRC, "LICENSE")
    if os.path.exists(lic):
        shutil.copyfile(lic, os.path.join(DST, "LICENSE"))
    p = os.path.join(DST, "ART_np2", "ART", "ModuleDefects.py")
    text = open(p).read()
    for old, new in PATCHES:
        if text.count(old) != 1:
            raise RuntimeError("numpy-2 compatibility patch does not apply: expected exactly one occurrence of " + repr(old))
        text = text.replace(old, new)
    open(p, "w").write(text)
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"oracle/_ref: {len(manifest['files'])} reference files copied unmodified from {src_pkg}; "
              f"ART_np2 carries the {len(PATCHES)}-line numpy>=2 patch of ModuleDefects.py")
    return DST


def available():
    return os.path.isdir(os.path.join(DST, "ART"))


if __name__ == "__main__":
    build()
    sys.exit(0)
