#!/usr/bin/env python
"""Builds tuning variants of libart_b200.so into build/variants/ (git-ignored; they travel to the GPU box with
gpurun) for tools/tune_variants.py:   python tools/build_variants.py name=-DFLAG,-DFLAG ...
Without arguments: the block-shape study of round 2 (threads per block of the trace kernel classes)."""
import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from attosecondraytracing_b200 import build as b  # noqa: E402

DEFAULT = {
    "bt192": ["-DART_BT_QUADRIC=192", "-DART_BT_TOROID=192", "-DART_BT_ANY=192"],
    "bt224": ["-DART_BT_QUADRIC=224", "-DART_BT_TOROID=224", "-DART_BT_ANY=224"],
    "bt320q": ["-DART_BT_QUADRIC=320"],
    "bt128x4": ["-DART_BT_QUADRIC=128", "-DART_BT_TOROID=128", "-DART_BT_ANY=128", "-DART_MINB=4", "-DART_GRID_PER_SM=4"],
    "def224": ["-DART_BT_DEF=224"],
}


def one(item):
    name, flags = item
    out = os.path.join(ROOT, "build", "variants", name + ".so")
    cmd = [b.nvcc()] + b.NVCC_FLAGS + flags + ["-o", out] + b.SOURCES
    res = subprocess.run(cmd, cwd=b.CSRC, capture_output=True, text=True)
    return name, res.returncode, res.stderr[-800:]


def main():
    todo = dict(DEFAULT)
    if len(sys.argv) > 1:
        todo = {a.split("=", 1)[0]: [f for f in a.split("=", 1)[1].split(",") if f] for a in sys.argv[1:]}
    os.makedirs(os.path.join(ROOT, "build", "variants"), exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
        for name, rc, err in ex.map(one, todo.items()):
            print(name, "ok" if rc == 0 else "FAILED\n" + err, flush=True)


if __name__ == "__main__":
    main()
