"""The C ABI from C: the header compiles as strict C99 and a C program links libart_b200.so and calls the
GPU-free entry points (no compute is launched)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "attosecondraytracing_b200")
SRC = os.path.join(ROOT, "tests", "c_abi", "host_only.c")


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_is_plain_c_and_library_links_from_c(tmp_path):
    from attosecondraytracing_b200 import build
    build.build()
    exe = str(tmp_path / "host_only")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
           "-o", exe, "-L", LIBDIR, "-l:libart_b200.so", "-lm", "-Wl,-rpath," + LIBDIR]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert run.stdout.startswith("c-abi ok 200 40 88 184 64")
