"""The drop-in boundary is a C ABI: compile a plain-C caller against include/arraylib_b200.h, link it to
libarraylib_b200.so and run it on the GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_program_drives_the_library(al, tmp_path):
    libdir = os.path.join(ROOT, "arraylib_b200", "lib")
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-O1", os.path.join(ROOT, "tests", "abi_smoke.c"),
                           "-I", os.path.join(ROOT, "include"), "-L", libdir, "-larraylib_b200", "-lm",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "C ABI smoke ok" in out.stdout and "ValueError: can not broadcast." in out.stdout
