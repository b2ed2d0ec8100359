"""The C++ drop-in shim (include/common/cregistration_b200.hpp) keeps the reference's mm_lls_icp
signature and builds + links against the C-ABI library without PCL/Eigen (stand-in type headers in
tests/stubs). On a machine without a GPU the call itself reports the missing device and returns 0."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_builds_links_and_runs():
    libdir = os.path.join(ROOT, "mulls_b200", "csrc")
    assert os.path.exists(os.path.join(libdir, "libmulls_b200.so")), "build the library first"
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "shim_caller")
        subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-I", os.path.join(ROOT, "tests", "stubs"),
                               "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "stubs", "shim_caller.cpp"),
                               "-o", exe, "-L", libdir, "-lmulls_b200", f"-Wl,-rpath,{libdir}"])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "shim compiled and linked" in out.stdout
