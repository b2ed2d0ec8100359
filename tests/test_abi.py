"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/mulls_b200/abi.h declares, its PODs have the layout the ctypes mirror assumes, and the
product path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from mulls_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mulls_b200", "abi.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mulls_[a-z_0-9]+)\s*\(", src)) - {"mulls_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    names = declared_functions()
    assert set(names) == set(abi.EXPORTED_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in abi.h but not exported by libmulls_b200.so"


def test_struct_layout_matches_header():
    code = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "mulls_b200/abi.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(mulls_cloud_view), sizeof(mulls_icp_params), sizeof(mulls_icp_result),
             sizeof(mulls_icp_trace), sizeof(mulls_run_stats), sizeof(mulls_pca_out));
      printf("%zu %zu %zu %zu\n", offsetof(mulls_icp_params, target_bound), offsetof(mulls_icp_params, normal_bearing),
             offsetof(mulls_icp_result, sigma), offsetof(mulls_icp_trace, n_src));
      return 0; }"""
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(code)
        exe = os.path.join(td, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    sizes = [int(v) for v in out]
    assert sizes[:6] == [C.sizeof(abi.CloudView), C.sizeof(abi.IcpParams), C.sizeof(abi.IcpResult),
                         C.sizeof(abi.IcpTrace), C.sizeof(abi.RunStats), C.sizeof(abi.PcaOut)]
    assert sizes[6:] == [abi.IcpParams.target_bound.offset, abi.IcpParams.normal_bearing.offset,
                         abi.IcpResult.sigma.offset, abi.IcpTrace.n_src.offset]


def test_default_params_match_reference_defaults():
    lib = abi.load_library()
    p = abi.IcpParams()
    lib.mulls_icp_default_params(C.byref(p))
    q = abi.default_params()
    for name, _ in abi.IcpParams._fields_:
        a, b = getattr(p, name), getattr(q, name)
        if name == "target_bound":
            assert list(a) == list(b)
        else:
            assert a == b, name
    # cregistration.hpp:1115-1123
    assert p.max_iter_num == 20 and p.used_feature_type == b"111110" and p.weight_strategy == b"1101"
    assert abs(p.dis_thre_unit - 1.5) < 1e-7 and abs(p.normal_bearing - 45.0) < 1e-7


def test_product_path_has_no_cpu_fallback():
    """Without a CUDA device the library must refuse to create a context (and say why)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here; the no-device behaviour is checked on the CPU box")
    lib = abi.load_library()
    h = lib.mulls_create(0, 1, 1000, 1000)
    assert not h
    assert b"no CPU fallback" in lib.mulls_last_error(None)
    from mulls_b200.registration import Context

    with pytest.raises(RuntimeError):
        Context(0, 1, 1000, 1000)


def test_product_does_not_import_oracle():
    """No file of the product package may reference the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "mulls_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f == "__init__.py" and "oracle" not in text, (dirpath, f)
