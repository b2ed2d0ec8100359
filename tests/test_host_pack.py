"""Host-side wire format (csrc/host_pack.h, mulls_pack_rows): the packed layouts must carry exactly the bits of the
48-byte rows (utility.hpp:40) that the ingest kernel reads. CPU only — no compute call on the GPU."""
import ctypes as C

import numpy as np
import pytest

from mulls_b200 import abi


def _aligned(n_floats):
    raw = np.zeros(n_floats + 8, dtype=np.float32)
    off = (-raw.ctypes.data % 16) // 4
    return raw[off:off + n_floats]


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 127, 16384, 16385, 40001])
@pytest.mark.parametrize("fmt", [1, 2])
def test_pack_rows_bits(n, fmt):
    lib = abi.load_library()
    rng = np.random.default_rng(n * 7 + fmt)
    rows = rng.standard_normal((n, 12)).astype(np.float32)
    # arbitrary bit patterns (NaN payloads, denormals) must travel unchanged
    if n:
        rows.view(np.uint32)[::3, 5] = 0x7FC12345
        rows.view(np.uint32)[::5, 2] = 0x00000001
    out = _aligned(7 * n if fmt == 1 else 8 * n)
    fp = C.POINTER(C.c_float)
    rc = lib.mulls_pack_rows(rows.ctypes.data_as(fp), n, fmt, out.ctypes.data_as(fp))
    assert rc == 0
    pos = out[: 4 * n].reshape(n, 4).view(np.uint32)
    r = rows.view(np.uint32)
    assert np.array_equal(pos[:, :3], r[:, 0:3]) and np.array_equal(pos[:, 3], r[:, 8])
    if fmt == 1:
        nrm = out[4 * n:].reshape(n, 3).view(np.uint32)
        assert np.array_equal(nrm, r[:, 4:7])
    else:
        nrm = out[4 * n:].reshape(n, 4).view(np.uint32)
        assert np.array_equal(nrm[:, :3], r[:, 4:7]) and np.array_equal(nrm[:, 3], r[:, 9])


def test_pack_rows_rejects_bad_arguments():
    lib = abi.load_library()
    fp = C.POINTER(C.c_float)
    rows = np.zeros((4, 12), dtype=np.float32)
    out = _aligned(64)
    assert lib.mulls_pack_rows(rows.ctypes.data_as(fp), 4, 0, out.ctypes.data_as(fp)) == -101
    assert lib.mulls_pack_rows(rows.ctypes.data_as(fp), 4, 3, out.ctypes.data_as(fp)) == -101
    misaligned = out[1:]
    assert lib.mulls_pack_rows(rows.ctypes.data_as(fp), 4, 1, misaligned.ctypes.data_as(fp)) == -101
