"""The library's SSBW reader (superslam_b200/csrc/weights.cpp) on valid, truncated and corrupted archives: a damaged
weights file must come back as SSB_ERR_IO with a message, never as a crash or an out-of-bounds read.  The reader sits
behind the device check of every ssb_*_create, so it is reached here through a small test-side shim built with nvcc
(tests/weights_shim.cu + the library's own weights.cpp / runtime.cu)."""
import ctypes as C
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, SP_WEIGHTS

CSRC = os.path.join(ROOT, "superslam_b200", "csrc")


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    so = str(tmp_path_factory.mktemp("shim") / "libweights_shim.so")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-shared", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-x", "cu", os.path.join(ROOT, "tests", "weights_shim.cu"),
           os.path.join(CSRC, "weights.cpp"), os.path.join(CSRC, "runtime.cu"), "-o", so]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    lib = C.CDLL(so)
    lib.shim_load_archive.restype = C.c_int
    lib.shim_load_archive.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_char_p, C.c_int]
    return lib


def load(lib, path):
    n, s, err = C.c_int(), C.c_double(), C.create_string_buffer(512)
    st = lib.shim_load_archive(str(path).encode(), C.byref(n), C.byref(s), err, 512)
    return st, n.value, s.value, err.value.decode()


def test_reads_what_the_python_writer_writes(shim, tmp_path):
    from superslam_b200.weights_io import load_archive, save_archive

    st, n, s, _ = load(shim, SP_WEIGHTS)
    ref = load_archive(SP_WEIGHTS)
    assert st == 0 and n == len(ref) == 24
    assert abs(s - sum(float(v.astype(np.float64).sum()) for v in ref.values())) < 1e-3
    a = {"a": np.arange(6, dtype=np.float32).reshape(2, 3), "empty": np.zeros((0, 4), np.float32), "scalar": np.float32([2.5])}
    p = tmp_path / "t.ssbw"
    save_archive(str(p), a)
    assert load(shim, p)[:3] == (0, 3, 17.5)


def test_missing_truncated_and_foreign_files(shim, tmp_path):
    assert load(shim, tmp_path / "nope.ssbw")[0] == 3                       # SSB_ERR_IO
    blob = open(SP_WEIGHTS, "rb").read()
    for cut in (0, 3, 11, 12, 40, 200, 1000, len(blob) // 2, len(blob) - 1):
        p = tmp_path / f"cut{cut}.ssbw"
        p.write_bytes(blob[:cut])
        st, _, _, err = load(shim, p)
        assert st == 3 and err, cut
    p = tmp_path / "foreign.bin"
    p.write_bytes(b"PK\x03\x04" + bytes(4096))
    assert load(shim, p)[0] == 3
    p.write_bytes(b"SSBW" + struct.pack("<II", 2, 1) + bytes(64))           # unknown version
    assert load(shim, p)[0] == 3


def entry(name, dims, off, size, dtype=0):
    nb = name.encode()
    return struct.pack("<I", len(nb)) + nb + struct.pack("<II", dtype, len(dims)) + struct.pack(f"<{len(dims)}I", *dims) + \
        struct.pack("<QQ", off, size)


@pytest.mark.parametrize("case", ["offset_wraps", "dims_overflow", "size_mismatch", "bad_dtype", "too_many_dims",
                                  "name_past_end", "count_lies", "offset_past_end"])
def test_corrupted_headers_are_rejected_without_reading_out_of_bounds(shim, tmp_path, case):
    head = b"SSBW" + struct.pack("<II", 1, 1)
    body = {
        "offset_wraps": entry("w", [2], 2 ** 64 - 4, 8),                    # off + size wraps around to 4
        "dims_overflow": entry("w", [2 ** 31, 2 ** 31, 4], 64, 0),          # element count wraps to 0 == size / 4
        "size_mismatch": entry("w", [4], 64, 12),
        "bad_dtype": entry("w", [4], 64, 16, dtype=7),
        "too_many_dims": struct.pack("<I", 1) + b"w" + struct.pack("<II", 0, 9) + bytes(9 * 4 + 16),
        "name_past_end": struct.pack("<I", 10 ** 6) + b"w",
        "count_lies": b"",                                                   # header promises one tensor, file ends
        "offset_past_end": entry("w", [4], 10 ** 6, 16),
    }[case]
    p = tmp_path / f"{case}.ssbw"
    p.write_bytes(head + body + bytes(256))
    st, n, _, err = load(shim, p)
    assert st == 3 and n == 0 and err, (case, st, err)
