"""CPU-side checks of the boundary: the shared library loads, exports every symbol the header declares,
and fails loudly (status + message, no fallback) when there is no sm_100 device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, SP_WEIGHTS

HEADER = os.path.join(ROOT, "include", "superslam_b200.h")


def _header_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from superslam_b200 import _lib

    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 25
    bound = {name for name, _, _ in _lib.SYMBOLS}
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in bound, f"{name} declared in the header but not bound in _lib.SYMBOLS"


def test_no_gpu_fails_loudly_without_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from superslam_b200 import _lib, frontend

    assert _lib.load().ssb_device_count() < 0
    with pytest.raises(_lib.SsbError):
        frontend.SuperPoint(SP_WEIGHTS, 64)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "superslam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_only_the_allowed_places_touch_the_oracle():
    """oracle/ is test infrastructure: outside tests/ and oracle/ itself only bench.py (its CPU legs) and
    __graft_entry__.py (smoke) may import it; include/ and tools/ never mention it in code."""
    users = set()
    for dirpath, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in (".git", "gpurun_out", "build", "__pycache__", "tests", "oracle", "baseline")]
        for f in files:
            if f.endswith((".py", ".sh")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    users.add(os.path.relpath(os.path.join(dirpath, f), ROOT))
    assert users == {"bench.py", "__graft_entry__.py"}, users
    bench = open(os.path.join(ROOT, "bench.py")).read()
    # bench.py: every oracle import sits inside oracle_pair_seconds (the cpu_baseline / --impl reference legs)
    body = bench[bench.index("def oracle_pair_seconds"):bench.index("def ", bench.index("def oracle_pair_seconds") + 10)]
    assert bench.count("from oracle") == body.count("from oracle") > 0
    for header in os.listdir(os.path.join(ROOT, "include")):
        assert "oracle/" not in re.sub(r"//.*|/\*.*?\*/", "", open(os.path.join(ROOT, "include", header)).read(), flags=re.S)


def test_weight_archive_roundtrip(tmp_path):
    import numpy as np

    from superslam_b200.weights_io import load_archive, save_archive

    a = {"x.weight": np.arange(24, dtype=np.float32).reshape(2, 3, 4), "y": np.array([1.5], np.float32)}
    p = str(tmp_path / "t.ssbw")
    save_archive(p, a)
    b = load_archive(p)
    assert list(b) == list(a) and all(np.array_equal(a[k], b[k]) for k in a)
    sp = load_archive(SP_WEIGHTS)
    assert sp["conv1a.weight"].shape == (64, 1, 3, 3) and sp["convPb.weight"].shape == (65, 256, 1, 1)
    assert sum(v.size for v in sp.values()) == 1300865


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 on its own (no C++ types, no torch types)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-x", "c", HEADER],
                         capture_output=True, text=True)
    assert res.returncode == 0 and res.stderr.strip() == "", res.stderr


def test_committed_superpoint_archive_is_the_converted_reference_checkpoint(tmp_path):
    """Provenance of superslam_b200/weights/superpoint_v1.ssbw: tools/convert_superpoint_weights.py on the reference's
    own weights/superpoint_v1.pth gives the committed file, byte for byte."""
    import subprocess
    import sys

    src = "/root/reference/weights/superpoint_v1.pth"
    if not os.path.exists(src):
        pytest.skip("reference checkpoint not on this machine")
    dst = tmp_path / "sp.ssbw"
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "convert_superpoint_weights.py"), src, str(dst)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    assert open(dst, "rb").read() == open(SP_WEIGHTS, "rb").read()


def test_eigenplaces_converter_keeps_names_and_drops_bn_counters(tmp_path):
    import subprocess
    import sys

    import torch

    from superslam_b200.eigenplaces_weights import make_random_weights, save_state_dict
    from superslam_b200.weights_io import load_archive

    sd = make_random_weights(5)
    ckpt = {}
    for k, v in sd.items():                       # as torch saves it: DataParallel prefix, BN step counters
        ckpt["module." + k] = v
        if k.endswith("running_var"):
            ckpt["module." + k.replace("running_var", "num_batches_tracked")] = torch.tensor(1234)
    want, src, dst = tmp_path / "want.ssbw", tmp_path / "ep.pth", tmp_path / "ep.ssbw"
    save_state_dict(sd, str(want))
    torch.save(ckpt, src)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "convert_eigenplaces_weights.py"), str(src), str(dst)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    a, b = load_archive(str(want)), load_archive(str(dst))
    assert list(a) == list(b) and all(np.array_equal(a[k], b[k]) for k in a)
    assert not any("num_batches_tracked" in k for k in b) and b["aggregation.1.p"].shape == (1,)
