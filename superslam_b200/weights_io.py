"""Flat tensor-archive format ("SSBW") shared by the Python host code and the C++ runtime.

Layout (little endian):
    char[4]  magic = "SSBW"
    u32      version = 1
    u32      n_tensors
    repeat n_tensors:
        u32  name_len, char[name_len] name (no NUL)
        u32  dtype  (0 = f32)
        u32  ndim, u32[ndim] dims
        u64  byte_offset (from start of file, 64-byte aligned), u64 byte_size
    raw data blobs

The C++ reader is superslam_b200/csrc/weights.cpp (same layout).  SuperPoint tensors keep the key
names of the reference checkpoint weights/superpoint_v1.pth (conv1a.weight ... convDb.bias, see
/root/reference/utils/convert_superpoint_to_onnx.py:38-49); LightGlue tensors keep cvg/LightGlue's
state-dict names (transformers.{i}.self_attn.Wqkv.weight ...).
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

MAGIC = b"SSBW"


def save_archive(path: str, tensors: "OrderedDict[str, np.ndarray]") -> None:
    names = list(tensors.keys())
    arrs = [np.ascontiguousarray(tensors[k], dtype=np.float32) for k in names]
    header_size = 12
    for n, a in zip(names, arrs):
        header_size += 4 + len(n.encode()) + 4 + 4 + 4 * a.ndim + 16
    off = (header_size + 63) // 64 * 64
    entries = []
    for a in arrs:
        entries.append((off, a.nbytes))
        off = (off + a.nbytes + 63) // 64 * 64
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", 1, len(names)))
        for n, a, (o, s) in zip(names, arrs, entries):
            nb = n.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<II", 0, a.ndim))
            f.write(struct.pack("<%dI" % a.ndim, *a.shape))
            f.write(struct.pack("<QQ", o, s))
        for a, (o, s) in zip(arrs, entries):
            f.seek(o)
            f.write(a.tobytes())
        # pad the tail so the file size is a multiple of 64
        end = (f.tell() + 63) // 64 * 64
        if end > f.tell():
            f.seek(end - 1)
            f.write(b"\0")


def load_archive(path: str) -> "OrderedDict[str, np.ndarray]":
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:4] != MAGIC:
        raise ValueError(f"{path}: not an SSBW archive")
    ver, n = struct.unpack_from("<II", buf, 4)
    if ver != 1:
        raise ValueError(f"{path}: unsupported SSBW version {ver}")
    p = 12
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", buf, p)
        p += 4
        name = buf[p : p + ln].decode()
        p += ln
        dtype, ndim = struct.unpack_from("<II", buf, p)
        p += 8
        dims = struct.unpack_from("<%dI" % ndim, buf, p)
        p += 4 * ndim
        off, size = struct.unpack_from("<QQ", buf, p)
        p += 16
        if dtype != 0:
            raise ValueError("only f32 tensors are supported")
        out[name] = np.frombuffer(buf, dtype=np.float32, count=size // 4, offset=off).reshape(dims).copy()
    return out
