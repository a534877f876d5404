"""PCD reader on the boundary: the formats pcl::io::loadPCDFile accepts for the reference's
clouds (Dialog/PCLViewer.cpp:80-89) — ASCII and binary, x y z first or not, extra fields skipped."""
import os
import struct

import numpy as np
import pytest

from conftest import ROOT
from dialog_b200 import pcd


def _write_ascii(path, xyz, extra=True):
    with open(path, "w") as f:
        f.write("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n")
        f.write("FIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\n" if extra else
                "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n")
        f.write(f"WIDTH {len(xyz)}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {len(xyz)}\nDATA ascii\n")
        for p in xyz:
            f.write(f"{p[0]:.9g} {p[1]:.9g} {p[2]:.9g}" + (" 4294901760\n" if extra else "\n"))


def _write_binary(path, xyz):
    n = len(xyz)
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS rgb x y z\nSIZE 4 4 4 4\nTYPE U F F F\n"
           f"COUNT 1 1 1 1\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary\n")
    with open(path, "wb") as f:
        f.write(hdr.encode())
        for p in xyz:
            f.write(struct.pack("<Ifff", 0xFF00FF00, *[float(v) for v in p]))


def test_ascii_and_binary_roundtrip(tmp_path):
    rng = np.random.default_rng(5)
    xyz = rng.normal(size=(257, 3)).astype(np.float32)
    xyz[7] = [np.nan, 1, 2]
    for writer, name in ((_write_ascii, "a.pcd"), (_write_binary, "b.pcd")):
        path = tmp_path / name
        writer(path, xyz)
        got = pcd.read_pcd_xyz(str(path))
        assert got.shape == (257, 4) and got.dtype == np.float32
        assert np.array_equal(got[:, :3], xyz, equal_nan=True) and (got[:, 3] == 1).all()
    _write_ascii(tmp_path / "c.pcd", xyz[:10], extra=False)
    assert np.array_equal(pcd.read_pcd_xyz(str(tmp_path / "c.pcd"))[:, :3], xyz[:10], equal_nan=True)


def _lzf_literal_only(data: bytes) -> bytes:
    """A valid LZF stream made of literal runs only (control byte < 32 = run length - 1)."""
    out = bytearray()
    for i in range(0, len(data), 32):
        chunk = data[i:i + 32]
        out.append(len(chunk) - 1)
        out += chunk
    return bytes(out)


def test_binary_compressed(tmp_path):
    """DATA binary_compressed as pcl::io::savePCDFileBinaryCompressed writes it: two uint32 sizes, an LZF stream, the
    payload field-major (all x, then all y, ...), here with a double-precision z and a back reference in the stream."""
    rng = np.random.default_rng(6)
    n = 300
    xyz = rng.normal(size=(n, 3)).astype(np.float32)
    rgb = np.full(n, 0xFF102030, np.uint32)
    payload = rgb.tobytes() + xyz[:, 0].tobytes() + xyz[:, 1].tobytes() + xyz[:, 2].astype("<f8").tobytes()
    comp = bytearray(_lzf_literal_only(payload[:64]))
    # one back reference: copy 8 bytes from 64 bytes back (ctrl = (len - 2) << 5 | (off - 1) >> 8, then (off - 1) & 255)
    payload = payload[:64] + payload[0:8] + payload[72:]
    comp += bytes([((8 - 2) << 5) | 0, 63])
    comp += _lzf_literal_only(payload[72:])
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS rgb x y z\nSIZE 4 4 4 8\nTYPE U F F F\n"
           f"COUNT 1 1 1 1\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary_compressed\n")
    path = tmp_path / "c.pcd"
    with open(path, "wb") as f:
        f.write(hdr.encode())
        f.write(struct.pack("<II", len(comp), len(payload)))
        f.write(bytes(comp))
    want = np.frombuffer(payload, np.uint8)
    exp = np.ones((n, 4), np.float32)
    exp[:, 0] = np.frombuffer(want[4 * n: 8 * n].tobytes(), "<f4")
    exp[:, 1] = np.frombuffer(want[8 * n: 12 * n].tobytes(), "<f4")
    exp[:, 2] = np.frombuffer(want[12 * n: 20 * n].tobytes(), "<f8").astype(np.float32)
    got = pcd.read_pcd_xyz(str(path))
    assert np.array_equal(got, exp)
    bad = tmp_path / "bad.pcd"
    with open(bad, "wb") as f:
        f.write(hdr.encode())
        f.write(struct.pack("<II", len(comp) - 5, len(payload)))
        f.write(bytes(comp[:-5]))
    with pytest.raises(ValueError):
        pcd.read_pcd_xyz(str(bad))


def test_reads_the_reference_fixture_format(double_shadow, tmp_path):
    # same header as Dialog/double_shadow.pcd (FIELDS x y z rgb, TYPE F F F U, DATA ascii)
    _write_ascii(tmp_path / "ds.pcd", double_shadow)
    got = pcd.read_pcd_xyz(str(tmp_path / "ds.pcd"))
    assert np.array_equal(got[:, :3], double_shadow)


def test_rejects_unsupported(tmp_path):
    p = tmp_path / "bad.pcd"
    p.write_text("VERSION 0.7\nFIELDS a b\nSIZE 4 4\nTYPE F F\nCOUNT 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n1 2\n")
    with pytest.raises(ValueError):
        pcd.read_pcd_xyz(str(p))
