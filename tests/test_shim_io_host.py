"""CPU: the C++ shim's file formats (image / volume loaders, result-table writers and loaders) exercised by a small
host-only C++ program -- no GPU call is made (only classes that do not derive from DIC/DVC are used)."""
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_tiff(path, vol, big_endian=False):
    """8-bit grayscale, uncompressed, one strip per page; little- or big-endian (the reference's stereo images are 'MM')."""
    e = ">" if big_endian else "<"
    nz, ny, nx = vol.shape
    out = bytearray((b"MM\x00*" if big_endian else b"II*\x00") + b"\x00\x00\x00\x00")
    prev = 4
    for z in range(nz):
        off = len(out)
        out += vol[z].astype(np.uint8).tobytes()
        if len(out) % 2:
            out += b"\x00"
        struct.pack_into(e + "I", out, prev, len(out))
        tags = [(256, 3, nx), (257, 3, ny), (258, 3, 8), (259, 3, 1), (262, 3, 1), (273, 4, off), (277, 3, 1), (278, 3, ny), (279, 4, nx * ny)]
        out += struct.pack(e + "H", len(tags))
        for tag, typ, val in tags:
            out += struct.pack(e + "HHI", tag, typ, 1) + (struct.pack(e + "HH", val, 0) if typ == 3 else struct.pack(e + "I", val))
        prev = len(out)
        out += struct.pack(e + "I", 0)
    with open(path, "wb") as f:
        f.write(out)


def test_shim_file_formats(tmp_path):
    r, c = np.meshgrid(np.arange(12), np.arange(16), indexing="ij")
    _write_tiff(tmp_path / "img.tif", ((7 * r + 3 * c) % 251)[None], big_endian=True)
    z, y, x = np.meshgrid(np.arange(4), np.arange(6), np.arange(8), indexing="ij")
    vol = (5 * z + 7 * y + 3 * x) % 251
    _write_tiff(tmp_path / "stack.tif", vol)
    with open(tmp_path / "vol.bin", "wb") as f:   # int32[3] header + float32 payload (src/oc_image.cpp:76-110)
        f.write(np.array([8, 6, 4], np.int32).tobytes())
        f.write((vol.astype(np.float32) + 0.5).tobytes())
    exe = str(tmp_path / "shim_io_test")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.join(ROOT, "opencorr_b200", "lib")
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-fopenmp", "-I", os.path.join(ROOT, "include", "opencorr"), "-o", exe,
                           os.path.join(ROOT, "tests", "native", "shim_io_test.cpp"), "-L", libdir, "-lopencorr_b200",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("ok")
