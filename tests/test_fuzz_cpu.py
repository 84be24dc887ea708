"""Corrupted input files must end in a status, never in a memory error (include/cdsgpu.h: "nothing aborts across the ABI").

tests/fuzz_formats_main.cpp is compiled together with the library's host-side readers (csrc/cds_tiff.cpp, csrc/cds_formats.cpp)
under AddressSanitizer + UBSan and run on mutations of the reference's own test files (TIFF: PackBits / stored / LZW; 16-bit PNG
gradients; a zip archive of them).  Found by this harness and fixed: cds_zip_read copied `size` bytes of a stored entry after
checking only `compressed_size`; cds_tiff_to_packbits allocated whatever size the file's tags stated."""
import io
import os
import shutil
import subprocess
import zipfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "colormipsearch_b200", "csrc")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"      # cds_tiff.h names cudaStream_t in device-side prototypes


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("fuzz")
    exe = str(d / "fuzz_formats")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           "-I", CUDA_INC, os.path.join(ROOT, "tests", "fuzz_formats_main.cpp"), os.path.join(CSRC, "cds_formats.cpp"),
           os.path.join(CSRC, "cds_tiff.cpp"), "-o", exe, "-lz", "-lpthread"]
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr and "cannot find" in r.stderr:
        pytest.skip("libasan / libubsan are not installed")
    assert r.returncode == 0, r.stderr[-3000:]
    with np.load(os.path.join(ROOT, "tests", "golden", "tiff_fixtures.npz")) as t, \
            np.load(os.path.join(ROOT, "tests", "golden", "format_fixtures.npz")) as f:
        files = {"pack1.tif": t["file_pack1"].tobytes(), "stored1.tif": t["file_stored1"].tobytes(), "em_12191.tif": t["file_em_12191"].tobytes(),
                 "lzw1.tif": t["file_lzw1"].tobytes(), "lzw2.tif": f["file_lzw2"].tobytes(),
                 "grad_BJD.png": f["file_grad_BJD"].tobytes(), "grad_VT016795.png": f["file_grad_VT016795"].tobytes()}
    b = io.BytesIO()
    with zipfile.ZipFile(b, "w") as z:
        z.writestr(zipfile.ZipInfo("a/"), b"")
        z.writestr("a/b/img_1.tif", files["pack1.tif"], compress_type=zipfile.ZIP_STORED)
        z.writestr("a/img_2.tif", files["lzw1.tif"], compress_type=zipfile.ZIP_DEFLATED)
        z.writestr("c.png", files["grad_VT016795.png"], compress_type=zipfile.ZIP_DEFLATED)
    files["lib.zip"] = b.getvalue()
    paths = []
    for name, data in files.items():
        p = str(d / name)
        with open(p, "wb") as fh:
            fh.write(data)
        paths.append(p)
    return exe, paths


def _run(exe, iterations, seed, paths):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([exe, str(iterations), str(seed)] + paths, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-6000:])
    assert "0 failed expectations" in r.stdout
    return r.stdout


@pytest.mark.parametrize("seed", [1, 7])
def test_mutated_files_end_in_a_status(harness, seed):
    exe, paths = harness
    out = _run(exe, 2400, seed, paths)
    assert " tiff, " in out and " png, " in out and " zip " in out


def test_unmutated_seeds_are_read(harness):
    """0 iterations still parses the arguments; the seeds themselves are covered by test_formats_cpu.py / test_tiff_cpu.py"""
    exe, paths = harness
    _run(exe, 0, 1, paths)


def test_zip_reader_refuses_inconsistent_entries():
    """cds_zip_read with an entry the caller (or a corrupted directory) made inconsistent: a status, not a read past the archive."""
    from colormipsearch_b200 import capi
    payload = bytes(range(256)) * 4
    b = io.BytesIO()
    with zipfile.ZipFile(b, "w") as z:
        z.writestr("x.bin", payload, compress_type=zipfile.ZIP_STORED)
    za = capi.ZipArchive(b.getvalue())
    assert za.n == 1 and za.read(0) == payload
    # an entry record edited after indexing (a stored entry whose size is not its stored size) is refused by the read itself
    za.entries[0].size = len(payload) * 1000
    with pytest.raises(capi.CdsError):
        za.read(0)
    # the same inconsistency inside the archive's directory is refused when the directory is read
    data = bytearray(b.getvalue())
    cd = data.rfind(b"PK\x01\x02")
    data[cd + 24:cd + 28] = (len(payload) * 1000).to_bytes(4, "little")
    with pytest.raises(capi.CdsError):
        capi.ZipArchive(bytes(data))


def test_to_packbits_refuses_a_size_the_buffer_cannot_hold():
    """A size tag of 2^31 - 1 in a 13 kB file must not turn into an allocation of that many pixels."""
    import ctypes as C
    from colormipsearch_b200 import capi
    with np.load(os.path.join(ROOT, "tests", "golden", "tiff_fixtures.npz")) as t:
        data = bytearray(t["file_pack1"].tobytes())
    assert data[:2] == b"II"
    ifd = int.from_bytes(data[4:8], "little")
    n = int.from_bytes(data[ifd:ifd + 2], "little")
    for e in range(n):
        at = ifd + 2 + 12 * e
        if int.from_bytes(data[at:at + 2], "little") in (256, 257):      # ImageWidth, ImageLength
            data[at + 2:at + 4] = (4).to_bytes(2, "little")               # LONG
            data[at + 8:at + 12] = (0x7FFFFFFF).to_bytes(4, "little")
    buf = np.frombuffer(bytes(data), np.uint8)
    out = np.empty(1 << 20, np.uint8)
    got = C.c_int64(-1)
    st = capi.lib().cds_tiff_to_packbits(buf.ctypes.data_as(C.POINTER(C.c_uint8)), len(buf), out.ctypes.data_as(C.POINTER(C.c_uint8)), len(out), C.byref(got))
    assert st != 0
