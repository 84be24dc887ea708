"""The reference's own JUnit scenarios replayed through the C++ mirror of its API (colormipsearch_b200/host/cds_host.hpp over the
C ABI): tests/host_mirror_main.cpp is compiled with g++, linked against libcdsgpu.so and run on the fixture images."""
import os
import subprocess

import numpy as np
import pytest

from colormipsearch_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    exe = str(tmp_path / "host_mirror")
    so = B.build()
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "host_mirror_main.cpp"), "-o", exe,
           "-L" + os.path.dirname(so), "-lcdsgpu", "-Wl,-rpath," + os.path.dirname(so),
           "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_host_mirror_compiles_and_links(tmp_path):
    """CPU-only check: the header-only mirror compiles against include/cdsgpu.h and links against the built library."""
    assert os.path.exists(_compile(tmp_path))


@pytest.mark.gpu
def test_reference_scenarios_through_the_cpp_mirror(tmp_path, fixtures):
    exe = _compile(tmp_path)
    H, W = fixtures["em_12191"].shape[:2]
    (tmp_path / "dims.txt").write_text("%d %d\n" % (W, H))
    for k in ("em_12191", "em_12191_FL", "lm_VT033614", "lm_BJD", "lm_VT016795", "zgap_BJD"):
        np.ascontiguousarray(fixtures[k], np.uint8).tofile(tmp_path / (k + ".rgb"))
    np.ascontiguousarray(fixtures["grad_BJD"], np.uint16).tofile(tmp_path / "grad_BJD.g16")
    # the same images as TIFF files: the reference's own PackBits file for the EM mask, our writer (PackBits / stored) for the rest
    from colormipsearch_b200 import capi
    with np.load(os.path.join(ROOT, "tests", "golden", "tiff_fixtures.npz")) as z:
        (tmp_path / "em_12191.tif").write_bytes(z["file_em_12191"].tobytes())
    for k, (rps, comp) in {"em_12191_FL": (566, 1), "lm_VT033614": (8, 32773), "lm_BJD": (1, 32773), "lm_VT016795": (566, 32773)}.items():
        (tmp_path / (k + ".tif")).write_bytes(capi.tiff_encode_rgb(fixtures[k], rps, comp))
    out = subprocess.run([exe, str(tmp_path)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    lines = {" ".join(ln.split()[:2]): ln.split()[2:] for ln in out}
    # PixelMatchColorDepthSearchAlgorithmTest.java:72-103
    exp = {"12191xVT033614": (439, 0, 10299), "12191xBJD": (414, 0, 10299), "FLxVT033614": (515, 0, 17340),
           "FLxVT016795": (483, 0, 17340), "12191xVT016795": (426, 1, 10299)}
    for name, (score, mir, size) in exp.items():
        got = lines["pixel " + name]
        assert (int(got[0]), int(got[1]), int(got[3])) == (score, mir, size), name
        assert int(got[5]) == 1                                   # isMatch at pctPositivePixels 1
        assert abs(float(got[7]) - np.float32(score / size)) < 1e-6
    assert lines["odd_xyshift IllegalArgumentException"][:2] == ["XY", "shift"]
    assert "size_mismatch IllegalArgumentException" in lines
    assert lines["empty_mask 0"][:3] == ["0", "querySize", "0"]
    # batched seam: per mask, descending matchingPixels
    batched = [ln.split() for ln in out if ln.startswith("batched")]
    got = [(int(b[2]), int(b[4]), int(b[6]), int(b[8])) for b in batched]
    assert got[:3] == [(0, 0, 439, 0), (0, 2, 426, 1), (0, 1, 414, 0)]
    assert (1, 0, 515, 0) in got and (1, 2, 483, 0) in got
    every = [ln.split() for ln in out if ln.startswith("allpairs")]
    assert [(int(b[2]), int(b[4]), int(b[6]), int(b[8])) for b in every] == got       # 3 targets, K = 3: the same six pairs
    tiffp = [ln.split() for ln in out if ln.startswith("tiffpairs")]
    assert [(int(b[2]), int(b[4]), int(b[6]), int(b[8])) for b in tiffp] == got       # TIFF files in, the same pairs out
    tiffb = [ln.split() for ln in out if ln.startswith("tiffbest")]
    assert [(int(b[2]), int(b[4]), int(b[6]), int(b[8])) for b in tiffb] == [(0, 0, 439, 0)]
    # Shape2DMatchColorDepthSearchAlgorithmTest.java:53-54, 230-291
    assert lines["shape_masks 17340"][:1] == ["70640"] and lines["shape_masks 17340"][2] == "2"
    assert lines["shape 12191xBJD_zgapfile"] == ["33884", "523", "34058", "0"]
    assert lines["shape_missing -1"] == ["-1", "-1"]
    assert out[-1].startswith("normalized") and abs(float(out[-1].split()[1]) - 46833.58) < 0.1
