"""Image ingest, CPU side: the oracle's TIFF / PackBits restatement against the reference's own fixtures,
and the host-only entry points of the C ABI (tag parsing, TIFF writer)."""
import os
import struct

import numpy as np
import pytest

from colormipsearch_b200 import capi
from oracle import tiff as OT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tiffs():
    with np.load(os.path.join(ROOT, "tests", "golden", "tiff_fixtures.npz")) as z:
        return {k: z[k] for k in z.files}


DECODABLE = ["pack1", "pack2", "stored1", "em_12191", "em_LPLC2", "lm_GMR"]


@pytest.mark.parametrize("name", DECODABLE)
def test_oracle_decodes_reference_fixtures(tiffs, name):
    """packBitsUncompress / readImageArrayRangeWithTiffReader restated == the pixels ImageJ (here: Pillow) reads."""
    got = OT.read_tiff_rgb(tiffs["file_" + name].tobytes())
    assert np.array_equal(got, tiffs["pixels_" + name])


def test_oracle_matches_cdsearch_fixtures(tiffs, fixtures):
    """The colour-depth MIPs decoded here are the images every scoring test of the reference uses."""
    assert np.array_equal(OT.read_tiff_rgb(tiffs["file_em_12191"].tobytes()), fixtures["em_12191"])
    assert np.array_equal(OT.read_tiff_rgb(tiffs["file_em_LPLC2"].tobytes()), fixtures["em_LPLC2"])


@pytest.mark.parametrize("name", ["pack1", "pack2"])
def test_oracle_range_read_like_reference_test(tiffs, name):
    """ImageArrayUtilsTest.readImageRangeForPackBits (:19-44): reading the range of the image's bounding rows gives the image's
    pixels inside those rows and black outside."""
    px = tiffs["pixels_" + name]
    Hh, Ww = px.shape[:2]
    nz = np.argwhere(px.any(axis=2))
    miny, maxy = nz[:, 0].min(), nz[:, 0].max() + 1
    minx, maxx = nz[:, 1].min(), nz[:, 1].max() + 1
    got = OT.read_tiff_rgb(tiffs["file_" + name].tobytes(), int(miny) * Ww, int(maxy) * Ww + int(maxx))
    assert np.array_equal(got[miny:maxy + 1], px[miny:maxy + 1])       # the test's `y <= boundaries[3]` rows
    assert not got[:miny].any()
    assert not got[maxy + 1:].any()


def test_probe_reads_the_tags(tiffs):
    info = capi.tiff_probe(tiffs["file_em_12191"].tobytes())
    assert (info["width"], info["height"], info["compression"], info["rows_per_strip"], info["n_strips"]) == (1210, 566, 32773, 8, 71)
    assert info["big_endian"] == 1 and info["decodable"] == 1 and info["samples_per_pixel"] == 3
    info = capi.tiff_probe(tiffs["file_pack1"].tobytes())
    assert (info["width"], info["height"], info["compression"], info["n_strips"], info["big_endian"]) == (256, 256, 32773, 1, 0)
    info = capi.tiff_probe(tiffs["file_lzw1"].tobytes())
    assert info["compression"] == 5 and info["decodable"] == 0
    info = capi.tiff_probe(tiffs["file_stored1"].tobytes())
    assert info["compression"] == 1 and info["decodable"] == 1 and info["data_bytes"] == 256 * 256 * 3
    # the oracle's tag reader agrees
    for name in DECODABLE + ["lzw1"]:
        a = capi.tiff_probe(tiffs["file_" + name].tobytes())
        b = OT.tiff_info(tiffs["file_" + name].tobytes())
        assert (a["width"], a["height"], a["compression"], a["n_strips"]) == (b["width"], b["height"], b["compression"], len(b["strip_offsets"]))
        assert a["data_bytes"] == sum(b["strip_lengths"])


def test_probe_rejects_what_is_not_a_tiff(tiffs):
    for bad in (b"", b"II", b"XX*\0\0\0\0\0", b"II+\0\x08\0\0\0\0\0\0\0", b"II*\0\xff\xff\xff\x7f"):
        with pytest.raises(capi.CdsError):
            capi.tiff_probe(bad)
    # a directory that runs past the end of the file
    data = bytearray(tiffs["file_pack1"].tobytes())
    (ifd,) = struct.unpack("<I", data[4:8])
    with pytest.raises(capi.CdsError):
        capi.tiff_probe(bytes(data[:ifd + 20]))


@pytest.mark.parametrize("rows_per_strip,compression", [(8, 32773), (1, 32773), (566, 32773), (7, 32773), (8, 1), (0, 1)])
def test_writer_round_trips_through_the_oracle(rows_per_strip, compression):
    rng = np.random.default_rng(rows_per_strip * 7 + compression)
    W, H = 301, 57
    img = np.zeros((H, W, 3), np.uint8)
    img[5:30, 40:200] = rng.integers(0, 256, (25, 160, 3))
    img[33:35, :] = 9                       # long runs, longer than 128 bytes
    img[40, 10:12] = 200                    # runs of two
    img[41, 0:130] = rng.integers(0, 2, (130, 1)) * 255   # short runs of a three-byte pattern
    data = capi.tiff_encode_rgb(img, rows_per_strip, compression)
    info = capi.tiff_probe(data)
    rps = rows_per_strip if 0 < rows_per_strip <= H else H
    assert info["decodable"] == 1 and info["n_strips"] == -(-H // rps) and info["compression"] == compression
    assert np.array_equal(OT.read_tiff_rgb(data), img)
    if compression == 32773:
        assert len(data) < img.size // 2


def test_writer_worst_case_fits_the_bound():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (33, 129, 3)).astype(np.uint8)       # incompressible
    data = capi.tiff_encode_rgb(img, 4, 32773)
    assert np.array_equal(OT.read_tiff_rgb(data), img)
    assert len(data) <= capi.lib().cds_tiff_encode_bound(129, 33, 4)


def test_probe_survives_corrupted_files(tiffs):
    """The tag parser reads files that come from outside: random corruption of valid files must end in a clean status (OK or an
    error), never in a crash, and whatever is reported as decodable must have its strips inside the file."""
    rng = np.random.default_rng(20241018)
    seeds = [tiffs["file_pack1"].tobytes(), tiffs["file_em_12191"].tobytes(), tiffs["file_stored1"].tobytes()[:4096] + b"\0" * 64,
             capi.tiff_encode_rgb(np.zeros((9, 17, 3), np.uint8), 2, 32773)]
    for it in range(3000):
        data = bytearray(seeds[it % len(seeds)])
        n = len(data)
        (ifd,) = struct.unpack("<I" if data[:2] == b"II" else ">I", data[4:8])
        for _ in range(int(rng.integers(1, 6))):
            # most flips inside the header / directory, where they matter
            pos = int(rng.integers(0, 8)) if rng.random() < 0.2 else int(min(n - 1, ifd + rng.integers(0, 200))) if rng.random() < 0.8 else int(rng.integers(0, n))
            data[pos] = int(rng.integers(0, 256))
        if rng.random() < 0.2:
            data = data[:int(rng.integers(0, n))]
        try:
            info = capi.tiff_probe(bytes(data))
        except capi.CdsError:
            continue
        assert info["width"] > 0 and info["height"] > 0
        if info["decodable"]:
            assert 0 <= info["data_bytes"] <= len(data) * max(1, info["n_strips"])


def _ifd_file(entries, extra=b""):
    """A little-endian TIFF whose IFD (at offset 8) holds `entries` = [(tag, type, count, value)], followed by `extra` bytes."""
    out = b"II" + struct.pack("<HI", 42, 8) + struct.pack("<H", len(entries))
    for tag, typ, count, value in entries:
        out += struct.pack("<HHII", tag, typ, count, value)
    return out + struct.pack("<I", 0) + extra


def test_probe_tiled_file_with_strip_tags_and_mismatched_strip_tables():
    """A tile tag next to StripOffsets without (or with fewer) StripByteCounts used to index the byte-count table out of bounds
    (a crash inside cds_tiff_probe, i.e. inside the JVM); strip tables of different lengths are a malformed file either way."""
    end = 8 + 2 + 4 * 12 + 4
    # width, height, 3 strip offsets (out of line), TileWidth -- no StripByteCounts at all
    tiled = _ifd_file([(256, 4, 1, 64), (257, 4, 1, 64), (273, 4, 3, end), (322, 4, 1, 16)], struct.pack("<III", 8, 8, 8))
    info = capi.tiff_probe(tiled)
    assert info["decodable"] == 0 and info["n_strips"] == 0
    # not tiled: 3 offsets, 2 byte counts
    end = 8 + 2 + 4 * 12 + 4
    bad = _ifd_file([(256, 4, 1, 64), (257, 4, 1, 64), (273, 4, 3, end), (279, 4, 2, end + 12)], struct.pack("<IIIII", 8, 8, 8, 1, 1))
    with pytest.raises(capi.CdsIllegalArgument):
        capi.tiff_probe(bad)
    # tiled with both tables present but of different lengths: still only reported as not decodable
    both = _ifd_file([(256, 4, 1, 64), (257, 4, 1, 64), (273, 4, 3, end + 12), (279, 4, 2, end + 24), (322, 4, 1, 16)],
                     struct.pack("<IIIIIIII", 0, 0, 0, 8, 8, 8, 1, 1))
    assert capi.tiff_probe(both)["decodable"] == 0
