"""Host side of the remaining image ingest: LZW TIFF, PNG container, zip archives -- pinned on the reference's
own test files (tests/golden/format_fixtures.npz, made by make_format_fixtures.py) and on Python's zlib / zipfile."""
import io
import os
import struct
import zipfile
import zlib

import numpy as np
import pytest

from colormipsearch_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fmt():
    with np.load(os.path.join(ROOT, "tests", "golden", "format_fixtures.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def tiffs():
    with np.load(os.path.join(ROOT, "tests", "golden", "tiff_fixtures.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", ["lzw1", "lzw2"])
def test_lzw_tiff_decodes_like_imagej(fmt, name):
    """ImageArrayUtilsTest.readImageRangeForOtherCompression (:46-64): the LZW files (with the horizontal predictor) must give the
    pixels ImageJ's Opener gives (here: Pillow's, identical for this format)."""
    data = fmt["file_" + name].tobytes()
    px = fmt["pixels_" + name]
    info = capi.tiff_probe(data)
    assert info["compression"] == 5 and info["decodable"] == 0          # not decodable on the DEVICE
    got = capi.tiff_decode_rgb_host(data, px.shape[1], px.shape[0])
    assert np.array_equal(got, px)
    # rewritten as PackBits it is what the device ingest takes, pixel for pixel
    pb = capi.tiff_to_packbits(data)
    info2 = capi.tiff_probe(pb)
    assert info2["compression"] == 32773 and info2["decodable"] == 1 and info2["rows_per_strip"] == 8
    assert np.array_equal(capi.tiff_decode_rgb_host(pb, px.shape[1], px.shape[0]), px)


@pytest.mark.parametrize("name", ["pack1", "pack2", "stored1", "em_12191", "lm_GMR"])
def test_host_decoder_on_the_other_tiffs(tiffs, name):
    px = tiffs["pixels_" + name]
    assert np.array_equal(capi.tiff_decode_rgb_host(tiffs["file_" + name].tobytes(), px.shape[1], px.shape[0]), px)
    with pytest.raises(capi.CdsIllegalArgument):
        capi.tiff_decode_rgb_host(tiffs["file_" + name].tobytes(), px.shape[1] + 1, px.shape[0])


def _png_unfilter(raw, W, H, bps):
    """PNG specification section 9, in numpy / Python: the checker for the writer and for what the device must reproduce."""
    rb = W * bps
    out = np.zeros((H, rb), np.uint8)
    prev = np.zeros(rb, np.int32)
    for y in range(H):
        f = raw[y * (1 + rb)]
        x = np.frombuffer(raw[y * (1 + rb) + 1:(y + 1) * (1 + rb)], np.uint8).astype(np.int32)
        cur = np.zeros(rb, np.int32)
        if f == 0:
            cur = x
        elif f == 2:
            cur = (x + prev) & 255
        else:
            for i in range(rb):
                a = cur[i - bps] if i >= bps else 0
                b = prev[i]
                c = prev[i - bps] if i >= bps else 0
                if f == 1:
                    pred = a
                elif f == 3:
                    pred = (a + b) >> 1
                else:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (x[i] + pred) & 255
        out[y] = cur
        prev = cur
    return out


def _png_idat(data):
    pos, idat = 8, b""
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        if typ == b"IDAT":
            idat += data[pos + 8:pos + 8 + n]
        pos += 12 + n
    return idat


def test_png_probe_and_reference_files(fmt):
    for k in ("grad_BJD", "grad_VT016795", "grad_VT033614"):
        data = fmt["file_" + k].tobytes()
        info = capi.png_probe(data)
        assert (info["width"], info["height"], info["bit_depth"], info["color_type"], info["interlace"], info["decodable"]) == (1210, 566, 16, 0, 0, 1)
        # the files' own scanlines (zlib + filters undone in Python) are the fixture pixels: pins the checker used below
        raw = zlib.decompress(_png_idat(data))
        px = _png_unfilter(raw, 1210, 566, 2).reshape(566, 1210, 2)
        assert np.array_equal((px[:, :, 0].astype(np.uint16) << 8) | px[:, :, 1], fmt["pixels_" + k])
    for bad in (b"", b"\x89PNG\r\n\x1a\n", b"II*\0" + b"\0" * 40):
        with pytest.raises(capi.CdsError):
            capi.png_probe(bad)


@pytest.mark.parametrize("mode", [-1, 0, 1, 2, 3, 4])
def test_png_writer_round_trips(mode):
    rng = np.random.default_rng(mode + 7)
    W, H = 67, 23
    px = np.zeros((H, W), np.uint16)
    px[3:20, 5:60] = rng.integers(0, 700, (17, 55))
    px[10] = np.arange(W) * 900 % 65536
    data = capi.png_encode_gray16(px, mode)
    info = capi.png_probe(data)
    assert (info["width"], info["height"], info["bit_depth"], info["decodable"]) == (W, H, 16, 1)
    raw = zlib.decompress(_png_idat(data))
    if mode >= 0:
        assert all(raw[y * (1 + 2 * W)] == mode for y in range(H))
    got = _png_unfilter(raw, W, H, 2).reshape(H, W, 2)
    assert np.array_equal((got[:, :, 0].astype(np.uint16) << 8) | got[:, :, 1], px)


def test_zip_archives(tiffs):
    files = {"lib/a/em_12191.tif": tiffs["file_em_12191"].tobytes(), "lib/b/pack1.tif": tiffs["file_pack1"].tobytes(),
             "lib/b/em_LPLC2.tif": tiffs["file_em_LPLC2"].tobytes()}
    for method in (zipfile.ZIP_STORED, zipfile.ZIP_DEFLATED):
        bio = io.BytesIO()
        with zipfile.ZipFile(bio, "w", method) as zf:
            zf.writestr("lib/", b"")
            for k, v in files.items():
                zf.writestr(k, v)
        data = bio.getvalue()
        za = capi.ZipArchive(data)
        assert za.names() == ["lib/"] + list(files)
        for k, v in files.items():
            i = za.find(k)
            assert i >= 0 and za.read(i) == v
            span = za.stored_span(i)
            if method == zipfile.ZIP_STORED:
                assert data[span[0]:span[0] + span[1]] == v           # a stored entry IS the file, in place
            else:
                assert span is None
        # NeuronMIPUtils.openZipEntryStream :193-208: unknown path -> the first entry with the same file name
        assert za.names()[za.find("elsewhere/pack1.tif")] == "lib/b/pack1.tif"
        assert za.find("nothing.tif") == -1
        # a corrupted entry is caught by its CRC
        if method == zipfile.ZIP_STORED:
            bad = bytearray(data)
            span = za.stored_span(za.find("lib/b/pack1.tif"))
            bad[span[0] + 100] ^= 0xFF
            zb = capi.ZipArchive(bytes(bad))
            with pytest.raises(capi.CdsError):
                zb.read(zb.find("lib/b/pack1.tif"))
    with pytest.raises(capi.CdsError):
        capi.ZipArchive(b"not a zip archive at all, really not")
