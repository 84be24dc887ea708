"""Device side of the PNG ingest and the file-fed shape score: the reference's own gradient PNG files through
cds_png_decode_gray16 (zlib streams inflated on the device one warp per stream -- or by host threads, "device_inflate" 0 --, device
filter reconstruction + byte swap) and through cds_shape_score_pairs_files."""
import os
import struct
import zlib

import numpy as np
import pytest

from colormipsearch_b200 import capi
from oracle import oracle as O
from tests import golden_vectors as GV

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 1210, 566


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(n_dev=1)
    yield c
    c.close()


@pytest.fixture(scope="module")
def fmt():
    with np.load(os.path.join(ROOT, "tests", "golden", "format_fixtures.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(params=[1, 0, 2], ids=["device_inflate", "host_inflate", "device_inflate_with_forced_fallbacks"])
def inflate_mode(ctx, request):
    ctx.set_option("device_inflate", request.param)
    yield request.param
    ctx.set_option("device_inflate", 1)


def test_reference_gradient_pngs(ctx, fmt, inflate_mode):
    names = ["grad_BJD", "grad_VT016795", "grad_VT033614"]
    got = capi.png_decode_gray16(ctx, [fmt["file_" + k].tobytes() for k in names], W, H)
    for i, k in enumerate(names):
        assert np.array_equal(got[i], fmt["pixels_" + k]), k
    assert ctx.last_stats()["host_inflate_fallbacks"] == {1: 0, 0: 0, 2: 1}[inflate_mode]


def _png16(px, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, idat=1 << 30, wbits=15, tail=b""):
    """a 16-bit grayscale PNG, filter 0 on every row, with the zlib stream made the way the arguments say and cut into IDAT chunks of
    `idat` bytes; `tail` = bytes deflated after the image's own (a stream longer than the image)"""
    raw = b"".join(b"\0" + row.astype(">u2").tobytes() for row in px) + tail
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
    z = c.compress(raw) + c.flush()

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    body = b"".join(chunk(b"IDAT", z[i:i + idat]) for i in range(0, len(z), idat))
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", px.shape[1], px.shape[0], 16, 0, 0, 0, 0)) + body + chunk(b"IEND", b"")


def test_png_streams_of_every_kind(ctx, inflate_mode):
    """Stored, fixed and dynamic blocks, Huffman-only and run-length streams, small windows, streams cut into many IDAT chunks and
    streams that go on after the image: the same pixels from the device's inflate as from zlib; damaged streams are errors in both."""
    rng = np.random.default_rng(12)
    Ws, Hs = 333, 61
    px = np.zeros((Hs, Ws), np.uint16)
    px[4:55, 10:300] = (rng.integers(0, 40, (51, 290)) * rng.integers(0, 2, (51, 290))).astype(np.uint16)
    px[20:30, 50:250] = rng.integers(0, 65536, (10, 200))
    files, want = [], []
    for level in (0, 1, 6, 9):
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
            files.append(_png16(px, level, strategy))
    files.append(_png16(px, 6, idat=1))                    # one IDAT chunk per byte
    files.append(_png16(px, 6, idat=100))
    files.append(_png16(px, 6, wbits=9))                   # 512-byte window
    files.append(_png16(px, 6, tail=b"more than the image holds" * 40))
    files.append(_png16(np.zeros_like(px), 9))             # one long run
    want = [px] * (len(files) - 1) + [np.zeros_like(px)]
    got = capi.png_decode_gray16(ctx, files, Ws, Hs)
    for i in range(len(files)):
        assert np.array_equal(got[i], want[i]), i
    if inflate_mode == 1:
        assert ctx.last_stats()["host_inflate_fallbacks"] == 0
    # a stream cut short, and one with a flipped bit in the middle of its Huffman data: errors that name the file (zlib has the last word)
    good = _png16(px, 6)
    pos = good.index(b"IDAT") + 4
    n = struct.unpack(">I", good[pos - 8:pos - 4])[0]
    z = good[pos:pos + n]

    def rebuild(zz):
        def chunk(t, d):
            return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
        return good[:pos - 8] + chunk(b"IDAT", zz) + chunk(b"IEND", b"")
    for bad in (rebuild(z[:len(z) // 2]), rebuild(z[:2] + b"\x07" + z[3:])):
        with pytest.raises(capi.CdsError) as e:
            capi.png_decode_gray16(ctx, [good, bad, good], Ws, Hs)
        assert "file 1" in str(e.value)


def _png8(px):
    """an 8-bit grayscale PNG with a different filter on every row (Python zlib), for the widening path"""
    Hh, Ww = px.shape
    raw = bytearray()
    prev = np.zeros(Ww, np.int32)
    for y in range(Hh):
        cur = px[y].astype(np.int32)
        f = y % 5
        line = np.zeros(Ww, np.int32)
        for i in range(Ww):
            a = cur[i - 1] if i else 0
            b = prev[i]
            c = prev[i - 1] if i else 0
            if f == 0: pred = 0
            elif f == 1: pred = a
            elif f == 2: pred = b
            elif f == 3: pred = (a + b) >> 1
            else:
                p = a + b - c
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            line[i] = (cur[i] - pred) & 255
        raw += bytes([f]) + line.astype(np.uint8).tobytes()
        prev = cur

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", Ww, Hh, 8, 0, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(bytes(raw))) + chunk(b"IEND", b"")


def test_every_png_filter_and_bit_depth(ctx, inflate_mode):
    rng = np.random.default_rng(3)
    Ws, Hs = 301, 57
    px = np.zeros((Hs, Ws), np.uint16)
    px[5:50, 20:280] = rng.integers(0, 65536, (45, 260))
    px[30] = (np.arange(Ws) * 257) % 65536
    files = [capi.png_encode_gray16(px, m) for m in (-1, 0, 1, 2, 3, 4)]
    got = capi.png_decode_gray16(ctx, files, Ws, Hs)
    for i in range(len(files)):
        assert np.array_equal(got[i], px), i
    # 8-bit files are widened; more files than one staging chunk, odd width
    p8 = rng.integers(0, 256, (Hs, Ws)).astype(np.uint8)
    p8[:, 100:150] = 0
    f8 = _png8(p8)
    got = capi.png_decode_gray16(ctx, [f8, files[0]] * 40, Ws, Hs)
    assert np.array_equal(got[0], p8.astype(np.uint16)) and np.array_equal(got[78], p8.astype(np.uint16)) and np.array_equal(got[79], px)
    # errors name the file
    with pytest.raises(capi.CdsError) as e:
        capi.png_decode_gray16(ctx, [files[0], b"\x89PNG\r\n\x1a\n" + b"\0" * 40, files[1]], Ws, Hs)
    assert "file 1" in str(e.value)
    with pytest.raises(capi.CdsIllegalArgument) as e:
        capi.png_decode_gray16(ctx, [files[0]], Ws + 1, Hs)
    assert e.value.status == capi.CDS_ERR_SIZE_MISMATCH


def test_shape_scores_from_files_golden(ctx, fixtures, fmt, inflate_mode):
    """The reference's shape vectors (Shape2DMatchColorDepthSearchAlgorithmTest.java:86-132) with the inputs as FILES: targets as
    PackBits TIFF, gradients as the reference's own PNG files."""
    rects = O.label_rects(W, H)
    sms = capi.ShapeMaskSet(ctx, W, H, 20, True, rects)
    mask_keys = ["em_12191", "em_12191_FL"]
    sms.add_rgb(np.stack([fixtures[k] for k in mask_keys]))
    cases = [c for c in GV.SHAPE if c[3] is None]
    tiffs = [capi.tiff_encode_rgb(fixtures[c[1]], 8, 32773) for c in cases]
    pngs = [fmt["file_" + c[2]].tobytes() for c in cases]
    pm = [mask_keys.index(c[0]) for c in cases]
    gap, he, mir = sms.score_pairs_files(tiffs, pngs, None, pm, list(range(len(cases))))
    for i, c in enumerate(cases):
        assert (int(gap[i]), int(he[i]), bool(mir[i])) == (c[4], c[5], c[7]), c
    # more targets than a window, a target without pairs in the middle, has_variants, and the pixel call as the checker
    rng = np.random.default_rng(8)
    targets = capi.synth_rgb_host(1, 55, 0, 40, W, H)
    grads = ctx.synth_gradient(55, 0, 40, W, H, on_device=True)
    tf = [capi.tiff_encode_rgb(t, 8, 32773) for t in targets]
    pf = [capi.png_encode_gray16(g, -1 if i % 2 else 0) for i, g in enumerate(grads)]
    pm = rng.integers(0, 2, 150)
    pt = rng.integers(0, 40, 150)
    pt[pt == 17] = 18
    has = (np.arange(40) % 9 != 4).astype(np.uint8)
    a = sms.score_pairs(targets, grads, None, pm, pt, has)
    b = sms.score_pairs_files(tf, pf, None, pm, pt, has)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    # (the scored targets span two windows of 32; mode 2 treats every odd image of a window as refused by the device)
    if inflate_mode == 1:
        assert ctx.last_stats()["host_inflate_fallbacks"] == 0
    elif inflate_mode == 2:
        assert ctx.last_stats()["host_inflate_fallbacks"] > 0
    bad = list(pf)
    bad[21] = b"\x89PNG\r\n\x1a\n" + b"\0" * 60
    with pytest.raises(capi.CdsError) as e:
        sms.score_pairs_files(tf, bad, None, pm, pt, has)
    assert "file 21" in str(e.value)
    sms.close()
