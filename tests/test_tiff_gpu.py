"""Image ingest on the device: TIFF strips decoded by tiff_decode_kernel == the oracle's restatement of the
reference's reader, and searches over TIFF files == searches over the decoded pixels."""
import os
import struct

import numpy as np
import pytest

from colormipsearch_b200 import capi
from oracle import oracle as O
from oracle import tiff as OT

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 1210, 566


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(n_dev=1)
    yield c
    c.close()


@pytest.fixture(scope="module")
def tiffs():
    with np.load(os.path.join(ROOT, "tests", "golden", "tiff_fixtures.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", ["pack1", "pack2", "stored1", "em_12191", "em_LPLC2", "lm_GMR"])
def test_device_decode_of_reference_fixtures(ctx, tiffs, name):
    px = tiffs["pixels_" + name]
    got = capi.tiff_decode_rgb(ctx, [tiffs["file_" + name].tobytes()], px.shape[1], px.shape[0])
    assert np.array_equal(got[0], px)
    assert np.array_equal(got[0], OT.read_tiff_rgb(tiffs["file_" + name].tobytes()))


def test_device_decode_batch_and_layouts(ctx):
    """Many files per call (more than one internal chunk), every strip layout the writer can produce, both compressions."""
    rng = np.random.default_rng(11)
    w, h = 333, 61
    files, imgs = [], []
    for i in range(150):
        img = np.zeros((h, w, 3), np.uint8)
        y0, x0 = rng.integers(0, h - 20), rng.integers(0, w - 90)
        img[y0:y0 + 20, x0:x0 + 90] = rng.integers(0, 256, (20, 90, 3))
        img[rng.integers(0, h)] = rng.integers(0, 256)
        imgs.append(img)
        files.append(capi.tiff_encode_rgb(img, [1, 3, 8, 61, 0][i % 5], 32773 if i % 7 else 1))
    got = capi.tiff_decode_rgb(ctx, files, w, h)
    assert np.array_equal(got, np.stack(imgs))
    for i in (0, 6, 7, 149):
        assert np.array_equal(got[i], OT.read_tiff_rgb(files[i]))
    assert capi.tiff_decode_rgb(ctx, [], w, h).shape == (0, h, w, 3)


def test_device_decode_errors(ctx, tiffs):
    with pytest.raises(capi.CdsError) as e:                       # LZW goes to ImageJ in the reference; not decodable here
        capi.tiff_decode_rgb(ctx, [tiffs["file_lzw1"].tobytes()], 256, 256)
    assert e.value.status == capi.CDS_ERR_UNSUPPORTED and "compression 5" in str(e.value)
    with pytest.raises(capi.CdsIllegalArgument) as e:             # wrong image size
        capi.tiff_decode_rgb(ctx, [tiffs["file_pack1"].tobytes()], 1210, 566)
    assert e.value.status == capi.CDS_ERR_SIZE_MISMATCH
    with pytest.raises(capi.CdsIllegalArgument):                  # not a TIFF
        capi.tiff_decode_rgb(ctx, [b"not a tiff at all"], 256, 256)
    good = tiffs["file_pack1"].tobytes()
    with pytest.raises(capi.CdsError) as e:                       # the second file is bad: the message names it
        capi.tiff_decode_rgb(ctx, [good, good[:100]], 256, 256)
    assert "file 1" in str(e.value)


def test_truncated_strip_leaves_zeros(ctx, tiffs):
    """A strip whose byte count stops early decodes as far as it goes; the rest of the image stays black, like the
    zero-initialised array of the reference (ImageArrayUtils.java:198) -- checked against the oracle's run decoder."""
    data = bytearray(tiffs["file_pack1"].tobytes())
    info = OT.tiff_info(bytes(data))
    (ifd,) = struct.unpack("<I", data[4:8])
    (n,) = struct.unpack("<H", data[ifd:ifd + 2])
    full = info["strip_lengths"][0]
    # cut at a run boundary so that the reference's decoder ends cleanly too
    strip = bytes(data[info["strip_offsets"][0]:info["strip_offsets"][0] + full])
    idx, cut = 0, 0
    while idx < full // 2:
        c = strip[idx]
        idx += (c + 2) if c < 128 else (2 if c != 128 else 1)
        cut = idx
    for i in range(n):
        at = ifd + 2 + 12 * i
        tag, typ, count = struct.unpack("<HHI", data[at:at + 8])
        if tag == 279:
            assert count == 1
            data[at + 8:at + 12] = struct.pack("<I", cut) if typ == 4 else struct.pack("<HH", cut, 0)
    exp = OT.read_tiff_rgb(bytes(data))
    got = capi.tiff_decode_rgb(ctx, [bytes(data)], 256, 256)[0]
    assert np.array_equal(got, exp)
    assert exp.any() and not exp[-1].any()


@pytest.fixture(scope="module")
def synth_files(ctx):
    masks = capi.synth_rgb_host(0, 0xC0FFEE, 0, 24, W, H)
    targets = capi.synth_rgb_host(1, 0xC0FFEE, 0, 70, W, H)
    files = [capi.tiff_encode_rgb(t, 8 if i % 3 else 566, 32773 if i % 5 else 1) for i, t in enumerate(targets)]
    return masks, targets, files


@pytest.mark.parametrize("chunk", [256, 16, 7])
def test_stream_search_over_tiff_files_equals_rgb(ctx, synth_files, chunk):
    masks, targets, files = synth_files
    rects = O.label_rects(W, H)
    ctx.set_option("stream_chunk", chunk)
    ctx.set_option("stream_chunk_tiff", chunk)
    try:
        for n_masks, k, pct in ((24, 300, 0.0), (24, 5, 1.0), (3, 9, 0.0)):
            ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
            ms.add_rgb(masks[:n_masks])
            exp = ms.search_stream(targets, k, pct)
            got = ms.search_stream_tiff(files, k, pct)
            st = ctx.last_stats()
            assert np.array_equal(got[3], exp[3])
            for m in range(n_masks):
                c = exp[3][m]
                for a, b in zip(got[:3], exp[:3]):
                    assert np.array_equal(a[m, :c], b[m, :c]), (chunk, n_masks, k, m)
            assert st["h2d_bytes"] < targets.nbytes                # the files crossed PCIe, not the pixels
            em = ms.search_stream_matches(targets, pct)
            gm = ms.search_stream_matches_tiff(files, pct)
            for a, b in zip(gm, em):
                assert np.array_equal(a, b)
            s, t, mir, cnt = ms.search_stream_tiff([], 4, 0.0)
            assert cnt.tolist() == [0] * n_masks
            ms.close()
    finally:
        ctx.set_option("stream_chunk", 256)
        ctx.set_option("stream_chunk_tiff", 4096)


def test_stream_search_two_kernel_ingest_equals_fused(ctx, synth_files):
    """cds_ctx_set_option("fused_ingest", 0): the streaming search decodes to RGB and encodes (its staging half is sized lazily) and returns
    the fused path's lists."""
    masks, targets, files = synth_files
    rects = O.label_rects(W, H)
    ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    ms.add_rgb(masks[:12])
    exp = ms.search_stream_tiff(files, 40, 0.0)
    try:
        ctx.set_option("fused_ingest", 0)
        for chunk in (4096, 16):
            ctx.set_option("stream_chunk_tiff", chunk)
            got = ms.search_stream_tiff(files, 40, 0.0)
            assert np.array_equal(got[3], exp[3])
            for m in range(12):
                c = exp[3][m]
                for a, b in zip(got[:3], exp[:3]):
                    assert np.array_equal(a[m, :c], b[m, :c]), (chunk, m)
    finally:
        ctx.set_option("fused_ingest", 1)
        ctx.set_option("stream_chunk_tiff", 4096)
    ms.close()


def test_stream_chunks_end_at_the_byte_cap(ctx, synth_files):
    """A chunk of the file search also ends where its files would exceed what the strip table addresses: forced here with a tiny cap."""
    masks, targets, files = synth_files
    rects = O.label_rects(W, H)
    ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    ms.add_rgb(masks[:12])
    exp = ms.search_stream_tiff(files, 50, 0.0)
    biggest = max(len(f) for f in files)
    try:
        for cap in (biggest + 1, 3 * biggest, 1):                    # one file per chunk at most / a few / a cap below every file (one file per chunk)
            ctx.set_option("stream_chunk_bytes", cap)
            got = ms.search_stream_tiff(files, 50, 0.0)
            assert np.array_equal(got[3], exp[3])
            for m in range(12):
                c = exp[3][m]
                for a, b in zip(got[:3], exp[:3]):
                    assert np.array_equal(a[m, :c], b[m, :c]), (cap, m)
    finally:
        ctx.set_option("stream_chunk_bytes", 0xC0000000)
    ms.close()


def test_stream_search_reports_the_bad_file(ctx, synth_files, tiffs):
    masks, targets, files = synth_files
    ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, O.label_rects(W, H))
    ms.add_rgb(masks[:3])
    bad = list(files[:20]) + [tiffs["file_pack1"].tobytes()] + list(files[20:30])
    with pytest.raises(capi.CdsIllegalArgument) as e:
        ms.search_stream_tiff(bad, 4, 0.0)
    assert "file 20" in str(e.value)
    # the context is still usable afterwards
    got = ms.search_stream_tiff(files[:10], 4, 0.0)
    exp = ms.search_stream(targets[:10], 4, 0.0)
    for a, b in zip(got, exp):
        assert np.array_equal(a, b)
    ms.close()


def test_library_from_tiff_files(ctx, synth_files):
    masks, targets, files = synth_files
    rects = O.label_rects(W, H)
    lib_rgb = capi.Library(ctx, W, H, 80)
    lib_rgb.add_rgb(targets)
    lib_tif = capi.Library(ctx, W, H, 80)
    assert lib_tif.add_tiff(files[:10]) == 0
    assert lib_tif.add_tiff(files[10:]) == 10                    # appends continue inside a block of 64
    assert len(lib_tif) == len(lib_rgb) == 70
    ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    ms.add_rgb(masks)
    a, am = ms.search_dense(lib_rgb)
    b, bm = ms.search_dense(lib_tif)
    assert np.array_equal(a, b) and np.array_equal(am, bm)
    ms.close(); lib_rgb.close(); lib_tif.close()


def test_masks_from_tiff_files(ctx, synth_files, tiffs, fixtures):
    """cds_maskset_add_tiff == cds_maskset_add_rgb of the decoded pixels; the reference's own EM mask file gives the mask size
    the reference's tests pin (PixelMatchColorDepthSearchAlgorithmTest: 12191_JRC2018U, threshold 20)."""
    masks, targets, files = synth_files
    rects = O.label_rects(W, H)
    mfiles = [capi.tiff_encode_rgb(m, 8, 32773) for m in masks] + [tiffs["file_em_12191"].tobytes()]
    mrgb = np.concatenate([masks, fixtures["em_12191"][None]])
    a = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    b = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    sa = a.add_rgb(mrgb)
    sb = b.add_tiff(mfiles[:5])
    sb = np.concatenate([sb, b.add_tiff(mfiles[5:])])
    assert np.array_equal(sa, sb)
    ra = a.search_stream(targets, 10, 0.0)
    rb = b.search_stream_tiff(files, 10, 0.0)
    for x, y in zip(ra, rb):
        assert np.array_equal(x, y)
    a.close(); b.close()


def _clamped_packbits(strip, out_len):
    """The decoder's documented behaviour for ANY input (include/cdsgpu.h): the reference's run loop, stopping at the end of the
    strip's input or of its rows; bytes a literal is missing, and everything that was not produced, are 0."""
    out = np.zeros(out_len, np.uint8)
    idx = pos = 0
    n = len(strip)
    while idx < n and pos < out_len:
        c = strip[idx]
        if c < 128:
            cnt = c + 1
            lit = strip[idx + 1:idx + 1 + cnt]
            lim = min(cnt, out_len - pos)
            out[pos:pos + min(lim, len(lit))] = np.frombuffer(bytes(lit[:lim]), np.uint8)
            idx += 1 + cnt
            pos += cnt
        elif c != 128:
            cnt = 257 - c
            v = strip[idx + 1] if idx + 1 < n else 0
            out[pos:min(out_len, pos + cnt)] = v
            idx += 2
            pos += cnt
        else:
            idx += 1
    return out


def test_device_decode_of_garbage_strips(ctx):
    """Random bytes inside a well-formed container: the decoder neither hangs nor writes outside the image, and produces what the
    clamped run loop produces."""
    rng = np.random.default_rng(99)
    w, h, rps = 67, 23, 5
    files, expect = [], []
    for i in range(40):
        img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
        data = bytearray(capi.tiff_encode_rgb(img, rps, 32773))
        info = OT.tiff_info(bytes(data))
        exp = np.zeros(h * w * 3, np.uint8)
        for s, (off, ln) in enumerate(zip(info["strip_offsets"], info["strip_lengths"])):
            garbage = rng.integers(0, 256, ln).astype(np.uint8)
            if i % 4 == 0:
                garbage[:] = rng.choice([0x81, 0x00, 0x80, 0x7f, 0xff], ln)      # long runs, no-ops, maximal literals
            data[off:off + ln] = garbage.tobytes()
            rows = min(rps, h - s * rps)
            exp[s * rps * w * 3:(s * rps + rows) * w * 3] = _clamped_packbits(bytes(garbage), rows * w * 3)
        files.append(bytes(data))
        expect.append(exp.reshape(h, w, 3))
    got = capi.tiff_decode_rgb(ctx, files, w, h)
    assert np.array_equal(got, np.stack(expect))


def test_failed_mask_append_leaves_the_mask_set_unchanged(ctx, synth_files):
    """cds_maskset_add_tiff with an undecodable file in its SECOND chunk of 64: the call fails naming the file, and the mask set
    is exactly what it was before the call (all or nothing) -- the next append and search behave like a fresh mask set's."""
    masks, targets, files = synth_files
    rects = O.label_rects(W, H)
    mfiles = [capi.tiff_encode_rgb(m, 8, 32773) for m in masks]
    many = [mfiles[i % len(mfiles)] for i in range(80)]
    bad = list(many)
    bad[70] = b"II*\0" + b"\0" * 60
    a = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    first = a.add_tiff(mfiles[:3])
    with pytest.raises(capi.CdsError) as e:
        a.add_tiff(bad)
    assert "file 70" in str(e.value)
    assert len(a) == 3 and np.array_equal(a.sizes(), first)
    more = a.add_tiff(many[:70])
    b = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    sb = np.concatenate([b.add_tiff(mfiles[:3]), b.add_tiff(many[:70])])
    assert np.array_equal(np.concatenate([first, more]), sb) and len(a) == len(b) == 73
    ra = a.search_stream_tiff(files, 10, 0.0)
    rb = b.search_stream_tiff(files, 10, 0.0)
    for x, y in zip(ra, rb):
        assert np.array_equal(x, y)
    a.close(); b.close()


def _expected_codes_and_valid(ctx, pixels, thr):
    """Code words of decoded pixels through the colour encoder (checked against the oracle's double predicate over all 2^24
    colours elsewhere), and the per-sector "can match" bits derived from them (cds_common.h: sector = SR / 32768 when the pixel is
    above the threshold and has a sector)."""
    n, Hh, Ww, _ = pixels.shape
    codes = ctx.debug_encode_colors(pixels.reshape(-1, 3), thr).reshape(n, Hh, Ww)
    vp = (((Ww + 31) // 32) + 3) // 4 * 4
    sr = (codes >> 8) & 0x3FFFF
    ok = ((codes & 0xC0000000) == 0) & (sr < 6 * 32768)
    valid = np.zeros((n, Hh, 6, vp), np.uint32)
    xs = np.arange(Ww)
    for s in range(6):
        bits = (ok & (sr // 32768 == s)).astype(np.uint32) << (xs % 32).astype(np.uint32)
        np.add.at(valid[:, :, s, :], (slice(None), slice(None), xs // 32), bits)       # distinct bits: add == or
    return codes, valid


def test_fused_ingest_equals_decode_then_encode(ctx, tiffs, synth_files):
    """tiff_encode_kernel (strips -> code words + valid bits, no RGB image in between) against decode + encode and against the
    colour encoder applied to the oracle-decoded pixels: the reference's own files, synthetic MIPs in several strip layouts
    (PackBits and stored), truncated and garbage strips."""
    # the reference's colour-depth MIP files (1210 x 566, 71 PackBits strips)
    names = ["em_12191", "em_LPLC2", "lm_GMR"]
    files = [tiffs["file_" + k].tobytes() for k in names]
    px = np.stack([tiffs["pixels_" + k] for k in names])
    for thr in (20, 100, 0):
        exp_c, exp_v = _expected_codes_and_valid(ctx, px, thr)
        for fused in (1, 0):
            c, v = ctx.debug_tiff_codes(files, W, H, thr, fused)
            assert np.array_equal(c, exp_c), (thr, fused)
            assert np.array_equal(v, exp_v), (thr, fused)
    # synthetic MIPs: rows per strip 8 / 566 / 1 / 7, PackBits and stored
    masks, targets, _ = synth_files
    imgs = np.concatenate([targets[:10], masks[:2]])
    layouts = [(8, 32773), (566, 32773), (1, 32773), (7, 32773), (8, 1), (566, 1), (3, 1)]
    sfiles = [capi.tiff_encode_rgb(im, *layouts[i % len(layouts)]) for i, im in enumerate(imgs)]
    exp_c, exp_v = _expected_codes_and_valid(ctx, imgs, 20)
    for fused in (1, 0):
        c, v = ctx.debug_tiff_codes(sfiles, W, H, 20, fused)
        assert np.array_equal(c, exp_c) and np.array_equal(v, exp_v), fused
    # small odd-sized images with noise (literal-heavy strips, runs that cross rows when the whole image is one strip)
    rng = np.random.default_rng(12)
    Ws, Hs = 333, 41
    small = np.zeros((4, Hs, Ws, 3), np.uint8)
    small[0] = rng.integers(0, 256, (Hs, Ws, 3))
    small[1, 5:30, 40:200] = rng.integers(0, 256, (25, 160, 3))
    small[2, :, ::7] = 77
    small[3, 11] = 255
    sf = [capi.tiff_encode_rgb(im, rps, comp) for im, (rps, comp) in zip(small, [(8, 32773), (Hs, 32773), (5, 1), (1, 32773)])]
    exp_c, exp_v = _expected_codes_and_valid(ctx, small, 20)
    for fused in (1, 0):
        c, v = ctx.debug_tiff_codes(sf, Ws, Hs, 20, fused)
        assert np.array_equal(c, exp_c) and np.array_equal(v, exp_v), fused


def test_fused_ingest_of_truncated_and_garbage_strips(ctx):
    """Malformed strips: both ingest paths follow the documented clamped PackBits rules (include/cdsgpu.h) and agree word for word."""
    rng = np.random.default_rng(99)
    Ws, Hs = 200, 24
    files, pixels = [], []
    for case in range(12):
        row = Ws * 3
        strips = []
        for s in range(3):                       # 3 strips of 8 rows
            n = int(rng.integers(1, 900))
            strips.append(bytes(rng.integers(0, 256, n).astype(np.uint8)))
        pixels.append(np.concatenate([_clamped_packbits(st, 8 * row) for st in strips]).reshape(Hs, Ws, 3))
        # a little-endian TIFF around the strips
        data = b"".join(strips)
        offs, o = [], 8
        for st in strips:
            offs.append(o); o += len(st)
        body = data + (b"\0" if len(data) & 1 else b"")
        p0 = 8 + len(body)
        extra = struct.pack("<HHH", 8, 8, 8) + struct.pack("<III", *offs) + struct.pack("<III", *[len(st) for st in strips])
        ifd = p0 + len(extra)
        ent = [(256, 4, 1, Ws), (257, 4, 1, Hs), (258, 3, 3, p0), (259, 3, 1, 32773), (262, 3, 1, 2), (273, 4, 3, p0 + 6), (277, 3, 1, 3),
               (278, 4, 1, 8), (279, 4, 3, p0 + 18), (284, 3, 1, 1)]
        d = b"II" + struct.pack("<HI", 42, ifd) + body + extra + struct.pack("<H", len(ent))
        for tag, typ, cnt, val in ent:
            d += struct.pack("<HHI", tag, typ, cnt) + (struct.pack("<HH", val, 0) if typ == 3 and cnt == 1 else struct.pack("<I", val))
        d += struct.pack("<I", 0)
        files.append(d)
    px = np.stack(pixels)
    got = capi.tiff_decode_rgb(ctx, files, Ws, Hs)
    assert np.array_equal(got, px)
    exp_c, exp_v = _expected_codes_and_valid(ctx, px, 20)
    for fused in (1, 0):
        c, v = ctx.debug_tiff_codes(files, Ws, Hs, 20, fused)
        assert np.array_equal(c, exp_c) and np.array_equal(v, exp_v), fused
