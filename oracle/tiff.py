"""CPU restatement of the reference's TIFF reading for colour-depth MIPs.

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
bench.py, never from colormipsearch_b200/.

Follows colormipsearch-api/src/main/java/org/janelia/colormipsearch/imageprocessing/ImageArrayUtils.java:
  * readImageArrayRangeWithTiffReader (:184-227): first image of the file; when the compression is PackBits, walk the strips in
    order, decoding each into ONE zero-initialised byte array with a running output offset (the next strip continues where the
    previous one stopped); other compressions go to ImageJ's Opener (not restated: Pillow plays that part in the tests).
  * packBitsUncompress (:229-258): the run decoder with its [start, end) window.
The tag reading mirrors what LocalTiffDecoder.getTiffInfo (same package, a fork of ImageJ's TiffDecoder) extracts: width,
height, compression, samples per pixel, strip offsets, strip byte counts.

Pinned on the reference's own fixtures (tests/golden/tiff_fixtures.npz: the PackBits / stored TIFFs of
src/test/resources/colormipsearch/api/{imageprocessing,cdsearch/ems} with the pixels Pillow decodes from them, which
ImageArrayUtilsTest.readImageRangeForPackBits asserts equal ImageJ's) -- tests/test_tiff_cpu.py.
"""
import struct

import numpy as np

PACK_BITS = 32773


def packbits_uncompress(inp, output, offset, start, end):
    """ImageArrayUtils.packBitsUncompress (:229-258).  inp: bytes of one strip; output: bytearray / uint8 array written in
    place; returns the new output position."""
    if end == 0:
        end = 2 ** 31 - 1
    index = 0
    pos = offset
    n_in = len(inp)
    n_out = len(output)
    while pos < end and pos < n_out and index < n_in:
        n = inp[index]
        index += 1
        if n >= 128:
            n -= 256
        if n >= 0:
            b = inp[index:index + n + 1]
            index += n + 1
            if len(b) != n + 1:
                raise IndexError("literal run past the end of the strip")          # Java: ArrayIndexOutOfBoundsException
            if pos >= start:
                if pos + len(b) > n_out:
                    raise IndexError("literal run past the end of the image")      # Java: System.arraycopy throws
                output[pos:pos + len(b)] = np.frombuffer(bytes(b), np.uint8)
            elif pos + len(b) >= start:
                output[start:pos + len(b)] = np.frombuffer(bytes(b[start - pos:]), np.uint8)
            pos += len(b)
        elif n != -128:
            length = -n + 1
            v = inp[index]
            index += 1
            for _ in range(length):
                if pos >= start:
                    output[pos] = v          # Java throws past the array; well-formed files never get there
                pos += 1
    return pos


def tiff_info(data):
    """The tags of the first image file directory (what LocalTiffDecoder.getTiffInfo keeps in fi_list[0])."""
    if data[:2] == b"II":
        e = "<"
    elif data[:2] == b"MM":
        e = ">"
    else:
        raise ValueError("not a TIFF file")
    magic, ifd = struct.unpack(e + "HI", data[2:8])
    if magic != 42:
        raise ValueError("not a classic TIFF file")
    (n,) = struct.unpack(e + "H", data[ifd:ifd + 2])
    info = {"compression": 1, "samples_per_pixel": 1, "rows_per_strip": None, "bits_per_sample": [1]}
    names = {256: "width", 257: "height", 258: "bits_per_sample", 259: "compression", 262: "photometric", 273: "strip_offsets",
             277: "samples_per_pixel", 278: "rows_per_strip", 279: "strip_lengths", 284: "planar_config"}
    for i in range(n):
        at = ifd + 2 + 12 * i
        tag, typ, count = struct.unpack(e + "HHI", data[at:at + 8])
        if tag not in names or typ not in (1, 3, 4):
            continue
        size = {1: 1, 3: 2, 4: 4}[typ]
        fmt = {1: "B", 3: "H", 4: "I"}[typ]
        where = at + 8 if size * count <= 4 else struct.unpack(e + "I", data[at + 8:at + 12])[0]
        vals = list(struct.unpack(e + fmt * count, data[where:where + size * count]))
        info[names[tag]] = vals if tag in (258, 273, 279) else vals[0]
    if info["rows_per_strip"] is None:
        info["rows_per_strip"] = info["height"]
    return info


def read_tiff_rgb(data, start=0, end=0):
    """readImageArrayRangeWithTiffReader for a PackBits RGB TIFF (:184-227); stored (uncompressed) files, which the reference
    hands to ImageJ, are read strip by strip.  Returns uint8 [H, W, 3]."""
    info = tiff_info(data)
    W, H = info["width"], info["height"]
    bpp = info["samples_per_pixel"]
    out = np.zeros(W * H * bpp, np.uint8)
    ioffset = 0
    maskpos_st, maskpos_ed = start * 3, end * 3
    if info["compression"] == PACK_BITS:
        for off, ln in zip(info["strip_offsets"], info["strip_lengths"]):
            ioffset = packbits_uncompress(data[off:off + ln], out, ioffset, maskpos_st, maskpos_ed)
            if maskpos_ed and ioffset >= maskpos_ed:
                break
    elif info["compression"] == 1:
        for off, ln in zip(info["strip_offsets"], info["strip_lengths"]):
            ln = min(ln, out.size - ioffset)
            out[ioffset:ioffset + ln] = np.frombuffer(data[off:off + ln], np.uint8)
            ioffset += ln
    else:
        raise NotImplementedError("compression %d is decoded by ImageJ in the reference" % info["compression"])
    return out.reshape(H, W, bpp)
