#!/usr/bin/env python
"""Regenerate tests/golden/format_fixtures.npz from the reference's own test files (build container only: needs /root/reference
and Pillow):

    python tests/golden/make_format_fixtures.py

What the reference reads besides PackBits TIFFs: the 16-bit grayscale PNG gradient images under
cdsearch/grad (read through ImageIO.read, ImageArrayUtils.java:98-121) and the LZW TIFFs of
ImageArrayUtilsTest.readImageRangeForOtherCompression (:46-64, which must equal ImageJ's Opener).  The GPU box has no
/root/reference, so the FILE BYTES and the pixels Pillow decodes from them (identical to ImageJ's / ImageIO's for these formats)
are stored here.  Only test DATA (BSD-3-Clause, Howard Hughes Medical Institute) is copied, no reference source code.

Keys: file_<name> : uint8 [file size];  pixels_<name> : uint8 [H, W, 3] (TIFF) or uint16 [H, W] (PNG).
"""
import os
import sys

import numpy as np
from PIL import Image

REF = "/root/reference/colormipsearch-api/src/test/resources/colormipsearch/api"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "format_fixtures.npz")

TIFFS = {"lzw1": "imageprocessing/compressed_lzw1.tif", "lzw2": "imageprocessing/compressed_lzw2.tif"}      # LZW + horizontal predictor
PNGS = {
    "grad_BJD": "cdsearch/grad/BJD_127B01_AE_01-20171124_64_H6-40x-Brain-JRC2018_Unisex_20x_HR-2483089192251293794-CH2-01_CDM.png",
    "grad_VT016795": "cdsearch/grad/VT016795_115C08_AE_01-20200221_61_I2-m-CH1_01.png",
    "grad_VT033614": "cdsearch/grad/VT033614_127B01_AE_01-20171124_64_H6-f-CH2_01.png",
}


def main():
    if not os.path.isdir(REF):
        sys.exit("reference fixtures not found at " + REF)
    arrays = {}
    for key, rel in TIFFS.items():
        path = os.path.join(REF, rel)
        arrays["file_" + key] = np.fromfile(path, dtype=np.uint8)
        arrays["pixels_" + key] = np.array(Image.open(path).convert("RGB"))
    for key, rel in PNGS.items():
        path = os.path.join(REF, rel)
        arrays["file_" + key] = np.fromfile(path, dtype=np.uint8)
        a = np.array(Image.open(path))
        assert a.dtype == np.uint16 and a.ndim == 2, (key, a.dtype, a.shape)
        arrays["pixels_" + key] = a
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
