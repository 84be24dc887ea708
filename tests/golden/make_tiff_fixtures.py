#!/usr/bin/env python
"""Regenerate tests/golden/tiff_fixtures.npz from the reference's own TIFF test files.

Run in the build container only (needs /root/reference and Pillow):

    python tests/golden/make_tiff_fixtures.py

The reference (JaneliaSciComp/colormipsearch v3.1.1, BSD-3-Clause, Howard Hughes Medical Institute) tests its TIFF reading on
src/test/resources/colormipsearch/api/imageprocessing/compressed_{pack,lzw}{1,2}.tif
(ImageArrayUtilsTest.java:19-64: the PackBits range reader must equal ImageJ's Opener) and reads the PackBits colour-depth MIPs
under cdsearch/{ems,lms} in every scoring test.  The GPU box has no /root/reference, so the FILE BYTES of a few small ones and
the pixels Pillow decodes from them (identical to ImageJ's for these formats) are stored here.  Only test DATA is copied, no
reference source code.

Keys: file_<name> : uint8 [file size];  pixels_<name> : uint8 [H, W, 3].
"""
import os
import sys

import numpy as np
from PIL import Image

REF = "/root/reference/colormipsearch-api/src/test/resources/colormipsearch/api"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiff_fixtures.npz")

FILES = {
    "pack1": "imageprocessing/compressed_pack1.tif",          # PackBits, little-endian, one strip, 256 x 256
    "pack2": "imageprocessing/compressed_pack2.tif",          # PackBits, little-endian, one strip, 384 x 384
    "lzw1": "imageprocessing/compressed_lzw1.tif",            # LZW: not decodable on the device
    "stored1": "imageprocessing/minmaxTest1.tif",             # stored, big-endian, one strip, 256 x 256
    "em_12191": "cdsearch/ems/12191_JRC2018U.tif",            # PackBits, big-endian, 71 strips of 8 rows, 1210 x 566
    "em_LPLC2": "cdsearch/ems/1752016801-LPLC2-RT_18U.tif",
    "lm_GMR": "cdsearch/lms/GMR_31G04_AE_01-20190813_66_F3-40x-Brain-JRC2018_Unisex_20x_HR-2704505419467849826-CH2-07_CDM.tif",
}


def main():
    if not os.path.isdir(REF):
        sys.exit("reference fixtures not found at " + REF)
    arrays = {}
    for key, rel in FILES.items():
        path = os.path.join(REF, rel)
        arrays["file_" + key] = np.fromfile(path, dtype=np.uint8)
        a = np.array(Image.open(path).convert("RGB"))
        assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3, (key, a.dtype, a.shape)
        arrays["pixels_" + key] = a
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
