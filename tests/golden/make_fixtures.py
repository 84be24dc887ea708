#!/usr/bin/env python
"""Regenerate tests/golden/cdsearch_fixtures.npz from the reference's own test images.

Run in the build container only (needs /root/reference and Pillow):

    python tests/golden/make_fixtures.py

The reference (JaneliaSciComp/colormipsearch v3.1.1, BSD-3-Clause, Howard Hughes Medical
Institute) keeps its known-answer inputs as TIFF / PNG files under
colormipsearch-api/src/test/resources/colormipsearch/api/cdsearch/{ems,lms,grad,zgap}.
The GPU box has neither /root/reference nor a guarantee of a TIFF decoder, so the decoded
pixel arrays (exactly what ImageJ's Opener hands to ImageArrayUtils.fromImagePlus,
colormipsearch-api/src/main/java/org/janelia/colormipsearch/imageprocessing/ImageArrayUtils.java:42-96)
are stored here as a compressed npz.  Only pixel DATA is copied, no reference source code.

Keys: em_<short>, lm_<short> : uint8 [H, W, 3] (R,G,B);  grad_<short> : uint16 [H, W];
zgap_<short> : uint8 [H, W, 3].
"""
import os
import sys
import numpy as np
from PIL import Image

REF = "/root/reference/colormipsearch-api/src/test/resources/colormipsearch/api/cdsearch"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cdsearch_fixtures.npz")

FILES = {
    "em_12191": "ems/12191_JRC2018U.tif",
    "em_12191_FL": "ems/12191_JRC2018U_FL.tif",
    "em_LPLC2": "ems/1752016801-LPLC2-RT_18U.tif",
    "lm_BJD": "lms/BJD_127B01_AE_01-20171124_64_H6-40x-Brain-JRC2018_Unisex_20x_HR-2483089192251293794-CH2-01_CDM.tif",
    "lm_GMR": "lms/GMR_31G04_AE_01-20190813_66_F3-40x-Brain-JRC2018_Unisex_20x_HR-2704505419467849826-CH2-07_CDM.tif",
    "lm_VT016795": "lms/VT016795_115C08_AE_01-20200221_61_I2-m-CH1_01.tif",
    "lm_VT033614": "lms/VT033614_127B01_AE_01-20171124_64_H6-f-CH2_01.tif",
    "grad_BJD": "grad/BJD_127B01_AE_01-20171124_64_H6-40x-Brain-JRC2018_Unisex_20x_HR-2483089192251293794-CH2-01_CDM.png",
    "grad_VT016795": "grad/VT016795_115C08_AE_01-20200221_61_I2-m-CH1_01.png",
    "grad_VT033614": "grad/VT033614_127B01_AE_01-20171124_64_H6-f-CH2_01.png",
    "zgap_BJD": "zgap/BJD_127B01_AE_01-20171124_64_H6-40x-Brain-JRC2018_Unisex_20x_HR-2483089192251293794-CH2-01_CDM.tif",
}


def main():
    if not os.path.isdir(REF):
        sys.exit("reference fixtures not found at " + REF)
    arrays = {}
    for key, rel in FILES.items():
        im = Image.open(os.path.join(REF, rel))
        a = np.array(im)
        if key.startswith("grad_"):
            assert a.dtype == np.uint16 and a.ndim == 2, (key, a.dtype, a.shape)
        else:
            assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3, (key, a.dtype, a.shape)
        assert a.shape[:2] == (566, 1210), (key, a.shape)
        arrays[key] = np.ascontiguousarray(a)
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(arrays), "arrays")


if __name__ == "__main__":
    main()
