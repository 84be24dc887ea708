"""Known answers asserted by the reference's own JUnit tests (APIT = colormipsearch-api/src/test/java/org/janelia/colormipsearch)."""

# APIT/cds/PixelMatchColorDepthSearchAlgorithmTest.java
# (mask, target, maskThr, dataThr, zTol, xyShift, mirror, colour-scale label width, expected score, expected mirrored)
PIXEL_MATCH = [
    # :33-53 direct constructor, regions x >= W-260 && y < 90 || x < 330 && y < 100
    ("em_LPLC2", "lm_GMR", 20, 20, 0.01, 2, True, 260, 87, False),
    # :72-103 provider, pixColorFluctuation 1 -> zTol 0.01, ImageTestUtils regions (270)
    ("em_12191", "lm_VT033614", 20, 20, 0.01, 2, True, 270, 439, False),
    ("em_12191", "lm_BJD", 20, 20, 0.01, 2, True, 270, 414, False),
    ("em_12191_FL", "lm_VT033614", 20, 20, 0.01, 2, True, 270, 515, False),
    ("em_12191_FL", "lm_VT016795", 20, 20, 0.01, 2, True, 270, 483, False),
    ("em_12191", "lm_VT016795", 20, 20, 0.01, 2, True, 270, 426, True),
]

# APIT/cds/Shape2DMatchColorDepthSearchAlgorithmTest.java:32-60
SHAPE_MASK_SIZES = [("em_12191_FL", 20, 17340, 70640)]

# (mask, target, gradient, zgap file or None (= synthesised maxFilter(10)), gap, highExpr, score, mirrored)
SHAPE = [
    # :86-132 via the provider
    ("em_12191", "lm_VT033614", "grad_VT033614", None, 21365, 731, 21608, False),
    ("em_12191", "lm_BJD", "grad_BJD", None, 23359, 523, 23533, False),
    ("em_12191", "lm_VT016795", "grad_VT016795", None, 40696, 17253, 46447, True),
    ("em_12191_FL", "lm_VT033614", "grad_VT033614", None, 65381, 677, 65606, False),
    ("em_12191_FL", "lm_VT016795", "grad_VT016795", None, 104449, 16803, 110050, True),
    # :230-291 direct constructor: on-disk zgap file, and a mismatched gradient
    ("em_12191", "lm_BJD", "grad_BJD", "zgap_BJD", 33884, 523, 34058, False),
    ("em_12191", "lm_BJD", "grad_VT033614", None, 23367, 523, 23541, False),
]

# APIT/cds/GradientAreaGapUtilsTest.java:30-49  (pix, gap, highExpr, maxPix, maxNeg, shape, normalized +-0.1)
NORMALIZE = [
    (636, 156, 1897, 679, 1114361, 788, 46833.58),
    (636, 233, 1644, 679, 1107088, 781, 46833.58),
    (636, 0, 1644, 679, 1114361, 548, 46833.58),
    (795, 123, 93, 875, 1606182, 154, 45428.57),
]
