/*
 * cds_oracle.c -- CPU restatement of colormipsearch's colour-depth matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * Nothing under colormipsearch_b200/ links, imports or calls it, and the product library has
 * no CPU fallback.
 *
 * Parity status: PINNED.  The reference (JaneliaSciComp/colormipsearch v3.1.1) is Java and
 * cannot be run in this image (no JVM), so this is a loop-for-loop restatement in plain C,
 * built with -O2 -ffp-contract=off (Java never contracts a*b+c).  It is pinned by every
 * known-answer vector the reference's own JUnit tests hold for this path
 * (tests/test_oracle_golden.py):  6 pixel-match vectors
 * (colormipsearch-api/src/test/java/org/janelia/colormipsearch/cds/PixelMatchColorDepthSearchAlgorithmTest.java:33-103),
 * 2 mask-size vectors + 11 shape-score vectors
 * (.../cds/Shape2DMatchColorDepthSearchAlgorithmTest.java:32-60, 86-132, 230-291) and 4 normalisation
 * vectors (.../cds/GradientAreaGapUtilsTest.java:30-49).
 *
 * Citations use the prefix API/ = colormipsearch-api/src/main/java/org/janelia/colormipsearch/.
 * Pixel layout everywhere: RGB images are uint8[H][W][3] in R,G,B order exactly as
 * API/imageprocessing/ColorImageArray.java:6-31 stores them; gray16 is uint16[H][W]
 * (API/imageprocessing/ShortImageArray.java:4-16); gray8 is uint8[H][W].
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CDSO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * a5: AbstractColorDepthSearchAlgorithm.calculatePixelGap
 * API/cds/AbstractColorDepthSearchAlgorithm.java:157-390
 * Sector ids: 1=BR 2=BG 3=GB 4=GR 5=RG 6=RB, 0 = none (ties for the maximum).
 * The reference keeps twelve "XY1/xy1" variables of which at most one pair is non-zero; a
 * (sector, ratio) pair is the same information.  The `== 255` branches (:266, :282 ...) are
 * unreachable because every ratio is < 1; they are kept for fidelity.
 * ------------------------------------------------------------------------------------------ */
static void cdso_classify(int red, int green, int blue, int *sector, double *ratio)
{
    *sector = 0;
    *ratio = 0.0;
    if (blue > red && blue > green) {                 /* :195 */
        if (red > green) {
            *sector = 1;                              /* BR; BR1 = blue+red > 0 */
            if (blue != 0 && red != 0) *ratio = (double) red / (double) blue;
        } else {
            *sector = 2;                              /* BG */
            if (blue != 0 && green != 0) *ratio = (double) green / (double) blue;
        }
    } else if (green > blue && green > red) {          /* :206 */
        if (blue > red) {
            *sector = 3;                              /* GB */
            if (green != 0 && blue != 0) *ratio = (double) blue / (double) green;
        } else {
            *sector = 4;                              /* GR */
            if (green != 0 && red != 0) *ratio = (double) red / (double) green;
        }
    } else if (red > blue && red > green) {            /* :217 */
        if (green > blue) {
            *sector = 5;                              /* RG */
            if (red != 0 && green != 0) *ratio = (double) green / (double) red;
        } else {
            *sector = 6;                              /* RB */
            if (red != 0 && blue != 0) *ratio = (double) blue / (double) red;
        }
    }
}

static double cdso_same_sector_gap(double a1, double a2)
{
    /* e.g. :262-270 */
    double pxGap = 10000;
    if (a1 > 0 && a2 > 0) {
        if (a1 != a2) pxGap = fabs(a2 - a1);
        else pxGap = 0;
        if ((a1 == 255) & (a2 == 255)) pxGap = 1000;
    }
    return pxGap;
}

CDSO_API double cdso_pixel_gap(int red1, int green1, int blue1, int red2, int green2, int blue2)
{
    const double BrBg = 0.354862745;
    const double BgGb = 0.996078431;
    const double GbGr = 0.505882353;
    const double GrRg = 0.996078431;
    const double RgRb = 0.505882353;
    double pxGap = 10000;
    int s1, s2;
    double r1, r2;
    cdso_classify(red1, green1, blue1, &s1, &r1);
    cdso_classify(red2, green2, blue2, &s2, &r2);

    if (s1 == 1) {                                     /* BR mask :260 */
        if (s2 == 1) {
            pxGap = cdso_same_sector_gap(r1, r2);
        } else if (s2 == 2) {                          /* :271 */
            if (r1 < 0.44 && r2 < 0.54) {
                double BrGap = r1 - BrBg;
                double BgGap = r2 - BrBg;
                pxGap = BrGap + BgGap;
            }
        }
    } else if (s1 == 2) {                              /* BG mask :278 */
        if (s2 == 2) {
            pxGap = cdso_same_sector_gap(r1, r2);
        } else if (s2 == 3) {                          /* :288 */
            if (r1 > 0.8 && r2 > 0.8) {
                double BgGap = BgGb - r1;
                double GbGap = BgGb - r2;
                pxGap = BgGap + GbGap;
            }
        } else if (s2 == 1) {                          /* :294 */
            if (r1 < 0.54 && r2 < 0.44) {
                double BgGap = r1 - BrBg;
                double BrGap = r2 - BrBg;
                pxGap = BrGap + BgGap;
            }
        }
    } else if (s1 == 3) {                              /* GB mask :301 */
        if (s2 == 3) {
            pxGap = cdso_same_sector_gap(r1, r2);
        } else if (s2 == 2) {                          /* :311 */
            if (r1 > 0.8 && r2 > 0.8) {
                double BgGap = BgGb - r1;
                double GbGap = BgGb - r2;
                pxGap = BgGap + GbGap;
            }
        } else if (s2 == 4) {                          /* :317 */
            if (r1 < 0.7 && r2 < 0.7) {
                double GbGap = r1 - GbGr;
                double GrGap = r2 - GbGr;
                pxGap = GbGap + GrGap;
            }
        }
    } else if (s1 == 4) {                              /* GR mask :324 */
        if (s2 == 4) {
            pxGap = cdso_same_sector_gap(r1, r2);
        } else if (s2 == 3) {                          /* :334 */
            if (r1 < 0.7 && r2 < 0.7) {
                double GrGap = r1 - GbGr;
                double GbGap = r2 - GbGr;
                pxGap = GrGap + GbGap;
            }
        } else if (s2 == 5) {                          /* :340 */
            if (r1 > 0.8 && r2 > 0.8) {
                double GrGap = GrRg - r1;
                double RgGap = GrRg - r2;
                pxGap = GrGap + RgGap;
            }
        }
    } else if (s1 == 5) {                              /* RG mask :347 */
        if (s2 == 5) {
            pxGap = cdso_same_sector_gap(r1, r2);
        } else if (s2 == 4) {                          /* :358 */
            if (r1 > 0.8 && r2 > 0.8) {
                double GrGap = GrRg - r2;
                double RgGap = GrRg - r1;
                pxGap = GrGap + RgGap;
            }
        } else if (s2 == 6) {                          /* :364 */
            if (r1 < 0.7 && r2 < 0.7) {
                double RgGap = r1 - RgRb;
                double RbGap = r2 - RgRb;
                pxGap = RbGap + RgGap;
            }
        }
    } else if (s1 == 6) {                              /* RB mask :371 */
        if (s2 == 6) {
            pxGap = cdso_same_sector_gap(r1, r2);
        } else if (s2 == 5) {                          /* :381 */
            if (r2 < 0.7 && r1 < 0.7) {
                double RgGap = r2 - RgRb;
                double RbGap = r1 - RgRb;
                pxGap = RgGap + RbGap;
            }
        }
    }
    return pxGap;
}

/* ------------------------------------------------------------------------------------------
 * Excluded (label) regions: a union of half-open rectangles [x0,x1) x [y0,y1), the shape of
 * colormipsearch-tools/.../cmd/AbstractColorDepthMatchArgs.java:101-119.
 * rects = int32[nrects][4] = {x0, y0, x1, y1}.
 * ------------------------------------------------------------------------------------------ */
static int cdso_in_rects(int x, int y, const int32_t *rects, int nrects)
{
    for (int i = 0; i < nrects; i++) {
        const int32_t *r = rects + 4 * i;
        if (x >= r[0] && x < r[2] && y >= r[1] && y < r[3]) return 1;
    }
    return 0;
}

/* a2: AbstractColorDepthSearchAlgorithm.getMaskPosArray  API/cds/AbstractColorDepthSearchAlgorithm.java:96-126
 * Returns P; writes ascending pixel indices into pos_out (capacity W*H) and the bounding box
 * {minx, miny, maxx, maxy} (maxx/maxy exclusive) into bbox_out[4] if non-NULL. */
CDSO_API int cdso_mask_positions(const uint8_t *rgb, int W, int H, int thresm,
                                 const int32_t *rects, int nrects,
                                 int32_t *pos_out, int32_t *bbox_out)
{
    int n = 0;
    int minx = W, miny = H, maxx = 0, maxy = 0;
    int sumpx = W * H;
    for (int pi = 0; pi < sumpx; pi++) {
        int x = pi % W;
        int y = pi / W;
        if (cdso_in_rects(x, y, rects, nrects)) continue;
        int red = rgb[3 * pi], green = rgb[3 * pi + 1], blue = rgb[3 * pi + 2];
        if (red > thresm || green > thresm || blue > thresm) {
            pos_out[n++] = pi;
            if (x < minx) minx = x;
            if (x + 1 > maxx) maxx = x + 1;
            if (y < miny) miny = y;
            if (y + 1 > maxy) maxy = y + 1;
        }
    }
    if (bbox_out) { bbox_out[0] = minx; bbox_out[1] = miny; bbox_out[2] = maxx; bbox_out[3] = maxy; }
    return n;
}

/* a3: shift offsets in the order of generateShiftedMasks, API/cds/PixelMatchColorDepthSearchAlgorithm.java:113-130.
 * Java allocates 1 + 8*(xyShift/2) slots but the loops emit 9*(xyShift/2) entries, so it only
 * runs for xyShift in {0, 2} (ArrayIndexOutOfBounds for >= 4).  For xyShift >= 4 this oracle is the
 * specification: the entries the loops intend, duplicates of (0,0) included --
 * a duplicate cannot change a max.  Returns the number of offsets written (<= cap). */
CDSO_API int cdso_shift_offsets(int xyshift, int32_t *dx_out, int32_t *dy_out, int cap)
{
    int n = 0;
    int nshifts = 1 + (xyshift / 2) * 8;
    if (nshifts > 1) {
        for (int i = 2; i <= xyshift; i += 2)
            for (int xx = -i; xx <= i; xx += i)
                for (int yy = -i; yy <= i; yy += i) {
                    if (n < cap) { dx_out[n] = xx; dy_out[n] = yy; }
                    n++;
                }
    } else {
        if (cap > 0) { dx_out[0] = 0; dy_out[0] = 0; }
        n = 1;
    }
    return n;
}

/* 1 when the Java reference itself would throw for this xyShift (see above). */
CDSO_API int cdso_reference_throws_for_xyshift(int xyshift)
{
    if (xyshift & 1) return 1; /* IllegalArgumentException, API/cds/ColorDepthSearchAlgorithmProviderFactory.java:57-60 */
    return 9 * (xyshift / 2) > 1 + 8 * (xyshift / 2) && xyshift > 0;
}

/* shiftMaskPosArray :132-144 */
static void cdso_shift_positions(const int32_t *pos, int P, int xshift, int yshift, int W, int H, int32_t *out)
{
    for (int i = 0; i < P; i++) {
        int pixelCoord = pos[i];
        int x = (pixelCoord % W) + xshift;
        int y = pixelCoord / W + yshift;
        if (x >= 0 && x < W && y >= 0 && y < H) out[i] = y * W + x;
        else out[i] = -1;
    }
}

/* mirrorMask :146-158 */
static void cdso_mirror_positions(const int32_t *pos, int P, int ypitch, int32_t *out)
{
    for (int i = 0; i < P; i++) {
        int pixelCoord = pos[i];
        if (pixelCoord == -1) out[i] = -1;
        else {
            int x = pixelCoord % ypitch;
            out[i] = pixelCoord + (ypitch - 1) - 2 * x;
        }
    }
}

/* A prepared mask = what the PixelMatchColorDepthSearchAlgorithm constructor holds (:29-101). */
typedef struct cdso_mask {
    int W, H;
    int P;
    int nvariants;           /* per orientation */
    int mirror;
    int target_threshold;
    double z_tolerance;
    uint8_t *rgb;            /* copy of the query image */
    int32_t *pos;            /* [P] */
    int32_t *variants;       /* [nvariants][P] */
    int32_t *mvariants;      /* [nvariants][P] or NULL */
    int32_t bbox[4];
} cdso_mask;

CDSO_API cdso_mask *cdso_mask_create(const uint8_t *rgb, int W, int H, int query_threshold, int mirror,
                                     int target_threshold, double z_tolerance, int xyshift,
                                     const int32_t *rects, int nrects)
{
    cdso_mask *m = (cdso_mask *) calloc(1, sizeof(cdso_mask));
    size_t npx = (size_t) W * H;
    m->W = W; m->H = H; m->mirror = mirror;
    m->target_threshold = target_threshold;
    m->z_tolerance = z_tolerance;
    m->rgb = (uint8_t *) malloc(npx * 3);
    memcpy(m->rgb, rgb, npx * 3);
    int32_t *tmp = (int32_t *) malloc(npx * sizeof(int32_t));
    m->P = cdso_mask_positions(rgb, W, H, query_threshold, rects, nrects, tmp, m->bbox);
    m->pos = (int32_t *) malloc((size_t) (m->P > 0 ? m->P : 1) * sizeof(int32_t));
    memcpy(m->pos, tmp, (size_t) m->P * sizeof(int32_t));
    free(tmp);
    int cap = 9 * (xyshift / 2) + 1;
    int32_t *dx = (int32_t *) malloc(cap * sizeof(int32_t));
    int32_t *dy = (int32_t *) malloc(cap * sizeof(int32_t));
    m->nvariants = cdso_shift_offsets(xyshift, dx, dy, cap);
    size_t vsz = (size_t) m->nvariants * (size_t) (m->P > 0 ? m->P : 1);
    m->variants = (int32_t *) malloc(vsz * sizeof(int32_t));
    for (int v = 0; v < m->nvariants; v++) {
        if (m->nvariants == 1) memcpy(m->variants, m->pos, (size_t) m->P * sizeof(int32_t));  /* :126 out[0] = pixelCoords */
        else cdso_shift_positions(m->pos, m->P, dx[v], dy[v], W, H, m->variants + (size_t) v * m->P);
    }
    if (mirror) {
        m->mvariants = (int32_t *) malloc(vsz * sizeof(int32_t));
        for (int v = 0; v < m->nvariants; v++)
            cdso_mirror_positions(m->variants + (size_t) v * m->P, m->P, W, m->mvariants + (size_t) v * m->P);
    }
    free(dx); free(dy);
    return m;
}

CDSO_API void cdso_mask_destroy(cdso_mask *m)
{
    if (!m) return;
    free(m->rgb); free(m->pos); free(m->variants); free(m->mvariants); free(m);
}

CDSO_API int cdso_mask_size(const cdso_mask *m) { return m->P; }
CDSO_API int cdso_mask_nvariants(const cdso_mask *m) { return m->nvariants * (m->mirror ? 2 : 1); }
CDSO_API void cdso_mask_get_positions(const cdso_mask *m, int32_t *out) { memcpy(out, m->pos, (size_t) m->P * sizeof(int32_t)); }
CDSO_API void cdso_mask_get_bbox(const cdso_mask *m, int32_t *out) { memcpy(out, m->bbox, sizeof(m->bbox)); }

/* calculateScore :235-263 */
static int cdso_calculate_score(const cdso_mask *m, const uint8_t *target, const int32_t *target_positions)
{
    int score = 0;
    for (int i = 0; i < m->P; i++) {
        int srcPos = m->pos[i];
        int targetPos = target_positions[i];
        if (targetPos == -1 || srcPos == -1) continue;
        int red2 = target[3 * targetPos], green2 = target[3 * targetPos + 1], blue2 = target[3 * targetPos + 2];
        if (red2 > m->target_threshold || green2 > m->target_threshold || blue2 > m->target_threshold) {
            int red1 = m->rgb[3 * srcPos], green1 = m->rgb[3 * srcPos + 1], blue1 = m->rgb[3 * srcPos + 2];
            double pxGap = cdso_pixel_gap(red1, green1, blue1, red2, green2, blue2);
            if (pxGap <= m->z_tolerance) score++;
        }
    }
    return score;
}

/* calculateMaxScoreForAllTargetTransformations :221-233 */
static int cdso_max_score(const cdso_mask *m, const uint8_t *target, const int32_t *variants)
{
    int maxScore = 0;
    for (int v = 0; v < m->nvariants; v++) {
        int score = cdso_calculate_score(m, target, variants + (size_t) v * m->P);
        if (score > maxScore) maxScore = score;
    }
    return maxScore;
}

/* a4: calculateMatchingScore :166-219 (negative-query branch omitted: the tools always pass a null
 * negative image, API/cds/ColorDepthSearchAlgorithmProviderFactory.java:61-71).
 * Size mismatch is the caller's job (returns -1 here = IllegalArgumentException at :171-175). */
CDSO_API int cdso_mask_score(const cdso_mask *m, const uint8_t *target, int tW, int tH,
                             int32_t *score_out, double *ratio_out, int32_t *mirrored_out)
{
    if (m->P == 0) { *score_out = 0; *ratio_out = 0; *mirrored_out = 0; return 0; }   /* :169-170 */
    if (tW != m->W || tH != m->H) return -1;
    int maxMatchingPixels = cdso_max_score(m, target, m->variants);
    int bestScoreMirrored = 0;
    if (m->mvariants) {
        int mirroredMax = cdso_max_score(m, target, m->mvariants);
        if (mirroredMax > maxMatchingPixels) {                       /* strict :189 */
            maxMatchingPixels = mirroredMax;
            bestScoreMirrored = 1;
        }
    }
    *score_out = maxMatchingPixels;
    *ratio_out = (double) maxMatchingPixels / (double) m->P;         /* :194 */
    *mirrored_out = bestScoreMirrored;
    return 0;
}

/* Per-variant counts, for debugging kernels: out[nvariants * (mirror?2:1)] normal first then mirrored. */
CDSO_API void cdso_mask_variant_scores(const cdso_mask *m, const uint8_t *target, int32_t *out)
{
    for (int v = 0; v < m->nvariants; v++)
        out[v] = m->P ? cdso_calculate_score(m, target, m->variants + (size_t) v * m->P) : 0;
    if (m->mvariants)
        for (int v = 0; v < m->nvariants; v++)
            out[m->nvariants + v] = m->P ? cdso_calculate_score(m, target, m->mvariants + (size_t) v * m->P) : 0;
}

/* a6: ColorMIPSearch.isMatch API/cds/ColorMIPSearch.java:42-45 with PixelMatchScore.getNormalizedScore
 * (API/cds/PixelMatchScore.java:24-26): the ratio is narrowed to float, then compared as double. */
CDSO_API int cdso_is_match(int score, double ratio, double pct_positive_pixels)
{
    double thr = pct_positive_pixels / 100;
    float normalized = (float) ratio;
    return score > 0 && (double) normalized > thr;
}

/* The mask x target sweep the tools run (LocalColorMIPSearchProcessor.java:55-116), threaded over
 * targets the way the reference fans targets out over its pool.  scores/mirrored are [nmasks][ntargets].
 * Returns the number of threads actually used. */
CDSO_API int cdso_search_dense(cdso_mask *const *masks, int nmasks, const uint8_t *targets, int64_t ntargets,
                               int nthreads, int32_t *scores, uint8_t *mirrored)
{
    int used = 1;
    if (nmasks == 0 || ntargets == 0) return used;
    size_t tsz = (size_t) masks[0]->W * masks[0]->H * 3;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) collapse(2)
#endif
    for (int64_t t = 0; t < ntargets; t++) {
        for (int mi = 0; mi < nmasks; mi++) {
            int32_t s, mir; double r;
            cdso_mask_score(masks[mi], targets + (size_t) t * tsz, masks[mi]->W, masks[mi]->H, &s, &r, &mir);
            scores[(size_t) mi * ntargets + t] = s;
            if (mirrored) mirrored[(size_t) mi * ntargets + t] = (uint8_t) mir;
        }
    }
    return used;
}

/* ------------------------------------------------------------------------------------------
 * a11: GradientAreaGapUtils  API/cds/GradientAreaGapUtils.java
 * ------------------------------------------------------------------------------------------ */
static const short CDSO_LUT[256][3] = {   /* :132-155, the colour-depth lookup table PsychedelicRainBow2 */
    {127, 0, 255}, {125, 3, 255}, {124, 6, 255}, {122, 9, 255}, {121, 12, 255}, {120, 15, 255}, {119, 18, 255}, {118, 21, 255}, {116, 24, 255}, {115, 27, 255}, {114, 30, 255}, {113, 33, 255},
    {112, 36, 255}, {110, 39, 255}, {109, 42, 255}, {108, 45, 255}, {106, 48, 255}, {105, 51, 255}, {104, 54, 255}, {103, 57, 255}, {101, 60, 255}, {100, 63, 255}, {99, 66, 255}, {98, 69, 255},
    {96, 72, 255}, {95, 75, 255}, {94, 78, 255}, {93, 81, 255}, {92, 84, 255}, {90, 87, 255}, {89, 90, 255}, {87, 93, 255}, {86, 96, 255}, {84, 99, 255}, {83, 102, 255}, {81, 105, 255},
    {80, 108, 255}, {78, 111, 255}, {77, 114, 255}, {75, 117, 255}, {74, 120, 255}, {72, 123, 255}, {71, 126, 255}, {69, 129, 255}, {68, 132, 255}, {66, 135, 255}, {65, 138, 255}, {63, 141, 255},
    {62, 144, 255}, {60, 147, 255}, {59, 150, 255}, {57, 153, 255}, {56, 156, 255}, {54, 159, 255}, {53, 162, 255}, {51, 165, 255}, {50, 168, 255}, {48, 171, 255}, {47, 174, 255}, {45, 177, 255},
    {44, 180, 255}, {42, 183, 255}, {41, 186, 255}, {39, 189, 255}, {38, 192, 255}, {36, 195, 255}, {35, 198, 255}, {33, 201, 255}, {32, 204, 255}, {30, 207, 255}, {29, 210, 255}, {27, 213, 255},
    {26, 216, 255}, {24, 219, 255}, {23, 222, 255}, {21, 225, 255}, {20, 228, 255}, {18, 231, 255}, {16, 234, 255}, {14, 237, 255}, {12, 240, 255}, {9, 243, 255}, {6, 246, 255}, {3, 249, 255},
    {1, 252, 255}, {0, 254, 255}, {3, 255, 252}, {6, 255, 249}, {9, 255, 246}, {12, 255, 243}, {15, 255, 240}, {18, 255, 237}, {21, 255, 234}, {24, 255, 231}, {27, 255, 228}, {30, 255, 225},
    {33, 255, 222}, {36, 255, 219}, {39, 255, 216}, {42, 255, 213}, {45, 255, 210}, {48, 255, 207}, {51, 255, 204}, {54, 255, 201}, {57, 255, 198}, {60, 255, 195}, {63, 255, 192}, {66, 255, 189},
    {69, 255, 186}, {72, 255, 183}, {75, 255, 180}, {78, 255, 177}, {81, 255, 174}, {84, 255, 171}, {87, 255, 168}, {90, 255, 165}, {93, 255, 162}, {96, 255, 159}, {99, 255, 156}, {102, 255, 153},
    {105, 255, 150}, {108, 255, 147}, {111, 255, 144}, {114, 255, 141}, {117, 255, 138}, {120, 255, 135}, {123, 255, 132}, {126, 255, 129}, {129, 255, 126}, {132, 255, 123}, {135, 255, 120},
    {138, 255, 117}, {141, 255, 114}, {144, 255, 111}, {147, 255, 108}, {150, 255, 105}, {153, 255, 102}, {156, 255, 99}, {159, 255, 96}, {162, 255, 93}, {165, 255, 90}, {168, 255, 87}, {171, 255, 84},
    {174, 255, 81}, {177, 255, 78}, {180, 255, 75}, {183, 255, 72}, {186, 255, 69}, {189, 255, 66}, {192, 255, 63}, {195, 255, 60}, {198, 255, 57}, {201, 255, 54}, {204, 255, 51}, {207, 255, 48},
    {210, 255, 45}, {213, 255, 42}, {216, 255, 39}, {219, 255, 36}, {222, 255, 33}, {225, 255, 30}, {228, 255, 27}, {231, 255, 24}, {234, 255, 21}, {237, 255, 18}, {240, 255, 15}, {243, 255, 12},
    {246, 255, 9}, {249, 255, 6}, {252, 255, 3}, {254, 255, 0}, {255, 252, 3}, {255, 249, 6}, {255, 246, 9}, {255, 243, 12}, {255, 240, 15}, {255, 237, 18}, {255, 234, 21}, {255, 231, 24}, {255, 228, 27},
    {255, 225, 30}, {255, 222, 33}, {255, 219, 36}, {255, 216, 39}, {255, 213, 42}, {255, 210, 45}, {255, 207, 48}, {255, 204, 51}, {255, 201, 54}, {255, 198, 57}, {255, 195, 60}, {255, 192, 63},
    {255, 189, 66}, {255, 186, 69}, {255, 183, 72}, {255, 180, 75}, {255, 177, 78}, {255, 174, 81}, {255, 171, 84}, {255, 168, 87}, {255, 165, 90}, {255, 162, 93}, {255, 159, 96}, {255, 156, 99},
    {255, 153, 102}, {255, 150, 105}, {255, 147, 108}, {255, 144, 111}, {255, 141, 114}, {255, 138, 117}, {255, 135, 120}, {255, 132, 123}, {255, 129, 126}, {255, 126, 129}, {255, 123, 132},
    {255, 120, 135}, {255, 117, 138}, {255, 114, 141}, {255, 111, 144}, {255, 108, 147}, {255, 105, 150}, {255, 102, 153}, {255, 99, 156}, {255, 96, 159}, {255, 93, 162}, {255, 90, 165}, {255, 87, 168},
    {255, 84, 171}, {255, 81, 173}, {255, 78, 174}, {255, 75, 175}, {255, 72, 176}, {255, 69, 177}, {255, 66, 178}, {255, 63, 179}, {255, 60, 180}, {255, 57, 181}, {255, 54, 182}, {255, 51, 183},
    {255, 48, 184}, {255, 45, 185}, {255, 42, 186}, {255, 39, 187}, {255, 36, 188}, {255, 33, 189}, {255, 30, 190}, {255, 27, 191}, {255, 24, 192}, {255, 21, 193}, {255, 18, 194}, {255, 15, 195},
    {255, 12, 196}, {255, 9, 197}, {255, 6, 198}, {255, 3, 199}, {255, 0, 200}
};

CDSO_API void cdso_lut(int32_t *out /* [256][3] */)
{
    for (int i = 0; i < 256; i++) for (int c = 0; c < 3; c++) out[3 * i + c] = CDSO_LUT[i][c];
}

/* findSliceNumberInLUT :131-197 */
static int cdso_find_slice_in_lut(int lutStartRange, int lutEndRange, double colorRatio)
{
    int sliceNumber = 0;
    double mingapratio = 1000;
    for (int icolor = lutStartRange; icolor <= lutEndRange; icolor++) {
        double lutRatio = 0;
        double colorR = CDSO_LUT[icolor][0];
        double colorG = CDSO_LUT[icolor][1];
        double colorB = CDSO_LUT[icolor][2];
        if (colorB > colorR && colorB > colorG) {
            if (colorR > colorG) lutRatio = colorR / colorB;
            else if (colorG > colorR) lutRatio = colorG / colorB;
        } else if (colorG > colorR && colorG > colorB) {
            if (colorR > colorB) lutRatio = colorR / colorG;
            else if (colorB > colorR) lutRatio = colorB / colorG;
        } else if (colorR > colorG && colorR > colorB) {
            if (colorG > colorB) lutRatio = colorG / colorR;
            else if (colorB > colorG) lutRatio = colorB / colorR;
        }
        if (lutRatio == colorRatio) return icolor + 1;
        double gapratio = fabs(colorRatio - lutRatio);
        if (gapratio < mingapratio) {
            mingapratio = gapratio;
            sliceNumber = icolor + 1;
        }
    }
    return sliceNumber;
}

enum { CDSO_BLACK = 0, CDSO_RED = 1, CDSO_GREEN = 2, CDSO_BLUE = 3 };

/* findSliceNumber :107-129 */
static int cdso_find_slice_number(int maxColor, int secondMaxColor, double colorRatio)
{
    switch (maxColor) {
        case CDSO_RED:
            if (secondMaxColor == CDSO_GREEN) return cdso_find_slice_in_lut(171, 212, colorRatio);
            else if (secondMaxColor == CDSO_BLUE) return cdso_find_slice_in_lut(213, 255, colorRatio);
            break;
        case CDSO_GREEN:
            if (secondMaxColor == CDSO_RED) return cdso_find_slice_in_lut(128, 170, colorRatio);
            if (secondMaxColor == CDSO_BLUE) return cdso_find_slice_in_lut(86, 127, colorRatio);
            break;
        case CDSO_BLUE:
            if (secondMaxColor == CDSO_RED) return cdso_find_slice_in_lut(0, 29, colorRatio);
            if (secondMaxColor == CDSO_GREEN) return cdso_find_slice_in_lut(30, 85, colorRatio);
            break;
    }
    return 0;
}

/* first half of calculateSliceGap :18-99: the slice number of one colour. */
CDSO_API int cdso_slice_number(int red, int green, int blue)
{
    int max1 = 0, max2 = 0, c1 = CDSO_BLACK, c2 = CDSO_BLACK;
    if (red >= green && red >= blue) {
        max1 = red; c1 = CDSO_RED;
        if (green >= blue) { max2 = green; c2 = CDSO_GREEN; } else { max2 = blue; c2 = CDSO_BLUE; }
    } else if (green >= red && green >= blue) {
        max1 = green; c1 = CDSO_GREEN;
        if (red >= blue) { c2 = CDSO_RED; max2 = red; } else { max2 = blue; c2 = CDSO_BLUE; }
    } else if (blue >= red && blue >= green) {
        max1 = blue; c1 = CDSO_BLUE;
        if (red >= green) { max2 = red; c2 = CDSO_RED; } else { max2 = green; c2 = CDSO_GREEN; }
    }
    double ratio = (double) max2 / (double) max1;       /* NaN for black: every compare fails -> 0 */
    return cdso_find_slice_number(c1, c2, ratio);
}

/* Batch forms for the exhaustive tests (all 2^24 colours): out[i] = slice number / gray of colour rgb[3i..3i+2]. */
CDSO_API void cdso_slice_numbers(const uint8_t *rgb, int64_t n, uint16_t *out)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < n; i++) out[i] = (uint16_t) cdso_slice_number(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
}

/* calculateSliceGap :18-105; rgb ints are 0xAARRGGBB */
CDSO_API int cdso_slice_gap(int rgb1, int rgb2)
{
    int maskslinumber = cdso_slice_number((rgb1 >> 16) & 0xff, (rgb1 >> 8) & 0xff, rgb1 & 0xff);
    int dataslinumber = cdso_slice_number((rgb2 >> 16) & 0xff, (rgb2 >> 8) & 0xff, rgb2 & 0xff);
    if (dataslinumber == 0 || maskslinumber == 0) return dataslinumber;
    return abs(maskslinumber - dataslinumber);
}

/* a12 */
CDSO_API int64_t cdso_shape_score_2d(int64_t gradientAreaGap, int64_t highExpressionArea)   /* :199-207 */
{
    if (gradientAreaGap >= 0 && highExpressionArea >= 0) return gradientAreaGap + highExpressionArea / 3;
    return -1;
}

CDSO_API double cdso_normalized_score(int pixelMatchScore, int64_t shapeScore, int64_t maxPixelMatch, int64_t maxShapeScore) /* :219-235 */
{
    if (pixelMatchScore == 0 || maxPixelMatch == 0 || shapeScore < 0 || maxShapeScore <= 0) {
        return pixelMatchScore;
    } else {
        double normalizedPixelScore = (double) pixelMatchScore / maxPixelMatch;
        double normalizedShapeScore = (double) shapeScore / maxShapeScore;
        double boundedShapeScore = fmin(fmax(normalizedShapeScore * 2.5, 0.002), 1.);
        return normalizedPixelScore / boundedShapeScore * 100;
    }
}

/* ------------------------------------------------------------------------------------------
 * a9: ColorTransformation  API/imageprocessing/ColorTransformation.java
 * ------------------------------------------------------------------------------------------ */
CDSO_API int cdso_rgb_to_gray(int r, int g, int b)     /* rgbToGrayNoGammaCorrection :40-54 with maxGrayValue = 255 (:87, :103) */
{
    if (r == 0 && g == 0 && b == 0) return 0;
    float maxGrayValue = 255;
    double rw = 1 / 3.;
    double gw = 1 / 3.;
    double bw = 1 / 3.;
    return (int) ((maxGrayValue / 255) * (r * rw + g * gw + b * bw + 0.5));
}

CDSO_API void cdso_rgb_to_gray_batch(const uint8_t *rgb, int64_t n, uint8_t *out)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < n; i++) out[i] = (uint8_t) cdso_rgb_to_gray(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
}

/* ColorTransformation.mask(threshold) on RGB (:29-38, :114-132): black when every channel <= threshold */
static inline int cdso_rgb_passes(const uint8_t *p, int threshold)
{
    return p[0] > threshold || p[1] > threshold || p[2] > threshold;
}

CDSO_API void cdso_mask_rgb(const uint8_t *src, int W, int H, int threshold, uint8_t *dst)
{
    size_t n = (size_t) W * H;
    for (size_t i = 0; i < n; i++) {
        if (cdso_rgb_passes(src + 3 * i, threshold)) { dst[3*i] = src[3*i]; dst[3*i+1] = src[3*i+1]; dst[3*i+2] = src[3*i+2]; }
        else { dst[3*i] = dst[3*i+1] = dst[3*i+2] = 0; }
    }
}

/* ImageTransformation.clearRegion :182-193 */
CDSO_API void cdso_clear_regions(const uint8_t *src, int W, int H, const int32_t *rects, int nrects, uint8_t *dst)
{
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            size_t i = (size_t) y * W + x;
            if (cdso_in_rects(x, y, rects, nrects)) { dst[3*i] = dst[3*i+1] = dst[3*i+2] = 0; }
            else { dst[3*i] = src[3*i]; dst[3*i+1] = src[3*i+1]; dst[3*i+2] = src[3*i+2]; }
        }
}

/* ------------------------------------------------------------------------------------------
 * a8: max filter.  makeLineRadii  API/imageprocessing/ImageTransformation.java:549-572 gives the ImageJ
 * RankFilters disc: rows dy in [-kRadius, kRadius] with half-width dx(dy).  With border 0 and row-major
 * traversal unsafeMaxFilter (:353-535) is the plain per-channel dilation by that disc with pixels outside
 * the image ignored (the histograms drop zeros, :99-104, and the row cache is zero padded, :498-502); this
 * equivalence is what the 70 640 high-expression vector and the maxFilter(10) zgap vectors pin.
 * Returns kRadius; writes half-widths for dy = -kRadius..kRadius into dx_out (capacity 2*kRadius+1) if non-NULL.
 * ------------------------------------------------------------------------------------------ */
CDSO_API int cdso_line_radii(double radiusArg, int32_t *dx_out, int cap)
{
    double radius;
    if (radiusArg >= 1.5 && radiusArg < 1.75) radius = 1.75;
    else if (radiusArg >= 2.5 && radiusArg < 2.85) radius = 2.85;
    else radius = radiusArg;
    int r2 = (int) (radius * radius) + 1;
    int kRadius = (int) (sqrt(r2 + 1e-10));
    if (dx_out) {
        for (int y = -kRadius; y <= kRadius; y++) {
            int dx = (y == 0) ? kRadius : (int) (sqrt(r2 - y * y + 1e-10));
            if (y + kRadius < cap) dx_out[y + kRadius] = dx;
        }
    }
    return kRadius;
}

/* brute force, obviously-correct version (use on small images / small radii) */
CDSO_API void cdso_max_filter_bruteforce(const uint8_t *src, int W, int H, int nch, double radius, uint8_t *dst)
{
    int k = cdso_line_radii(radius, NULL, 0);
    int32_t *dxs = (int32_t *) malloc((size_t) (2 * k + 1) * sizeof(int32_t));
    cdso_line_radii(radius, dxs, 2 * k + 1);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            for (int c = 0; c < nch; c++) {
                int m = 0;
                for (int dy = -k; dy <= k; dy++) {
                    int yy = y + dy;
                    if (yy < 0 || yy >= H) continue;
                    int dx = dxs[dy + k];
                    for (int xx = x - dx; xx <= x + dx; xx++) {
                        if (xx < 0 || xx >= W) continue;
                        int v = src[((size_t) yy * W + xx) * nch + c];
                        if (v > m) m = v;
                    }
                }
                dst[((size_t) y * W + x) * nch + c] = (uint8_t) m;
            }
    free(dxs);
}

/* running max of one row/channel with window [x-dx, x+dx] (van Herk / Gil-Werman), zero outside */
static void cdso_row_max(const uint8_t *row, int W, int stride, int dx, uint8_t *out, uint8_t *pre, uint8_t *suf)
{
    int win = 2 * dx + 1;
    /* padded coordinates: p = x + dx, padded length L = W + 2 dx; value 0 in the pads */
    int L = W + 2 * dx;
    for (int p = 0; p < L; p++) {
        int x = p - dx;
        uint8_t v = (x >= 0 && x < W) ? row[(size_t) x * stride] : 0;
        pre[p] = (p % win == 0) ? v : (pre[p - 1] > v ? pre[p - 1] : v);
    }
    for (int p = L - 1; p >= 0; p--) {
        int x = p - dx;
        uint8_t v = (x >= 0 && x < W) ? row[(size_t) x * stride] : 0;
        suf[p] = (p % win == win - 1 || p == L - 1) ? v : (suf[p + 1] > v ? suf[p + 1] : v);
    }
    for (int x = 0; x < W; x++) {
        /* window in padded coords: [x, x + 2dx] */
        uint8_t a = suf[x], b = pre[x + 2 * dx];
        out[x] = a > b ? a : b;
    }
}

CDSO_API void cdso_max_filter(const uint8_t *src, int W, int H, int nch, double radius, uint8_t *dst)
{
    int k = cdso_line_radii(radius, NULL, 0);
    int32_t *dxs = (int32_t *) malloc((size_t) (2 * k + 1) * sizeof(int32_t));
    cdso_line_radii(radius, dxs, 2 * k + 1);
    memset(dst, 0, (size_t) W * H * nch);
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        uint8_t *rowmax = (uint8_t *) malloc((size_t) W);
        uint8_t *pre = (uint8_t *) malloc((size_t) W + 2 * (size_t) k + 2);
        uint8_t *suf = (uint8_t *) malloc((size_t) W + 2 * (size_t) k + 2);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int y = 0; y < H; y++) {
            for (int dy = -k; dy <= k; dy++) {
                int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
                for (int c = 0; c < nch; c++) {
                    cdso_row_max(src + (size_t) yy * W * nch + c, W, nch, dxs[dy + k], rowmax, pre, suf);
                    uint8_t *d = dst + (size_t) y * W * nch + c;
                    for (int x = 0; x < W; x++)
                        if (rowmax[x] > d[(size_t) x * nch]) d[(size_t) x * nch] = rowmax[x];
                }
            }
        }
        free(rowmax); free(pre); free(suf);
    }
    free(dxs);
}

/* ------------------------------------------------------------------------------------------
 * a7: per-mask preparation of the shape score
 * API/cds/ColorDepthSearchAlgorithmProviderFactory.java:76-127
 *   Q  = query with label regions cleared (:96,103)
 *   HE = signal0(gray( max20(Q) != 0 ? black : max60(Q) ))   (:105-109)
 *   QM = signal2(gray(Q))                                     (:111)
 * border: LImage.toImageArray() only visits pixels inside the borders (API/imageprocessing/LImage.java:89-151),
 * so QM / HE stay 0 in the border frame.  Only border = 0 is in scope for the GPU path;
 * the oracle restates the border frame but NOT the sliding-window start quirk for border > 0.
 * Outputs: q_out uint8[H][W][3], qm_out uint8[H][W] in {0,1}, he_out uint8[H][W] in {0,1}.
 * ------------------------------------------------------------------------------------------ */
CDSO_API void cdso_shape_prepare_mask(const uint8_t *rgb, int W, int H, int border,
                                      const int32_t *rects, int nrects,
                                      uint8_t *q_out, uint8_t *qm_out, uint8_t *he_out)
{
    size_t n = (size_t) W * H;
    cdso_clear_regions(rgb, W, H, rects, nrects, q_out);
    uint8_t *m60 = (uint8_t *) malloc(n * 3);
    uint8_t *m20 = (uint8_t *) malloc(n * 3);
    cdso_max_filter(q_out, W, H, 3, 60, m60);
    cdso_max_filter(q_out, W, H, 3, 20, m20);
    memset(qm_out, 0, n);
    memset(he_out, 0, n);
    for (int y = border; y < H - border; y++)
        for (int x = border; x < W - border; x++) {
            size_t i = (size_t) y * W + x;
            int p2nz = m20[3*i] | m20[3*i+1] | m20[3*i+2];
            int g = p2nz ? 0 : cdso_rgb_to_gray(m60[3*i], m60[3*i+1], m60[3*i+2]);
            he_out[i] = g > 0 ? 1 : 0;
            qm_out[i] = cdso_rgb_to_gray(q_out[3*i], q_out[3*i+1], q_out[3*i+2]) > 2 ? 1 : 0;
        }
    free(m60); free(m20);
}

/* PIXEL_GAP_OP  API/cds/Shape2DMatchColorDepthSearchAlgorithm.java:26-42 */
static inline int cdso_pixel_gap_op(int queryPix, int queryMask, int targetGradPix, int targetDilatedPix)
{
    int gap;
    if ((queryPix & 0xFFFFFF) != 0 && (targetDilatedPix & 0xFFFFFF) != 0) {
        int pxGapSlice = cdso_slice_gap(queryPix, targetDilatedPix);
        if (40 <= pxGapSlice - 40) gap = pxGapSlice - 40;
        else gap = queryMask * targetGradPix;
    } else {
        gap = queryMask * targetGradPix;
    }
    return gap > 3 ? gap : 0;
}

static inline int cdso_rgb_int(const uint8_t *p) { return (int) 0xFF000000u | (p[0] << 16) | (p[1] << 8) | p[2]; }

/* ColorTransformation.mask(pt, p, m) :134-143 used for the ROI: black/0 where the ROI pixel is black */
static inline int cdso_roi_keep(const uint8_t *roi, size_t i) { return roi == NULL || (roi[3*i] | roi[3*i+1] | roi[3*i+2]) != 0; }

/* a10: calculateNegativeScores :196-245 for one orientation.
 * q/qm/he: prepared mask planes.  roi: label-cleared ROI RGB or NULL (not mirrored, :205-218).
 * t: label-cleared target RGB.  grad: gray16 (or gray8 widened) [H][W].  z: zgap RGB with mask(threshold) applied. */
static void cdso_negative_scores(const uint8_t *q, const uint8_t *qm, const uint8_t *he, const uint8_t *roi,
                                 const uint8_t *t, const uint16_t *grad, const uint8_t *z,
                                 int W, int H, int border, int query_threshold, int mirror,
                                 int64_t *gap_out, int64_t *he_out)
{
    int64_t gapsum = 0, hesum = 0;
    for (int y = border; y < H - border; y++)
        for (int x = border; x < W - border; x++) {
            size_t i = (size_t) y * W + x;
            int sx = mirror ? W - x - 1 : x;                         /* horizontalMirror :158-165 */
            size_t si = (size_t) y * W + sx;
            int keep = cdso_roi_keep(roi, i);
            int queryPix = keep ? cdso_rgb_int(q + 3 * si) : (int) 0xFF000000u;
            int queryMask = keep ? qm[si] : 0;
            int heMask = keep ? he[si] : 0;
            int targetGradPix = grad[i];                             /* gradient is NOT mirrored :222 */
            int targetDilatedPix = cdso_rgb_int(z + 3 * si);         /* zgap IS mirrored :223 */
            gapsum += cdso_pixel_gap_op(queryPix, queryMask, targetGradPix, targetDilatedPix);
            if (heMask == 1 && cdso_rgb_passes(t + 3 * i, query_threshold)) hesum += 1;   /* :226-239 */
        }
    *gap_out = gapsum;
    *he_out = hesum;
}

/* a10: Shape2DMatchColorDepthSearchAlgorithm.calculateMatchingScore :150-186.
 * target/zgap are the RAW images; label clearing of the target (:159) and mask(threshold) of the zgap (:161)
 * happen here.  grad may be NULL or zgap may be NULL -> (-1,-1) (:155-158).
 * Returns 0; outputs gap, high expression area, mirrored flag. */
CDSO_API int cdso_shape_score(const uint8_t *q, const uint8_t *qm, const uint8_t *he, const uint8_t *roi,
                              int W, int H, int border, int query_threshold, int mirror_query,
                              const int32_t *rects, int nrects,
                              const uint8_t *target, const uint16_t *grad, const uint8_t *zgap,
                              int64_t *gap_out, int64_t *he_out, int32_t *mirrored_out)
{
    if (grad == NULL || zgap == NULL) { *gap_out = -1; *he_out = -1; *mirrored_out = 0; return 0; }
    size_t n = (size_t) W * H;
    uint8_t *t = (uint8_t *) malloc(n * 3);
    uint8_t *z = (uint8_t *) malloc(n * 3);
    cdso_clear_regions(target, W, H, rects, nrects, t);
    cdso_mask_rgb(zgap, W, H, query_threshold, z);
    int64_t g0, h0;
    cdso_negative_scores(q, qm, he, roi, t, grad, z, W, H, border, query_threshold, 0, &g0, &h0);
    *gap_out = g0; *he_out = h0; *mirrored_out = 0;
    if (mirror_query) {
        int64_t g1, h1;
        cdso_negative_scores(q, qm, he, roi, t, grad, z, W, H, border, query_threshold, 1, &g1, &h1);
        /* ShapeMatchScore.getScore narrows to int, API/cds/ShapeMatchScore.java:29-33 */
        int s0 = (int) cdso_shape_score_2d(g0, h0);
        int s1 = (int) cdso_shape_score_2d(g1, h1);
        if (s1 < s0) { *gap_out = g1; *he_out = h1; *mirrored_out = 1; }        /* strict :181 */
    }
    free(t); free(z);
    return 0;
}

/* The zgap image the reference tests synthesise when no file exists
 * (colormipsearch-api/src/test/java/.../cds/Shape2DMatchColorDepthSearchAlgorithmTest.java:171-174):
 * clear labels -> mask(threshold) -> unsafeMaxFilter(10). */
CDSO_API void cdso_make_zgap(const uint8_t *target, int W, int H, int threshold,
                             const int32_t *rects, int nrects, uint8_t *zgap_out)
{
    size_t n = (size_t) W * H;
    uint8_t *a = (uint8_t *) malloc(n * 3);
    uint8_t *b = (uint8_t *) malloc(n * 3);
    cdso_clear_regions(target, W, H, rects, nrects, a);
    cdso_mask_rgb(a, W, H, threshold, b);
    cdso_max_filter(b, W, H, 3, 10, zgap_out);
    free(a); free(b);
}

CDSO_API int cdso_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
