// cds_shape.cu -- 2D shape (gradient area gap) scoring on the device.
//
// Reference semantics (API/ = colormipsearch-api/src/main/java/org/janelia/colormipsearch/):
//   per-mask preparation : ColorDepthSearchAlgorithmProviderFactory.createShapeMatchCDSAlgorithmProvider, API/cds/...ProviderFactory.java:76-127
//   max filter           : ImageTransformation.unsafeMaxFilter / makeLineRadii, API/imageprocessing/ImageTransformation.java:353-572
//   gray / signal / mask : API/imageprocessing/ColorTransformation.java:29-54, 97-160
//   slice numbers        : GradientAreaGapUtils.calculateSliceGap / findSliceNumber(InLUT), API/cds/GradientAreaGapUtils.java:18-197
//   pair score           : Shape2DMatchColorDepthSearchAlgorithm.calculateMatchingScore / calculateNegativeScores / PIXEL_GAP_OP,
//                          API/cds/Shape2DMatchColorDepthSearchAlgorithm.java:26-42, 150-245
//
// B200 formulation.  PIXEL_GAP_OP is zero wherever the (label-cleared, ROI-masked) query pixel is black -- its "queryMask * gradient"
// term needs gray(query) > 2 -- so the gradient-area gap is a sum over the query's NON-BLACK pixels only (~1-3 % of the image).
// Per mask the device keeps (a) that pixel list with the query's slice number folded in, (b) the high-expression mask and its
// row-mirrored copy as bitmaps.  Per target it keeps the slice-number plane of the thresholded zgap image, the gradient plane and
// the "above threshold outside labels" bitmap.  One CTA scores one (mask, target) pair in both orientations: a gather over the
// pixel list for the gaps, and an AND + POPC sweep over the bitmaps for the high-expression area.
#include <algorithm>
#include <cmath>
#include <vector>

#include "cds_lut.h"
#include "cds_runtime.h"
#include "cds_tiff.h"

using namespace cds;

namespace cds {

// ------------------------------------------------------------------------------------------------------------------ slice numbers
__constant__ uint8_t c_shape_lut[256 * 3];
static bool g_shape_lut_uploaded[64] = {false};

static cudaError_t ensure_shape_lut(int dev)
{
    if (dev < 64 && g_shape_lut_uploaded[dev]) return cudaSuccess;
    cudaError_t e = cudaMemcpyToSymbol(c_shape_lut, kColorDepthLut, sizeof(kColorDepthLut));
    if (e == cudaSuccess && dev < 64) g_shape_lut_uploaded[dev] = true;
    return e;
}

// findSliceNumberInLUT, API/cds/GradientAreaGapUtils.java:131-197 (IEEE double, no contraction: the file is built with --fmad=false)
__device__ int find_slice_in_lut(int lo, int hi, double colorRatio)
{
    int sliceNumber = 0;
    double mingapratio = 1000;
    for (int i = lo; i <= hi; i++) {
        const double colorR = c_shape_lut[3 * i], colorG = c_shape_lut[3 * i + 1], colorB = c_shape_lut[3 * i + 2];
        double lutRatio = 0;
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
        if (lutRatio == colorRatio) return i + 1;
        const double gapratio = fabs(colorRatio - lutRatio);
        if (gapratio < mingapratio) { mingapratio = gapratio; sliceNumber = i + 1; }
    }
    return sliceNumber;
}

// the slice of one colour: first half of calculateSliceGap (:18-99) + findSliceNumber (:107-129).  0 for black.
__device__ int slice_number(int red, int green, int blue)
{
    if ((red | green | blue) == 0) return 0;
    int max1, max2, lo, hi;
    if (red >= green && red >= blue) {
        max1 = red;
        if (green >= blue) { max2 = green; lo = 171; hi = 212; } else { max2 = blue; lo = 213; hi = 255; }
    } else if (green >= red && green >= blue) {
        max1 = green;
        if (red >= blue) { max2 = red; lo = 128; hi = 170; } else { max2 = blue; lo = 86; hi = 127; }
    } else {
        max1 = blue;
        if (red >= green) { max2 = red; lo = 0; hi = 29; } else { max2 = green; lo = 30; hi = 85; }
    }
    return find_slice_in_lut(lo, hi, (double) max2 / (double) max1);
}

__device__ __forceinline__ bool shape_in_rects(const RectSet &r, int x, int y)
{
    bool in = false;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (i < r.n) in |= (x >= r.x0[i] && x < r.x1[i] && y >= r.y0[i] && y < r.y1[i]);
    return in;
}

// rgbToGrayNoGammaCorrection with maxGrayValue 255 (ColorTransformation.java:40-54, :103) in its exact integer form
// floor((2 (r + g + b) + 3) / 6): the double expression is never within 1/6 of an integer boundary (checked over all 2^24 colours
// against the oracle, tests/test_oracle_golden.py::test_gray_integer_form)
__device__ __forceinline__ int gray_of(int r, int g, int b) { return (r | g | b) == 0 ? 0 : (2 * (r + g + b) + 3) / 6; }

// ------------------------------------------------------------------------------------------------------------------ max filter
// Disc dilation with ImageJ's RankFilters disc (makeLineRadii): dst(x,y,c) = max over dy in [-k,k], |dx| <= dxs[dy+k] of
// src(x+dx, y+dy, c), pixels outside the image ignored.  One CTA produces a TR x TC tile of one channel: the tile plus a k-wide
// halo is loaded into shared memory, a sparse table of horizontal running maxima over windows 1,2,4,.. is built next to it, and
// every disc row then costs two table reads.  Tiles whose halo is entirely zero (most of a colour-depth MIP) are written as zeros.
constexpr int kMfTR = 16, kMfTC = 64;

struct DiscSpec {
    int k;
    int dxs[121];      // half widths for dy = -k..k (k <= 60)
};

__global__ void __launch_bounds__(256) max_filter_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int W, int H, int nch,
                                                         DiscSpec disc, int nlev)
{
    extern __shared__ uint8_t s_lev[];               // [nlev][rows][cols]
    __shared__ int s_row_lev[121], s_row_off[121];
    const int k = disc.k;
    const int rows = kMfTR + 2 * k, cols = kMfTC + 2 * k;
    const int plane = rows * cols;
    const int x0 = blockIdx.x * kMfTC, y0 = blockIdx.y * kMfTR;
    const int img = blockIdx.z / nch, ch = blockIdx.z % nch;
    const uint8_t *s = src + (size_t) img * W * H * nch + ch;
    uint8_t *d = dst + (size_t) img * W * H * nch + ch;

    int nz = 0;
    for (int i = threadIdx.x; i < plane; i += blockDim.x) {
        const int r = i / cols, c = i % cols;
        const int x = x0 - k + c, y = y0 - k + r;
        uint8_t v = 0;
        if (x >= 0 && x < W && y >= 0 && y < H) v = s[((size_t) y * W + x) * nch];
        s_lev[i] = v;
        nz |= v;
    }
    for (int i = threadIdx.x; i < 2 * k + 1; i += blockDim.x) {
        const int L = 2 * disc.dxs[i] + 1;
        int j = 0;
        while ((2 << j) <= L) j++;
        s_row_lev[i] = j;
        s_row_off[i] = L - (1 << j);
    }
    const int any = __syncthreads_or(nz);
    if (!any) {
        for (int i = threadIdx.x; i < kMfTR * kMfTC; i += blockDim.x) {
            const int x = x0 + i % kMfTC, y = y0 + i / kMfTC;
            if (x < W && y < H) d[((size_t) y * W + x) * nch] = 0;
        }
        return;
    }
    for (int j = 1; j < nlev; j++) {
        const uint8_t *prev = s_lev + (size_t) (j - 1) * plane;
        uint8_t *cur = s_lev + (size_t) j * plane;
        const int step = 1 << (j - 1);
        for (int i = threadIdx.x; i < plane; i += blockDim.x) {
            const int c = i % cols;
            uint8_t v = prev[i];
            if (c + step < cols) v = max(v, prev[i + step]);
            cur[i] = v;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < kMfTR * kMfTC; i += blockDim.x) {
        const int ox = i % kMfTC, oy = i / kMfTC;
        const int x = x0 + ox, y = y0 + oy;
        if (x >= W || y >= H) continue;
        int m = 0;
        for (int dy = -k; dy <= k; dy++) {
            const int w = disc.dxs[dy + k];
            const uint8_t *row = s_lev + (size_t) s_row_lev[dy + k] * plane + (size_t) (oy + k + dy) * cols + (ox + k - w);
            m = max(m, max((int) row[0], (int) row[s_row_off[dy + k]]));
        }
        d[((size_t) y * W + x) * nch] = (uint8_t) m;
    }
}

static DiscSpec make_disc(double radiusArg)
{
    // makeLineRadii, API/imageprocessing/ImageTransformation.java:549-572
    DiscSpec d{};
    double radius;
    if (radiusArg >= 1.5 && radiusArg < 1.75) radius = 1.75;
    else if (radiusArg >= 2.5 && radiusArg < 2.85) radius = 2.85;
    else radius = radiusArg;
    const int r2 = (int) (radius * radius) + 1;
    const int k = (int) (std::sqrt(r2 + 1e-10));
    d.k = k;
    if (k > 60) return d;
    for (int y = -k; y <= k; y++) d.dxs[y + k] = (y == 0) ? k : (int) (std::sqrt(r2 - y * y + 1e-10));
    return d;
}

// dilates n images of nch interleaved channels; returns false when the radius is not supported
static bool launch_max_filter(const uint8_t *src, uint8_t *dst, int64_t n, int W, int H, int nch, double radius, cudaStream_t s)
{
    DiscSpec disc = make_disc(radius);
    if (disc.k > 60 || disc.k < 0) return false;
    int nlev = 1;
    while ((1 << nlev) <= 2 * disc.k + 1) nlev++;
    const size_t smem = (size_t) nlev * (kMfTR + 2 * disc.k) * (kMfTC + 2 * disc.k);
    cudaFuncSetAttribute(max_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    for (int64_t i0 = 0; i0 < n; i0 += 8192) {
        const int64_t cnt = std::min<int64_t>(8192, n - i0);
        dim3 grid((W + kMfTC - 1) / kMfTC, (H + kMfTR - 1) / kMfTR, (unsigned) (cnt * nch));
        max_filter_kernel<<<grid, 256, smem, s>>>(src + (size_t) i0 * W * H * nch, dst + (size_t) i0 * W * H * nch, W, H, nch, disc, nlev);
    }
    return true;
}

// ------------------------------------------------------------------------------------------------------------------ image transforms
// dst = clearRegion(src) then optionally mask(threshold): black inside label rectangles, black where every channel <= threshold
__global__ void clear_and_mask_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int64_t n, int W, int H, RectSet rects,
                                      int threshold, int apply_mask)
{
    const size_t total = (size_t) n * W * H;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        const int x = (int) (i % W), y = (int) ((i / W) % H);
        int r = src[3 * i], g = src[3 * i + 1], b = src[3 * i + 2];
        bool keep = !shape_in_rects(rects, x, y);
        if (apply_mask) keep = keep && (r > threshold || g > threshold || b > threshold);
        if (!keep) r = g = b = 0;
        dst[3 * i] = (uint8_t) r; dst[3 * i + 1] = (uint8_t) g; dst[3 * i + 2] = (uint8_t) b;
    }
}

// target side of a pair: zslice = slice number of mask(threshold)(zgap) (0 = black), tsig bit = clearLabels(target) above threshold
__global__ void __launch_bounds__(256) target_planes_kernel(const uint8_t *__restrict__ target, const uint8_t *__restrict__ zgap, int W, int H,
                                                            RectSet rects, int threshold, int bpitch, uint16_t *__restrict__ zslice,
                                                            uint32_t *__restrict__ tsig)
{
    const int y = blockIdx.x;
    const int64_t img = blockIdx.y;
    const uint8_t *trow = target + ((size_t) img * H + y) * W * 3;
    const uint8_t *zrow = zgap + ((size_t) img * H + y) * W * 3;
    uint16_t *zs = zslice + ((size_t) img * H + y) * W;
    uint32_t *ts = tsig + ((size_t) img * H + y) * bpitch;
    const int lane = threadIdx.x & 31;
    for (int x0 = (threadIdx.x >> 5) * 32; x0 < bpitch * 32; x0 += (int) blockDim.x) {
        const int x = x0 + lane;
        bool sig = false;
        if (x < W) {
            const int r = trow[3 * x], g = trow[3 * x + 1], b = trow[3 * x + 2];
            sig = !shape_in_rects(rects, x, y) && (r > threshold || g > threshold || b > threshold);
            const int zr = zrow[3 * x], zg = zrow[3 * x + 1], zb = zrow[3 * x + 2];
            int sl = 0;
            if (zr > threshold || zg > threshold || zb > threshold) sl = slice_number(zr, zg, zb);   // mask(threshold) then slice
            zs[x] = (uint16_t) sl;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, sig);
        if (lane == 0) ts[x0 >> 5] = bal;
    }
}

// mask side: from Q (label-cleared query), max60(Q), max20(Q) and the label-cleared ROI (or NULL):
//   QM = gray(Q) > 2, HE = max20 == black && gray(max60) > 0            (ProviderFactory :105-111)
//   gap list entry for every pixel with Q != black that the ROI keeps in at least one orientation:
//       x | y << 11 | (slice(Q) - 1) << 21 | QM << 29 | keep_normal << 30 | keep_mirrored << 31
//   bitmaps he_n = HE & ROI, he_m = mirror(HE) & ROI;  counters[0] += sum QM, counters[1] += sum HE, counters[2] = list length
__global__ void __launch_bounds__(256) mask_planes_kernel(const uint8_t *__restrict__ q, const uint8_t *__restrict__ m60, const uint8_t *__restrict__ m20,
                                                          const uint8_t *__restrict__ roi, int W, int H, int bpitch,
                                                          uint32_t *__restrict__ he_n, uint32_t *__restrict__ he_m,
                                                          uint32_t *__restrict__ gap_list, unsigned long long *__restrict__ counters)
{
    const int y = blockIdx.x;
    const uint8_t *qrow = q + (size_t) y * W * 3;
    const uint8_t *arow = m60 + (size_t) y * W * 3;
    const uint8_t *brow = m20 + (size_t) y * W * 3;
    const uint8_t *rrow = roi ? roi + (size_t) y * W * 3 : nullptr;
    const int lane = threadIdx.x & 31;
    int qm_cnt = 0, he_cnt = 0;
    for (int x0 = (threadIdx.x >> 5) * 32; x0 < bpitch * 32; x0 += (int) blockDim.x) {
        const int x = x0 + lane;          // normal orientation: output pixel x takes HE(x)
        const int xs = W - 1 - x;         // mirrored orientation: output pixel x takes HE(W-1-x)
        bool hn = false, hm = false;
        if (x < W) {
            const bool keep = !rrow || (rrow[3 * x] | rrow[3 * x + 1] | rrow[3 * x + 2]) != 0;
            const bool he_here = (brow[3 * x] | brow[3 * x + 1] | brow[3 * x + 2]) == 0 && gray_of(arow[3 * x], arow[3 * x + 1], arow[3 * x + 2]) > 0;
            const bool he_mir = (brow[3 * xs] | brow[3 * xs + 1] | brow[3 * xs + 2]) == 0 && gray_of(arow[3 * xs], arow[3 * xs + 1], arow[3 * xs + 2]) > 0;
            hn = he_here && keep;
            hm = he_mir && keep;
            he_cnt += he_here ? 1 : 0;
            const int r = qrow[3 * x], g = qrow[3 * x + 1], b = qrow[3 * x + 2];
            if ((r | g | b) != 0) {
                const int qm = gray_of(r, g, b) > 2 ? 1 : 0;
                qm_cnt += qm;
                // this query pixel lands on output x (normal) and on output W-1-x (mirrored); the ROI is tested at the output
                const bool keep_m = !rrow || (rrow[3 * xs] | rrow[3 * xs + 1] | rrow[3 * xs + 2]) != 0;
                if (keep || keep_m) {
                    const uint32_t e = (uint32_t) x | ((uint32_t) y << 11) | ((uint32_t) (slice_number(r, g, b) - 1) << 21) |
                                       ((uint32_t) qm << 29) | ((uint32_t) keep << 30) | ((uint32_t) keep_m << 31);
                    const unsigned long long slot = atomicAdd(&counters[2], 1ull);
                    gap_list[slot] = e;
                }
            }
        }
        const unsigned bn = __ballot_sync(0xffffffffu, hn), bm = __ballot_sync(0xffffffffu, hm);
        if (lane == 0) { he_n[(size_t) y * bpitch + (x0 >> 5)] = bn; he_m[(size_t) y * bpitch + (x0 >> 5)] = bm; }
    }
    qm_cnt = __reduce_add_sync(0xffffffffu, qm_cnt);
    he_cnt = __reduce_add_sync(0xffffffffu, he_cnt);
    if (lane == 0) {
        if (qm_cnt) atomicAdd(&counters[0], (unsigned long long) qm_cnt);
        if (he_cnt) atomicAdd(&counters[1], (unsigned long long) he_cnt);
    }
}

// ------------------------------------------------------------------------------------------------------------------ pair kernel
struct ShapeMaskDesc {
    const uint32_t *gap_list;
    const uint32_t *he_n, *he_m;
    int n_gap;
    int pad;
};

__global__ void __launch_bounds__(256) shape_pair_kernel(const ShapeMaskDesc *__restrict__ masks, const int32_t *__restrict__ pair_mask,
                                                         const int64_t *__restrict__ pair_target, const uint8_t *__restrict__ has_variants,
                                                         const uint16_t *__restrict__ zslice, const uint16_t *__restrict__ grad,
                                                         const uint32_t *__restrict__ tsig, int W, int H, int bpitch, int mirror,
                                                         long long *__restrict__ gap_out, long long *__restrict__ he_out, uint8_t *__restrict__ mir_out)
{
    __shared__ long long s_red[4][8];
    const int64_t pr = blockIdx.x;
    const int64_t t = pair_target[pr];
    if (has_variants && !has_variants[t]) {                      // missing gradient / zgap supplier: (-1, -1, not mirrored), Shape2DMatch...:155-158
        if (threadIdx.x == 0) { gap_out[pr] = -1; he_out[pr] = -1; mir_out[pr] = 0; }
        return;
    }
    const ShapeMaskDesc md = masks[pair_mask[pr]];
    const uint16_t *zs = zslice + (size_t) t * W * H;
    const uint16_t *gr = grad + (size_t) t * W * H;
    const uint32_t *ts = tsig + (size_t) t * H * bpitch;
    long long gap_n = 0, gap_m = 0, he_n = 0, he_m = 0;
    for (int i = threadIdx.x; i < md.n_gap; i += blockDim.x) {
        const uint32_t e = __ldg(md.gap_list + i);
        const int x = (int) (e & 0x7FFu), y = (int) ((e >> 11) & 0x3FFu);
        const int qs = (int) ((e >> 21) & 0xFFu) + 1;
        const int qm = (int) ((e >> 29) & 1u);
        const size_t rowoff = (size_t) y * W;
        const int z = zs[rowoff + x];                            // the zgap image is sampled at the query's own pixel in both orientations
        int big = 0;
        if (z != 0) {                                            // PIXEL_GAP_OP :30-37
            const int pxGapSlice = abs(qs - z);
            if (40 <= pxGapSlice - 40) big = pxGapSlice - 40;
        }
        if (e & 0x40000000u) {
            const int gap = big ? big : qm * (int) gr[rowoff + x];
            if (gap > 3) gap_n += gap;
        }
        if (mirror && (e & 0x80000000u)) {
            const int gap = big ? big : qm * (int) gr[rowoff + (W - 1 - x)];   // the gradient is NOT mirrored with the query :222
            if (gap > 3) gap_m += gap;
        }
    }
    const int n_words = H * bpitch;
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) {
        const uint32_t tw = ts[i];
        he_n += __popc(__ldg(md.he_n + i) & tw);
        if (mirror) he_m += __popc(__ldg(md.he_m + i) & tw);
    }
    long long v[4] = {gap_n, gap_m, he_n, he_m};
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long tot[4] = {0, 0, 0, 0};
        for (int k = 0; k < 4; k++) for (int w = 0; w < 8; w++) tot[k] += s_red[k][w];
        long long g = tot[0], h = tot[2];
        int mir = 0;
        if (mirror) {
            // ShapeMatchScore.getScore narrows calculate2DShapeScore to int; mirrored wins when strictly smaller (:181)
            const int s0 = (int) (tot[0] + tot[2] / 3);
            const int s1 = (int) (tot[1] + tot[3] / 3);
            if (s1 < s0) { g = tot[1]; h = tot[3]; mir = 1; }
        }
        gap_out[pr] = g; he_out[pr] = h; mir_out[pr] = (uint8_t) mir;
    }
}

}  // namespace cds

// ------------------------------------------------------------------------------------------------------------------ C ABI
struct cds_shape_maskset {
    cds_ctx *ctx = nullptr;
    int W = 0, H = 0, bpitch = 0;
    int query_threshold = 0, mirror = 0;
    RectSet rects{};
    uint8_t *d_roi = nullptr;                        // label-cleared ROI on device 0, or nullptr
    struct Mask { uint32_t *gap_list = nullptr; uint32_t *he_n = nullptr; uint32_t *he_m = nullptr; int n_gap = 0; };
    std::vector<Mask> masks;
    ShapeMaskDesc *d_descs = nullptr;
    bool descs_dirty = true;
};

#define SH_TRY(expr) do { cds_status _s = (expr); if (_s != CDS_OK) return _s; } while (0)
#define SH_CUDA(ctx, expr) SH_TRY((ctx)->check((expr), #expr))

static RectSet to_rectset(const cds_rect *rects, int n)
{
    RectSet r{};
    r.n = n;
    for (int i = 0; i < n; i++) { r.x0[i] = rects[i].x0; r.y0[i] = rects[i].y0; r.x1[i] = rects[i].x1; r.y1[i] = rects[i].y1; }
    return r;
}

extern "C" cds_status cds_shape_maskset_create(cds_ctx *ctx, int32_t width, int32_t height, int32_t query_threshold, int32_t border,
                                               int32_t mirror, const cds_rect *rects, int32_t n_rects, const uint8_t *roi_rgb,
                                               cds_shape_maskset **out)
{
    if (!ctx || !out) { set_tls_error("cds_shape_maskset_create: NULL argument"); return CDS_ERR_BAD_ARG; }
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    *out = nullptr;
    if (width <= 0 || height <= 0 || n_rects < 0 || n_rects > CDS_MAX_RECTS || (n_rects > 0 && !rects))
        return ctx->fail(CDS_ERR_BAD_ARG, "cds_shape_maskset_create: bad arguments");
    if (width > 2048 || height > 1024) return ctx->fail(CDS_ERR_UNSUPPORTED, "cds_shape_maskset_create: images larger than 2048 x 1024 are not supported");
    if (border != 0) return ctx->fail(CDS_ERR_UNSUPPORTED, "cds_shape_maskset_create: only border = 0 is supported");
    auto sms = new cds_shape_maskset();
    sms->ctx = ctx; sms->W = width; sms->H = height; sms->bpitch = occupancy_valid_pitch(width);   // row-layout bitmaps: one bit per pixel, 32-pixel words
    sms->query_threshold = query_threshold; sms->mirror = mirror ? 1 : 0;
    sms->rects = to_rectset(rects, n_rects);
    DevState &d0 = ctx->devs[0];
    cds_status st = ctx->check(cudaSetDevice(d0.dev), "cudaSetDevice");
    if (st == CDS_OK) st = ctx->check(ensure_shape_lut(d0.dev), "lut upload");
    if (st == CDS_OK && roi_rgb) {
        const size_t bytes = (size_t) width * height * 3;
        uint8_t *tmp = nullptr;
        st = ctx->check(cudaMalloc(&sms->d_roi, bytes), "cudaMalloc(roi)");
        if (st == CDS_OK) st = ctx->check(cudaMalloc(&tmp, bytes), "cudaMalloc(roi tmp)");
        if (st == CDS_OK) st = ctx->check(cudaMemcpyAsync(tmp, roi_rgb, bytes, cudaMemcpyHostToDevice, d0.stream), "roi H2D");
        if (st == CDS_OK) {
            clear_and_mask_kernel<<<148 * 4, 256, 0, d0.stream>>>(tmp, sms->d_roi, 1, width, height, sms->rects, 0, 0);   // ROI is label-cleared too (:97-101)
            st = ctx->check(cudaGetLastError(), "clear_and_mask_kernel");
        }
        if (st == CDS_OK) st = ctx->check(cudaStreamSynchronize(d0.stream), "roi");
        if (tmp) cudaFree(tmp);
    }
    if (st != CDS_OK) { cds_shape_maskset_destroy(sms); return st; }
    *out = sms;
    return CDS_OK;
}

extern "C" void cds_shape_maskset_destroy(cds_shape_maskset *sms)
{
    if (!sms) return;
    cds_ctx *ctx = sms->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    cudaSetDevice(ctx->devs[0].dev);
    cudaStreamSynchronize(ctx->devs[0].stream);
    DevPool &pool = ctx->devs[0].pool;
    for (auto &m : sms->masks) {
        pool.free(m.gap_list);
        pool.free(m.he_n);
        pool.free(m.he_m);
    }
    if (sms->d_roi) cudaFree(sms->d_roi);
    if (sms->d_descs) cudaFree(sms->d_descs);
    cudaGetLastError();
    delete sms;
}

extern "C" int32_t cds_shape_maskset_size(const cds_shape_maskset *sms) { return sms ? (int32_t) sms->masks.size() : 0; }

extern "C" cds_status cds_shape_maskset_add_rgb(cds_shape_maskset *sms, const uint8_t *rgb, int32_t n, int64_t *qm_size_out, int64_t *he_size_out)
{
    if (!sms) { set_tls_error("cds_shape_maskset_add_rgb: NULL mask set"); return CDS_ERR_BAD_ARG; }
    cds_ctx *ctx = sms->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (n < 0 || (n > 0 && !rgb)) return ctx->fail(CDS_ERR_BAD_ARG, "cds_shape_maskset_add_rgb: bad arguments");
    DevState &d0 = ctx->devs[0];
    SH_CUDA(ctx, cudaSetDevice(d0.dev));
    const int W = sms->W, H = sms->H;
    const size_t px = (size_t) W * H, bytes = px * 3;
    const size_t bm_words = (size_t) H * sms->bpitch;
    uint8_t *d_raw = nullptr, *d_q = nullptr, *d_m60 = nullptr, *d_m20 = nullptr;
    uint32_t *d_list = nullptr;
    unsigned long long *d_cnt = nullptr;
    DevPool &pool = d0.pool;
    cds_status st = ctx->check(pool.alloc((void **) &d_raw, bytes), "cudaMalloc");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_q, bytes), "cudaMalloc");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_m60, bytes), "cudaMalloc");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_m20, bytes), "cudaMalloc");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_list, px * sizeof(uint32_t)), "cudaMalloc");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_cnt, 4 * sizeof(unsigned long long)), "cudaMalloc");
    for (int i = 0; i < n && st == CDS_OK; i++) {
        cds_shape_maskset::Mask m;
        st = ctx->check(pool.alloc((void **) &m.he_n, bm_words * sizeof(uint32_t)), "cudaMalloc(he bitmap)");
        if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &m.he_m, bm_words * sizeof(uint32_t)), "cudaMalloc(he bitmap)");
        if (st == CDS_OK) st = ctx->check(cudaMemcpyAsync(d_raw, rgb + (size_t) i * bytes, bytes, cudaMemcpyHostToDevice, d0.stream), "mask H2D");
        if (st == CDS_OK) st = ctx->check(cudaMemsetAsync(d_cnt, 0, 4 * sizeof(unsigned long long), d0.stream), "memset");
        if (st == CDS_OK) {
            clear_and_mask_kernel<<<148 * 4, 256, 0, d0.stream>>>(d_raw, d_q, 1, W, H, sms->rects, 0, 0);
            launch_max_filter(d_q, d_m60, 1, W, H, 3, 60, d0.stream);
            launch_max_filter(d_q, d_m20, 1, W, H, 3, 20, d0.stream);
            mask_planes_kernel<<<H, 256, 0, d0.stream>>>(d_q, d_m60, d_m20, sms->d_roi, W, H, sms->bpitch, m.he_n, m.he_m, d_list, d_cnt);
            ctx->stats.kernel_launches += 4;
            st = ctx->check(cudaGetLastError(), "shape mask kernels");
        }
        unsigned long long cnt[4] = {0, 0, 0, 0};
        if (st == CDS_OK) st = ctx->check(cudaMemcpyAsync(cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, d0.stream), "counters D2H");
        if (st == CDS_OK) st = ctx->check(cudaStreamSynchronize(d0.stream), "shape mask");
        if (st == CDS_OK) {
            m.n_gap = (int) cnt[2];
            st = ctx->check(pool.alloc((void **) &m.gap_list, std::max<size_t>(1, (size_t) m.n_gap) * sizeof(uint32_t)), "cudaMalloc(gap list)");
            if (st == CDS_OK && m.n_gap) st = ctx->check(cudaMemcpyAsync(m.gap_list, d_list, (size_t) m.n_gap * sizeof(uint32_t), cudaMemcpyDeviceToDevice, d0.stream), "gap list copy");
            if (st == CDS_OK) st = ctx->check(cudaStreamSynchronize(d0.stream), "gap list");
        }
        if (st != CDS_OK) {
            pool.free(m.he_n);
            pool.free(m.he_m);
            pool.free(m.gap_list);
            break;
        }
        if (qm_size_out) qm_size_out[i] = (int64_t) cnt[0];
        if (he_size_out) he_size_out[i] = (int64_t) cnt[1];
        sms->masks.push_back(m);
        sms->descs_dirty = true;
    }
    cudaStreamSynchronize(d0.stream);
    pool.free(d_raw);
    pool.free(d_q);
    pool.free(d_m60);
    pool.free(d_m20);
    pool.free(d_list);
    pool.free(d_cnt);
    return st;
}

extern "C" cds_status cds_make_zgap(cds_ctx *ctx, const uint8_t *rgb, int64_t n, int32_t width, int32_t height, int32_t threshold,
                                    double radius, const cds_rect *rects, int32_t n_rects, uint8_t *zgap_out)
{
    if (!ctx) { set_tls_error("cds_make_zgap: NULL ctx"); return CDS_ERR_BAD_ARG; }
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (n < 0 || (n > 0 && (!rgb || !zgap_out)) || width <= 0 || height <= 0 || n_rects < 0 || n_rects > CDS_MAX_RECTS || (n_rects > 0 && !rects))
        return ctx->fail(CDS_ERR_BAD_ARG, "cds_make_zgap: bad arguments");
    if (make_disc(radius).k > 60) return ctx->fail(CDS_ERR_UNSUPPORTED, "cds_make_zgap: radius > 60 is not supported");
    DevState &d0 = ctx->devs[0];
    SH_CUDA(ctx, cudaSetDevice(d0.dev));
    const RectSet rs = to_rectset(rects, n_rects);
    const size_t bytes = (size_t) width * height * 3;
    const int64_t chunk = 32;
    uint8_t *d_a = nullptr, *d_b = nullptr;
    cds_status st = ctx->check(cudaMalloc(&d_a, chunk * bytes), "cudaMalloc");
    if (st == CDS_OK) st = ctx->check(cudaMalloc(&d_b, chunk * bytes), "cudaMalloc");
    for (int64_t i0 = 0; i0 < n && st == CDS_OK; i0 += chunk) {
        const int64_t cnt = std::min<int64_t>(chunk, n - i0);
        st = ctx->check(cudaMemcpyAsync(d_a, rgb + (size_t) i0 * bytes, (size_t) cnt * bytes, cudaMemcpyHostToDevice, d0.stream), "zgap H2D");
        if (st != CDS_OK) break;
        clear_and_mask_kernel<<<148 * 4, 256, 0, d0.stream>>>(d_a, d_b, cnt, width, height, rs, threshold, 1);
        launch_max_filter(d_b, d_a, cnt, width, height, 3, radius, d0.stream);
        ctx->stats.kernel_launches += 2;
        st = ctx->check(cudaGetLastError(), "zgap kernels");
        if (st == CDS_OK) st = ctx->check(cudaMemcpyAsync(zgap_out + (size_t) i0 * bytes, d_a, (size_t) cnt * bytes, cudaMemcpyDeviceToHost, d0.stream), "zgap D2H");
        if (st == CDS_OK) st = ctx->check(cudaStreamSynchronize(d0.stream), "zgap");
    }
    if (d_a) cudaFree(d_a);
    if (d_b) cudaFree(d_b);
    return st;
}

// Targets either as pixels (target_rgb) or as TIFF files stored back to back (blob + offsets, decoded on the device).
static cds_status shape_score_pairs_impl(cds_ctx *ctx, const cds_shape_maskset *sms_c, const uint8_t *target_rgb,
                                         const uint8_t *blob, const int64_t *offsets, const uint16_t *gradient,
                                         const uint8_t *zgap_rgb, const uint8_t *has_variants, int64_t n_targets,
                                         const int32_t *pair_mask, const int64_t *pair_target, int64_t n_pairs,
                                         int64_t *gap_out, int64_t *high_expr_out, uint8_t *mirrored_out)
{
    if (!ctx || !sms_c) { set_tls_error("cds_shape_score_pairs: NULL argument"); return CDS_ERR_BAD_ARG; }
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    cds_shape_maskset *sms = const_cast<cds_shape_maskset *>(sms_c);
    if (sms->ctx != ctx) return ctx->fail(CDS_ERR_BAD_ARG, "cds_shape_score_pairs: mask set belongs to another context");
    if (n_pairs < 0 || n_targets < 0) return ctx->fail(CDS_ERR_BAD_ARG, "cds_shape_score_pairs: negative count");
    if (n_pairs == 0) return CDS_OK;
    if (!pair_mask || !pair_target || !gap_out || !high_expr_out || !mirrored_out) return ctx->fail(CDS_ERR_BAD_ARG, "cds_shape_score_pairs: NULL pair arrays");
    const int M = (int) sms->masks.size();
    bool any_scored = false;
    for (int64_t i = 0; i < n_pairs; i++) {
        if (pair_mask[i] < 0 || pair_mask[i] >= M || pair_target[i] < 0 || pair_target[i] >= n_targets)
            return ctx->fail(CDS_ERR_BAD_ARG, "cds_shape_score_pairs: pair index out of range");
        if (!has_variants || has_variants[pair_target[i]]) any_scored = true;
    }
    // a missing gradient can only be expressed through has_variants; a NULL gradient array with scorable pairs is an error
    const bool from_files = blob != nullptr && offsets != nullptr;
    if (any_scored && ((!target_rgb && !from_files) || !gradient)) return ctx->fail(CDS_ERR_BAD_ARG, "cds_shape_score_pairs: target / gradient images are NULL");
    ctx->stats = cds_search_stats{};
    DevState &d0 = ctx->devs[0];
    SH_CUDA(ctx, cudaSetDevice(d0.dev));
    SH_CUDA(ctx, ensure_shape_lut(d0.dev));
    const int W = sms->W, H = sms->H, bpitch = sms->bpitch;
    const size_t px = (size_t) W * H, bytes = px * 3;
    const size_t bm_words = (size_t) H * bpitch;

    if (sms->descs_dirty) {
        std::vector<ShapeMaskDesc> h(std::max(M, 1));
        for (int i = 0; i < M; i++) {
            h[i].gap_list = sms->masks[i].gap_list; h[i].he_n = sms->masks[i].he_n; h[i].he_m = sms->masks[i].he_m;
            h[i].n_gap = sms->masks[i].n_gap; h[i].pad = 0;
        }
        if (sms->d_descs) { cudaFree(sms->d_descs); sms->d_descs = nullptr; }
        SH_CUDA(ctx, cudaMalloc(&sms->d_descs, h.size() * sizeof(ShapeMaskDesc)));
        SH_CUDA(ctx, cudaMemcpy(sms->d_descs, h.data(), h.size() * sizeof(ShapeMaskDesc), cudaMemcpyHostToDevice));
        sms->descs_dirty = false;
    }

    uint16_t *d_zslice = nullptr, *d_grad = nullptr;
    uint32_t *d_tsig = nullptr;
    uint8_t *d_t = nullptr, *d_z = nullptr, *d_tmp = nullptr, *d_has = nullptr, *d_comp = nullptr;
    TiffStrip *d_strips = nullptr;
    size_t comp_cap = 0, strips_cap = 0;
    std::vector<TiffStrip> strips;
    int32_t *d_pm = nullptr;
    int64_t *d_pt = nullptr;
    long long *d_gap = nullptr, *d_he = nullptr;
    uint8_t *d_mir = nullptr;
    const int64_t chunk = 32;
    const int64_t nt = std::max<int64_t>(n_targets, 1);
    // buffers come from the device's caching pool: consecutive calls (one per mask batch in gradientScores) reuse them
    DevPool &pool = d0.pool;
    cds_status st = ctx->check(pool.alloc((void **) &d_zslice, nt * px * sizeof(uint16_t)), "cudaMalloc(zslice)");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_grad, nt * px * sizeof(uint16_t)), "cudaMalloc(gradient)");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_tsig, nt * bm_words * sizeof(uint32_t)), "cudaMalloc(tsig)");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_t, 2 * chunk * bytes), "cudaMalloc");          // two upload halves
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_z, 2 * chunk * bytes), "cudaMalloc");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_tmp, chunk * bytes), "cudaMalloc");
    if (st == CDS_OK && from_files && n_targets > 0) {
        for (int64_t i = 0; i <= n_targets; i++)
            if (offsets[i] < 0 || (i > 0 && offsets[i] < offsets[i - 1])) st = ctx->fail(CDS_ERR_BAD_ARG, "cds_shape_score_pairs_tiff: offsets must be non-decreasing");
        if (st == CDS_OK) {
            ingest_bounds(offsets, n_targets, chunk, W, H, comp_cap, strips_cap);
            st = ctx->check(pool.alloc((void **) &d_comp, comp_cap), "cudaMalloc(files)");
            if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_strips, strips_cap * sizeof(TiffStrip)), "cudaMalloc(strips)");
        }
    }
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_pm, n_pairs * sizeof(int32_t)), "cudaMalloc");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_pt, n_pairs * sizeof(int64_t)), "cudaMalloc");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_gap, n_pairs * sizeof(long long)), "cudaMalloc");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_he, n_pairs * sizeof(long long)), "cudaMalloc");
    if (st == CDS_OK) st = ctx->check(pool.alloc((void **) &d_mir, n_pairs), "cudaMalloc");
    if (st == CDS_OK && has_variants) {
        st = ctx->check(pool.alloc((void **) &d_has, nt), "cudaMalloc");
        if (st == CDS_OK) st = ctx->check(cudaMemcpyAsync(d_has, has_variants, n_targets, cudaMemcpyHostToDevice, d0.stream), "has_variants H2D");
    }
    if (st == CDS_OK && any_scored) {
        // Uploads (target, its gradient, its zgap image when given) run on the copy stream one chunk ahead of the kernels that
        // turn chunk i into slice / signal planes on the main stream.
        if (st == CDS_OK) st = ctx->check(cudaStreamSynchronize(d0.stream), "shape pairs");     // pooled buffers may still be in use by an earlier call
        auto enqueue_upload = [&](int64_t i0) -> cds_status {
            const int64_t cnt = std::min<int64_t>(chunk, n_targets - i0);
            const int slot = (int) ((i0 / chunk) & 1);
            cds_status s2 = CDS_OK;
            if (i0 >= 2 * chunk) s2 = ctx->check(cudaStreamWaitEvent(d0.copy_stream, d0.up_free[slot], 0), "wait");
            if (s2 == CDS_OK && from_files) {
                // the files as stored, decoded on the copy stream (one buffer: uploads and decodes of consecutive chunks are ordered)
                s2 = ingest_chunk(ctx, "cds_shape_score_pairs_tiff", blob, offsets, i0, cnt, W, H, d_comp, comp_cap, d_strips, strips_cap,
                                  d_t + (size_t) slot * chunk * bytes, d0.copy_stream, strips);
            } else if (s2 == CDS_OK) {
                s2 = ctx->check(cudaMemcpyAsync(d_t + (size_t) slot * chunk * bytes, target_rgb + (size_t) i0 * bytes, (size_t) cnt * bytes, cudaMemcpyHostToDevice, d0.copy_stream), "target H2D");
                ctx->stats.h2d_bytes += (int64_t) (cnt * bytes);
            }
            if (s2 == CDS_OK) s2 = ctx->check(cudaMemcpyAsync(d_grad + (size_t) i0 * px, gradient + (size_t) i0 * px, (size_t) cnt * px * sizeof(uint16_t), cudaMemcpyHostToDevice, d0.copy_stream), "gradient H2D");
            ctx->stats.h2d_bytes += (int64_t) (cnt * px * sizeof(uint16_t));
            if (s2 == CDS_OK && zgap_rgb) {
                s2 = ctx->check(cudaMemcpyAsync(d_z + (size_t) slot * chunk * bytes, zgap_rgb + (size_t) i0 * bytes, (size_t) cnt * bytes, cudaMemcpyHostToDevice, d0.copy_stream), "zgap H2D");
                ctx->stats.h2d_bytes += (int64_t) (cnt * bytes);
            }
            if (s2 == CDS_OK) s2 = ctx->check(cudaEventRecord(d0.up_done[slot], d0.copy_stream), "record");
            return s2;
        };
        if (st == CDS_OK && n_targets > 0) st = enqueue_upload(0);
        for (int64_t i0 = 0; i0 < n_targets && st == CDS_OK; i0 += chunk) {
            const int64_t cnt = std::min<int64_t>(chunk, n_targets - i0);
            const int slot = (int) ((i0 / chunk) & 1);
            uint8_t *ct = d_t + (size_t) slot * chunk * bytes, *cz = d_z + (size_t) slot * chunk * bytes;
            if (i0 + chunk < n_targets) st = enqueue_upload(i0 + chunk);
            if (st != CDS_OK) break;
            st = ctx->check(cudaStreamWaitEvent(d0.stream, d0.up_done[slot], 0), "wait");
            if (st != CDS_OK) break;
            if (!zgap_rgb) {
                // the zgap image the reference's tests derive: maxFilter(10)(mask(threshold)(clearLabels(target)))
                clear_and_mask_kernel<<<148 * 4, 256, 0, d0.stream>>>(ct, d_tmp, cnt, W, H, sms->rects, sms->query_threshold, 1);
                launch_max_filter(d_tmp, cz, cnt, W, H, 3, 10, d0.stream);
                ctx->stats.kernel_launches += 2;
            }
            dim3 grid(H, (unsigned) cnt);
            target_planes_kernel<<<grid, 256, 0, d0.stream>>>(ct, cz, W, H, sms->rects, sms->query_threshold, bpitch,
                                                              d_zslice + (size_t) i0 * px, d_tsig + (size_t) i0 * bm_words);
            ctx->stats.kernel_launches++;
            st = ctx->check(cudaGetLastError(), "target planes");
            if (st == CDS_OK) st = ctx->check(cudaEventRecord(d0.up_free[slot], d0.stream), "record");
        }
    }
    if (st == CDS_OK) st = ctx->check(cudaMemcpyAsync(d_pm, pair_mask, n_pairs * sizeof(int32_t), cudaMemcpyHostToDevice, d0.stream), "pairs H2D");
    if (st == CDS_OK) st = ctx->check(cudaMemcpyAsync(d_pt, pair_target, n_pairs * sizeof(int64_t), cudaMemcpyHostToDevice, d0.stream), "pairs H2D");
    if (st == CDS_OK) {
        cudaEventRecord(d0.ev0, d0.stream);
        for (int64_t p0 = 0; p0 < n_pairs; p0 += (1 << 30)) {
            const int64_t cnt = std::min<int64_t>(1 << 30, n_pairs - p0);
            shape_pair_kernel<<<(unsigned) cnt, 256, 0, d0.stream>>>(sms->d_descs, d_pm + p0, d_pt + p0, d_has, d_zslice, d_grad, d_tsig, W, H, bpitch,
                                                                      sms->mirror, d_gap + p0, d_he + p0, d_mir + p0);
            ctx->stats.kernel_launches++;
            ctx->stats.match_kernel_launches++;
        }
        cudaEventRecord(d0.ev1, d0.stream);
        st = ctx->check(cudaGetLastError(), "shape_pair_kernel");
    }
    if (st == CDS_OK) st = ctx->check(cudaMemcpyAsync(gap_out, d_gap, n_pairs * sizeof(long long), cudaMemcpyDeviceToHost, d0.stream), "gap D2H");
    if (st == CDS_OK) st = ctx->check(cudaMemcpyAsync(high_expr_out, d_he, n_pairs * sizeof(long long), cudaMemcpyDeviceToHost, d0.stream), "he D2H");
    if (st == CDS_OK) st = ctx->check(cudaMemcpyAsync(mirrored_out, d_mir, n_pairs, cudaMemcpyDeviceToHost, d0.stream), "mirrored D2H");
    if (st == CDS_OK) st = ctx->check(cudaStreamSynchronize(d0.stream), "shape pairs");
    if (st == CDS_OK) {
        float ms = 0;
        cudaEventElapsedTime(&ms, d0.ev0, d0.ev1);
        ctx->stats.match_kernel_ms = ms;
        ctx->stats.total_device_ms = ms;
        ctx->stats.comparisons = n_pairs;
        ctx->stats.d2h_bytes = n_pairs * 17;
    }
    if (st != CDS_OK) { cudaStreamSynchronize(d0.copy_stream); cudaStreamSynchronize(d0.stream); cudaGetLastError(); }
    for (void *p : {(void *) d_zslice, (void *) d_grad, (void *) d_tsig, (void *) d_t, (void *) d_z, (void *) d_tmp, (void *) d_has, (void *) d_pm,
                    (void *) d_pt, (void *) d_gap, (void *) d_he, (void *) d_mir, (void *) d_comp, (void *) d_strips})
        pool.free(p);
    return st;
}

extern "C" cds_status cds_shape_score_pairs(cds_ctx *ctx, const cds_shape_maskset *sms, const uint8_t *target_rgb, const uint16_t *gradient,
                                            const uint8_t *zgap_rgb, const uint8_t *has_variants, int64_t n_targets,
                                            const int32_t *pair_mask, const int64_t *pair_target, int64_t n_pairs,
                                            int64_t *gap_out, int64_t *high_expr_out, uint8_t *mirrored_out)
{
    return shape_score_pairs_impl(ctx, sms, target_rgb, nullptr, nullptr, gradient, zgap_rgb, has_variants, n_targets, pair_mask, pair_target, n_pairs,
                                  gap_out, high_expr_out, mirrored_out);
}

extern "C" cds_status cds_shape_score_pairs_tiff(cds_ctx *ctx, const cds_shape_maskset *sms, const uint8_t *blob, const int64_t *offsets,
                                                 const uint16_t *gradient, const uint8_t *zgap_rgb, const uint8_t *has_variants, int64_t n_targets,
                                                 const int32_t *pair_mask, const int64_t *pair_target, int64_t n_pairs,
                                                 int64_t *gap_out, int64_t *high_expr_out, uint8_t *mirrored_out)
{
    if (n_targets > 0 && (!blob || !offsets)) { set_tls_error("cds_shape_score_pairs_tiff: NULL files"); return CDS_ERR_BAD_ARG; }
    return shape_score_pairs_impl(ctx, sms, nullptr, blob, offsets, gradient, zgap_rgb, has_variants, n_targets, pair_mask, pair_target, n_pairs,
                                  gap_out, high_expr_out, mirrored_out);
}
