// cds_synth.h -- deterministic synthetic colour-depth MIPs, bit-identical on host and device (integer arithmetic only).
//
// Bench / scale-test INPUT DATA, not part of the matching algorithm.  Images imitate what the reference's fixtures look
// like (colormipsearch-api/src/test/resources/colormipsearch/api/cdsearch/{ems,lms}): neurites drawn as chains of thick
// line segments ("capsules") whose colour walks along the 256-entry colour-depth LUT
// (colormipsearch-api/src/main/java/org/janelia/colormipsearch/cds/GradientAreaGapUtils.java:132-155), on black, plus
// non-black text-label blocks inside the two label regions the CLI excludes
// (colormipsearch-tools/src/main/java/org/janelia/colormipsearch/cmd/AbstractColorDepthMatchArgs.java:101-119).
//   kind 0 (EM-like mask)  : 1-3 neurites, full brightness, confined to a random box of ~W/2 x H/3
//   kind 1 (LM-like target): 4-28 neurites over the whole image, brightness 30..255 (part of them below the usual data
//                            thresholds), ~6 % of the neurites grey (no colour sector); every 10th target also contains a
//                            jittered (+-3 px, +-6 slices) copy of one of the first 1024 masks of the same seed, so that
//                            searches have true positives.
// The image is a pure function of (kind, seed, index, W, H).
#ifndef CDS_SYNTH_H
#define CDS_SYNTH_H

#include <stdint.h>

#ifdef __CUDACC__
#define CDS_HD __host__ __device__
#else
#define CDS_HD
#endif

namespace cds {

struct SynthCapsule {
    int16_t x0, y0, x1, y1;
    uint8_t r;          // half width in pixels
    uint8_t z0, z1;     // LUT slice at both ends
    uint8_t bright;     // 0..255
    uint8_t grey;       // 1: r = g = b = bright
    uint8_t pad[3];
};

#define CDS_SYNTH_MAX_CAPS 1024

struct SynthSpec {
    int32_t n;
    int32_t W, H;
    int32_t n_labels;
    int16_t label[2][4];          // x0, y0, x1, y1 of the label blocks
    uint8_t label_rgb[2][4];
    SynthCapsule caps[CDS_SYNTH_MAX_CAPS];
};

CDS_HD inline uint64_t synth_mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct SynthRng {
    uint64_t key, ctr;
    CDS_HD SynthRng(uint64_t seed, uint64_t kind, uint64_t index) : key(synth_mix(synth_mix(seed ^ (kind << 56)) + index)), ctr(0) {}
    CDS_HD uint32_t next() { return (uint32_t) (synth_mix(key + (ctr++) * 0xD1342543DE82EF95ull) >> 32); }
    CDS_HD int range(int lo, int hi) { return lo + (int) (next() % (uint32_t) (hi - lo + 1)); }   // inclusive
};

// 32 unit directions scaled by 256 (round(256 cos), round(256 sin)) -- a table, so no libm on either side
CDS_HD inline void synth_dir(int d, int &cx, int &cy)
{
    const int16_t C[32] = {256, 251, 237, 213, 181, 142, 98, 50, 0, -50, -98, -142, -181, -213, -237, -251,
                           -256, -251, -237, -213, -181, -142, -98, -50, 0, 50, 98, 142, 181, 213, 237, 251};
    d &= 31;
    cx = C[d];
    cy = C[(d + 24) & 31];   // sin(a) = cos(a - 90 deg)
}

CDS_HD inline int synth_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Appends one neurite (a random walk of capsules) to spec.
CDS_HD inline void synth_add_neurite(SynthSpec &s, SynthRng &rng, int bx0, int by0, int bx1, int by1,
                                     int n_seg, int step_lo, int step_hi, int r, int bright, int grey, int z_start, int z_drift)
{
    int x = rng.range(bx0, bx1), y = rng.range(by0, by1);
    int dir = rng.range(0, 31);
    int z = z_start;
    for (int k = 0; k < n_seg && s.n < CDS_SYNTH_MAX_CAPS; k++) {
        dir += rng.range(-3, 3);
        int len = rng.range(step_lo, step_hi);
        int cx, cy;
        synth_dir(dir, cx, cy);
        int nx = x + (cx * len) / 256, ny = y + (cy * len) / 256;
        if (nx < bx0 || nx > bx1 || ny < by0 || ny > by1) {      // bounce: turn around and stay inside the box
            dir += 16;
            nx = synth_clampi(nx, bx0, bx1);
            ny = synth_clampi(ny, by0, by1);
        }
        int nz = synth_clampi(z + rng.range(-z_drift, z_drift), 0, 255);
        SynthCapsule c;
        c.x0 = (int16_t) x; c.y0 = (int16_t) y; c.x1 = (int16_t) nx; c.y1 = (int16_t) ny;
        c.r = (uint8_t) r; c.z0 = (uint8_t) z; c.z1 = (uint8_t) nz; c.bright = (uint8_t) bright; c.grey = (uint8_t) grey;
        c.pad[0] = c.pad[1] = c.pad[2] = 0;
        s.caps[s.n++] = c;
        x = nx; y = ny; z = nz;
    }
}

CDS_HD inline void synth_add_labels(SynthSpec &s, SynthRng &rng)
{
    // text-like blocks inside the CLI's label regions: name label (x < 330, y < 100) and colour scale (x >= W - 270, y < 90)
    s.n_labels = 0;
    if (s.W >= 340 && s.H >= 110) {
        int16_t *l = s.label[s.n_labels];
        l[0] = 8; l[1] = 8; l[2] = (int16_t) rng.range(120, 320); l[3] = (int16_t) rng.range(24, 90);
        s.label_rgb[s.n_labels][0] = 255; s.label_rgb[s.n_labels][1] = 255; s.label_rgb[s.n_labels][2] = 255;
        s.n_labels++;
    }
    if (s.W > 620 && s.H >= 110) {
        int16_t *l = s.label[s.n_labels];
        l[0] = (int16_t) (s.W - 262); l[1] = 6; l[2] = (int16_t) (s.W - 6); l[3] = (int16_t) rng.range(20, 80);
        s.label_rgb[s.n_labels][0] = 255; s.label_rgb[s.n_labels][1] = (uint8_t) rng.range(100, 255); s.label_rgb[s.n_labels][2] = 40;
        s.n_labels++;
    }
}

CDS_HD inline void synth_mask_neurites(SynthSpec &s, uint64_t seed, int64_t index, int jx, int jy, int jz, int bright)
{
    // the neurites of mask `index`; (jx, jy, jz, bright) let a target embed a jittered, dimmer copy
    SynthRng rng(seed, 0, (uint64_t) index);
    const int W = s.W, H = s.H;
    const int bw = W / 2 > 40 ? W / 2 : W - 2, bh = H / 3 > 40 ? H / 3 : H - 2;
    // box below the label strip when the image is large enough
    const int top = H >= 220 ? 100 : 0;
    int bx0 = rng.range(1, W - bw - 1 > 1 ? W - bw - 1 : 1);
    int by0 = rng.range(top, H - bh - 1 > top ? H - bh - 1 : top);
    int bx1 = bx0 + bw - 1 < W - 2 ? bx0 + bw - 1 : W - 2;
    int by1 = by0 + bh - 1 < H - 2 ? by0 + bh - 1 : H - 2;
    const int n_neur = rng.range(1, 3);
    const int size_class = rng.range(0, 15);             // spreads the mask size over roughly a decade
    for (int k = 0; k < n_neur; k++) {
        int r = rng.range(1, 3);
        int n_seg = 10 + size_class * 4 + rng.range(0, 8);
        int z0 = rng.range(10, 245);
        int first = s.n;
        synth_add_neurite(s, rng, bx0, by0, bx1, by1, n_seg, 10, 34, r, bright, 0, z0, 5);
        for (int i = first; i < s.n; i++) {
            SynthCapsule &c = s.caps[i];
            c.x0 = (int16_t) synth_clampi(c.x0 + jx, 0, W - 1); c.x1 = (int16_t) synth_clampi(c.x1 + jx, 0, W - 1);
            c.y0 = (int16_t) synth_clampi(c.y0 + jy, 0, H - 1); c.y1 = (int16_t) synth_clampi(c.y1 + jy, 0, H - 1);
            c.z0 = (uint8_t) synth_clampi(c.z0 + jz, 0, 255); c.z1 = (uint8_t) synth_clampi(c.z1 + jz, 0, 255);
        }
    }
}

CDS_HD inline void synth_make_spec(int kind, uint64_t seed, int64_t index, int W, int H, SynthSpec &s)
{
    s.n = 0; s.W = W; s.H = H; s.n_labels = 0;
    if (kind == 0) {
        synth_mask_neurites(s, seed, index, 0, 0, 0, 255);
        SynthRng lr(seed, 2, (uint64_t) index);
        synth_add_labels(s, lr);
        return;
    }
    SynthRng rng(seed, 1, (uint64_t) index);
    const int n_neur = rng.range(4, 28);
    for (int k = 0; k < n_neur; k++) {
        int r = rng.range(0, 2);
        int n_seg = rng.range(6, 26);
        int bright = rng.range(30, 255);
        int grey = rng.range(0, 15) == 0;
        int z0 = rng.range(0, 255);
        synth_add_neurite(s, rng, 1, 1, W - 2, H - 2, n_seg, 8, 40, r, bright, grey, z0, 7);
    }
    if (index % 10 == 3) {
        int64_t mi = (index / 10) % 1024;
        int jx = rng.range(-3, 3), jy = rng.range(-3, 3), jz = rng.range(-6, 6);
        int bright = rng.range(90, 255);
        synth_mask_neurites(s, seed, mi, jx, jy, jz, bright);
    }
    SynthRng lr(seed, 3, (uint64_t) index);
    synth_add_labels(s, lr);
}

// squared distance from (px,py) to the capsule's axis and the clamped projection (for the z interpolation)
CDS_HD inline int64_t synth_axis_dist2(const SynthCapsule &c, int px, int py, int64_t &dot_out, int64_t &len2_out)
{
    const int64_t dx = c.x1 - c.x0, dy = c.y1 - c.y0;
    const int64_t ex = px - c.x0, ey = py - c.y0;
    const int64_t len2 = dx * dx + dy * dy;
    int64_t dot = ex * dx + ey * dy;
    int64_t d2;
    if (len2 == 0 || dot <= 0) { dot = 0; d2 = ex * ex + ey * ey; }
    else if (dot >= len2) { dot = len2; const int64_t fx = px - c.x1, fy = py - c.y1; d2 = fx * fx + fy * fy; }
    else {
        // distance^2 = cross^2 / len2, kept exact as a rational comparison by callers that need it; here floor
        const int64_t cross = ex * dy - ey * dx;
        d2 = (cross * cross) / len2;
    }
    dot_out = dot; len2_out = len2;
    return d2;
}

CDS_HD inline bool synth_capsule_hit(const SynthCapsule &c, int px, int py, int &z_out)
{
    const int rr = c.r;
    int xmin = c.x0 < c.x1 ? c.x0 : c.x1, xmax = c.x0 < c.x1 ? c.x1 : c.x0;
    int ymin = c.y0 < c.y1 ? c.y0 : c.y1, ymax = c.y0 < c.y1 ? c.y1 : c.y0;
    if (px < xmin - rr || px > xmax + rr || py < ymin - rr || py > ymax + rr) return false;
    int64_t dot, len2;
    const int64_t d2 = synth_axis_dist2(c, px, py, dot, len2);
    if (d2 > (int64_t) rr * rr) return false;
    int z = c.z0;
    if (len2 > 0) z = c.z0 + (int) (((int64_t) (c.z1 - c.z0) * dot) / len2);
    z_out = z < 0 ? 0 : (z > 255 ? 255 : z);
    return true;
}

// colour of the LUT entry z, the table of GradientAreaGapUtils.java:132-155 regenerated arithmetically is NOT possible
// (the table is hand-made), so callers pass the 256 x 3 table (host: static array, device: __constant__).
CDS_HD inline void synth_shade(const uint8_t *lut, const SynthCapsule &c, int z, uint8_t &r, uint8_t &g, uint8_t &b)
{
    if (c.grey) { r = g = b = c.bright; return; }
    r = (uint8_t) ((lut[3 * z] * (int) c.bright + 127) / 255);
    g = (uint8_t) ((lut[3 * z + 1] * (int) c.bright + 127) / 255);
    b = (uint8_t) ((lut[3 * z + 2] * (int) c.bright + 127) / 255);
}

// One pixel: the capsule with the highest index covering it wins; labels are drawn on top.
// `list`/`n_list` = indices of the capsules whose bounding rows include py (any order), or NULL = all.
CDS_HD inline void synth_pixel(const SynthSpec &s, const uint8_t *lut, const int16_t *list, int n_list, int px, int py,
                               uint8_t &r, uint8_t &g, uint8_t &b)
{
    r = g = b = 0;
    int best = -1, bestz = 0;
    const int n = list ? n_list : s.n;
    for (int k = 0; k < n; k++) {
        const int ci = list ? list[k] : k;
        if (ci < best) continue;
        int z;
        if (synth_capsule_hit(s.caps[ci], px, py, z)) { best = ci; bestz = z; }
    }
    if (best >= 0) synth_shade(lut, s.caps[best], bestz, r, g, b);
    for (int l = 0; l < s.n_labels; l++) {
        // a sparse glyph-like pattern so that label regions are not solid
        if (px >= s.label[l][0] && px < s.label[l][2] && py >= s.label[l][1] && py < s.label[l][3] &&
            (((px >> 1) + (py >> 2)) % 3 != 0)) {
            r = s.label_rgb[l][0]; g = s.label_rgb[l][1]; b = s.label_rgb[l][2];
        }
    }
}

CDS_HD inline uint32_t synth_isqrt(uint64_t v)
{
    uint64_t r = 0, bit = 1ull << 62;
    while (bit > v) bit >>= 2;
    while (bit) {
        if (v >= r + bit) { v -= r + bit; r = (r >> 1) + bit; }
        else r >>= 1;
        bit >>= 2;
    }
    return (uint32_t) r;
}

// gradient image of a synthetic target: distance (pixels, floor) to the nearest neurite surface, capped at 650
CDS_HD inline uint16_t synth_gradient_pixel(const SynthSpec &s, int px, int py)
{
    int best = 650;
    for (int k = 0; k < s.n; k++) {
        const SynthCapsule &c = s.caps[k];
        // cheap lower bound from the bounding box
        int xmin = c.x0 < c.x1 ? c.x0 : c.x1, xmax = c.x0 < c.x1 ? c.x1 : c.x0;
        int ymin = c.y0 < c.y1 ? c.y0 : c.y1, ymax = c.y0 < c.y1 ? c.y1 : c.y0;
        int bx = px < xmin ? xmin - px : (px > xmax ? px - xmax : 0);
        int by = py < ymin ? ymin - py : (py > ymax ? py - ymax : 0);
        int lb = (bx > by ? bx : by) - c.r;
        if (lb >= best) continue;
        int64_t dot, len2;
        const int64_t d2 = synth_axis_dist2(c, px, py, dot, len2);
        int d = (int) synth_isqrt((uint64_t) d2) - c.r;
        if (d < 0) d = 0;
        if (d < best) best = d;
    }
    return (uint16_t) best;
}

}  // namespace cds
#endif
