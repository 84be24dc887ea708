// cds_common.h -- data formats shared by the host runtime and the CUDA kernels of libcdsgpu.
//
// TARGET CODE WORD (one uint32 per target pixel, the form a target MIP takes in HBM)
//   A target pixel only ever enters the reference's predicate `calculatePixelGap(mask, target) <= zTolerance`
//   (colormipsearch-api/src/main/java/org/janelia/colormipsearch/cds/AbstractColorDepthSearchAlgorithm.java:157-390)
//   through (sector, ratio) -- sector in {BR,BG,GB,GR,RG,RB,none} picked by strict inequalities (:195-257), ratio =
//   second/max as an IEEE double -- and enters the threshold test (PixelMatchColorDepthSearchAlgorithm.java:250) through its
//   maximum channel.  There are exactly CDS_NUM_RANKS = 19820 distinct doubles a/b with 0 <= a < b <= 255, and IEEE rounding is
//   monotone, so for one mask pixel the set of matching target (sector, ratio) classes is at most two *rank intervals*:
//   one in the mask's own sector and one in a single neighbouring sector.  The device therefore never touches a double:
//
//     bits  0..7   max channel value                      (re-bake the threshold flag without the RGB)
//     bits  8..25  SR = sector * 32768 + rank(ratio)      sector 0..5 = BR,BG,GB,GR,RG,RB;  SR_NONE = 6*32768 for "no sector"
//     bit   30     pad: guard columns / guard rows around the image, never matches, survives re-baking
//     bit   31     below: max channel <= the data threshold currently baked into the library
//
//   A mask pixel carries [lo, lo+len] intervals pre-shifted by 8 (low byte of len = 0xFF), so the whole per-gather test is
//   (code - lo1) <=u len1  ||  (code - lo2) <=u len2, which is false for every word with bit 30 or 31 set.
#ifndef CDS_COMMON_H
#define CDS_COMMON_H

#include <stdint.h>

#define CDS_NUM_RANKS 19820          // distinct values of fl(a/b), 0 <= a < b <= 255 (checked at start-up)
#define CDS_SECTOR_STRIDE 32768
#define CDS_NUM_SECTORS 6
#define CDS_SR_NONE (CDS_NUM_SECTORS * CDS_SECTOR_STRIDE)
#define CDS_NUM_CLASSES (CDS_NUM_SECTORS * CDS_NUM_RANKS)   // mask classes that can match anything

#define CDS_CODE_SR_SHIFT 8
#define CDS_CODE_PAD_BIT 0x40000000u
#define CDS_CODE_BELOW_BIT 0x80000000u
#define CDS_CODE_PAD_WORD (CDS_CODE_PAD_BIT | CDS_CODE_BELOW_BIT | ((uint32_t) CDS_SR_NONE << CDS_CODE_SR_SHIFT))

#define CDS_EMPTY_LO 0x7F000000u     // no code word lies in [0x7F000000, 0x7F0000FF]: bits 26..29 of a code are always 0
#define CDS_EMPTY_LEN 0u

#define CDS_GUARD_ROWS 4             // pad rows above / below every plane = largest xyShift served by the band kernel
#define CDS_MIN_PAD_COLS 8           // pad columns after every row  (>= CDS_GUARD_ROWS)

// One mask pixel as the kernels see it (16 bytes, one LDG.128).
//   x | y << 16 ; lo1 ; lo2 ; len1 | len2 << 16   (lo* pre-shifted by CDS_CODE_SR_SHIFT, len* in SR units)
struct __attribute__((aligned(16))) cds_mask_record {
    uint32_t xy;
    uint32_t lo1;
    uint32_t lo2;
    uint32_t lens;
};

// Interval pair of one mask class, SR units (host side / device lookup table).
struct cds_class_interval {
    uint32_t lo1, len1, lo2, len2;   // empty interval: lo = CDS_IV_EMPTY, len = 0
};

#define CDS_IV_EMPTY 0xFFFFFFFFu

// COMPACT MASK RECORDS.  The two intervals of a mask pixel depend only on its colour class (sector, rank), and the masks
// of a group share few classes (colour-depth MIPs are rendered with a 256-entry LUT), so the band kernel keeps one
// PALETTE per group of CDS_PALETTE_GROUP masks in shared memory and streams 4-byte records:
//     record  = x | y << 11 | palette index << 21              (W <= 2048, H <= 1024, <= CDS_PALETTE_SIZE classes per group)
//     palette = { lo1 | len1 << 18 , lo2 | len2 << 18 }         SR units, 18 + 14 bits; empty interval: lo = CDS_PAL_EMPTY_LO
// Groups that do not fit (more classes, larger images, tolerances so wide that an interval spans >= 16384 ranks) use the
// 16-byte records above.
#define CDS_PALETTE_GROUP 1024
#define CDS_PALETTE_SIZE 2048
#define CDS_PAL_LO_BITS 18
#define CDS_PAL_EMPTY_LO 0x3FFFFu     // an SR no code word has (> CDS_SR_NONE)
#define CDS_PAL_MAX_LEN 16383u
#define CDS_CLASS_NONE_INDEX CDS_NUM_CLASSES   // class index of "no sector" pixels in the per-record class array

#define CDS_MAX_VARIANTS 34          // band kernel: 17 offsets x 2 orientations (xyShift 4)
#define CDS_MAX_SHIFT_OFFSETS 40     // generic kernel: 1 + 9*4 entries for xyShift 8

#endif
