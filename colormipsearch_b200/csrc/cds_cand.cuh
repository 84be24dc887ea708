// cds_cand.cuh -- the candidate kernel: batched pixel match driven by 32-pixel mask words (the hot path at scale).
#ifndef CDS_CAND_CUH
#define CDS_CAND_CUH

#include "cds_kernels.cuh"

namespace cds {

// WORD LIST of a mask group (built once per mask set, per device).  A mask's pixel set is kept as the non-zero 32-bit words
// of two bitmaps in TARGET coordinates: orientation 0 has bit x of row y set for a mask pixel (x, y); orientation 1 (only
// when the mask set mirrors) has bit W-1-x set.  The words of the CDS_PALETTE_GROUP masks of a group form ONE list ordered by
// image row (then mask, orientation, column), with a row-start table gstart[H+1], so the entries that concern a band of rows
// are one contiguous range that is cut into equal tickets regardless of mask boundaries.  One 16-byte entry per word:
//     bits : the word
//     meta : y | word column << 10 | orientation << 16 | mask index inside the group << 22      (H <= 1024, W <= 2048)
//     rec  : record index (inside its mask) of the word's LOWEST set bit; the record of set bit b is rec + popc(bits below b)
//            for orientation 0 and rec - popc(bits below b) for orientation 1 (mirroring reverses the order inside a row)
//     0
// ANDing `bits` with the library's occupancy word at the same (row, column) leaves exactly the mask pixels that can match
// in some shifted variant of that orientation -- 32 pixels per instruction instead of one.
constexpr int kWordMetaYBits = 10;
constexpr int kWordMetaOrientBit = 16;
constexpr int kWordMetaMaskShift = 22;

bool cand_kernel_supported(int xy_shift, const PlaneGeom &g);

// Construction, in this order:
//   launch_words_count      wcount[m][y] = entries of (mask m, row y)                         (masks[m].records / rowstart)
//   launch_words_group_rows wcount[m][y] -> offset of mask m inside its group's run of row y;  grow[g][y] = length of that run
//   (host) gstart = exclusive scan of grow over (group, row), absolute entry indices; gstart[g][H] = end of the group
//   launch_words_fill       writes the entries; masks[m].wstart must point at wcount[m]
void launch_words_count(const MaskDesc *masks, int n_masks, int W, int H, bool mirror, uint32_t *wcount, cudaStream_t s);
void launch_words_group_rows(uint32_t *wcount, int n_masks, int H, uint32_t *grow, cudaStream_t s);
void launch_words_fill(const MaskDesc *masks, int n_masks, int W, int H, bool mirror, const uint32_t *gstart, uint4 *words, cudaStream_t s);

// Same contract as launch_pixelmatch_band (cds_band.cuh); every group needs its word list (PaletteGroup::words / gstart).
int launch_pixelmatch_cand(const MaskDesc *masks, int n_masks, const uint32_t *planes, PlaneGeom g, int64_t n_targets,
                           const uint32_t *occ, int bpitch, const PaletteGroup *groups, int xy_shift, bool mirror,
                           int32_t *scores, cudaStream_t s);

}  // namespace cds
#endif
