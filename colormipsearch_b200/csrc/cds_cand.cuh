// cds_cand.cuh -- the candidate kernel: batched pixel match driven by 32-pixel mask words (the hot path at scale).
#ifndef CDS_CAND_CUH
#define CDS_CAND_CUH

#include "cds_kernels.cuh"

namespace cds {

// WORD LIST of a mask (built once per mask set, per device).  The mask's pixel set is kept as the non-zero 32-bit words
// of two bitmaps in TARGET coordinates: orientation 0 has bit x of row y set for a mask pixel (x, y); orientation 1 (only
// when the mask set mirrors) has bit W-1-x set.  One 16-byte entry {bits, meta, rec, 0} per non-zero word, ordered by row:
//     bits : the word
//     meta : y | word column << 10 | orientation << 16                      (H <= 1024, W <= 2048)
//     rec  : record index of the word's LOWEST set bit; the record of set bit b is rec + popc(bits below b) for
//            orientation 0 and rec - popc(bits below b) for orientation 1 (mirroring reverses the order inside a row)
// ANDing `bits` with the library's occupancy word at the same (row, column) leaves exactly the mask pixels that can match
// in some shifted variant of that orientation -- 32 pixels per instruction instead of one.
constexpr int kWordMetaYBits = 10;
constexpr int kWordMetaOrientBit = 16;

bool cand_kernel_supported(int xy_shift, const PlaneGeom &g);

// rows -> number of word-list entries of every (mask, row): wcount[m * (H + 1) + y]  (then scanned in place with
// launch_mask_scan_rows, which also yields the per-mask totals)
void launch_words_count(const MaskDesc *masks, int n_masks, int W, int H, bool mirror, uint32_t *wcount, cudaStream_t s);
// fills masks[m].words (n_words 16-byte entries, see above) given the scanned row starts masks[m].wstart
void launch_words_fill(const MaskDesc *masks, int n_masks, int W, int H, bool mirror, cudaStream_t s);

// Same contract as launch_pixelmatch_band (cds_band.cuh); every mask needs its word list.
int launch_pixelmatch_cand(const MaskDesc *masks, int n_masks, const uint32_t *planes, PlaneGeom g, int64_t n_targets,
                           const uint32_t *occ, int bpitch, const PaletteGroup *groups, int xy_shift, bool mirror,
                           int32_t *scores, cudaStream_t s);

}  // namespace cds
#endif
