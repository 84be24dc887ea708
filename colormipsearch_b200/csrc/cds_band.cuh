// cds_band.cuh -- the band kernel: batched pixel match with one mask pixel per lane.  Predecessor of the candidate kernel
// (cds_cand.cuh), kept as the batched fallback for mask sets whose tolerance is too wide for palettes and as a cross-check.
#ifndef CDS_BAND_CUH
#define CDS_BAND_CUH

#include "cds_kernels.cuh"

namespace cds {

// xyShift in {0, 2, 4} and a row pitch that fits at least a few rows in shared memory
bool band_kernel_supported(int xy_shift, const PlaneGeom &g);
// below this many masks per launch the gather kernel is used instead
int band_min_masks();
// Launches the band kernel for masks [0, n_masks) x targets [0, n_targets) of one device; returns the number of
// kernel launches issued (0 on configuration error, cudaGetLastError has it).
// `occ` are the library's occupancy bitmaps for this xy_shift (cds_kernels.cuh launch_occupancy), sector pitch `bpitch` words;
// this kernel reads their all-sector row.
// masks[0] must be the first mask of a palette group (index multiple of CDS_PALETTE_GROUP in its mask set) and `groups`
// that group's descriptor.
int launch_pixelmatch_band(const MaskDesc *masks, int n_masks, const uint32_t *planes, PlaneGeom g, int64_t n_targets,
                           const uint32_t *occ, int bpitch, const PaletteGroup *groups, int xy_shift, bool mirror,
                           int32_t *scores, const MatchScratch &scratch, cudaStream_t s);

}  // namespace cds
#endif
