// cds_topk.cuh -- per-mask top-K selection over the score words of one device.
#ifndef CDS_TOPK_CUH
#define CDS_TOPK_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace cds {

// key: ascending order == (score descending, local target index ascending)
//   bits 34..63 = 0x3FFFFFFF - score, bits 1..33 = local target index, bit 0 = mirrored
__host__ __device__ inline uint64_t topk_make_key(int32_t score, int64_t idx, int mir)
{
    return ((uint64_t) (0x3FFFFFFF - score) << 34) | ((uint64_t) idx << 1) | (uint64_t) (mir & 1);
}
inline void topk_decode_key(uint64_t key, int32_t &score, int64_t &idx, uint8_t &mir)
{
    score = 0x3FFFFFFF - (int32_t) (key >> 34);
    idx = (int64_t) ((key >> 1) & ((1ull << 33) - 1));
    mir = (uint8_t) (key & 1);
}

inline int topk_max_k() { return 4096; }

// scores: [n_masks][n_targets] score words (count | mirrored << 30).  For mask m keeps the entries with
// count >= min_score[m], selects the k best by (count desc, index asc) and writes their keys, sorted, to
// keys_out[m * k ..]; counts_out[m] = number written.  The index stored in a key is idx_base + column.
void launch_topk(const int32_t *scores, int n_masks, int64_t n_targets, const int32_t *min_score, int k, int64_t idx_base,
                 uint64_t *keys_out, int32_t *counts_out, cudaStream_t s);
// Streaming searches: folds one chunk's sorted lists into the running ones (run_keys[m * k ..], run_counts[m]).
void launch_topk_merge(uint64_t *run_keys, int32_t *run_counts, const uint64_t *chunk_keys, const int32_t *chunk_counts,
                       int n_masks, int k, cudaStream_t s);

// Appends every (mask, target) of scores[n_masks][n_targets] whose count reaches min_score[mask] to an unordered device list
// (keys as above with index idx_base + column; masks[slot] = first_mask + row).  *counter counts all of them, also past capacity.
void launch_collect_matches(const int32_t *scores, int n_masks, int64_t n_targets, const int32_t *min_score, int first_mask, int64_t idx_base,
                            uint64_t *keys, int32_t *masks, unsigned long long *counter, unsigned long long capacity, cudaStream_t s);

}  // namespace cds
#endif
