// cds_topk.cu -- per-mask top-K on the device.
//
// The reference keeps every pair that passes ColorMIPSearch.isMatch (API/cds/ColorMIPSearch.java:42-45) and sorts by
// descending matchingPixels with a stable sort (TOOLS/ColorDepthSearchCmd.java:403-409, API/results/ItemsHandling.java:61-63);
// the K best of that list, ties in ascending target order, is what this kernel returns.
// One CTA per mask: a 3-pass radix select (10 bits per pass) finds the K-th largest count exactly, then one ordered
// pass collects everything above it plus the first ties, and a shared-memory bitonic sort orders the <= 4096 keys.
#include "cds_topk.cuh"
#include "cds_kernels.cuh"

namespace cds {

namespace {
constexpr int kThreads = 256;
constexpr int kMaxK = 4096;

__device__ __forceinline__ int block_exclusive_scan_flags(bool flag, int *warp_tot, int &block_total)
{
    // returns the number of set flags in threads before this one; block_total = all of them
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned bal = __ballot_sync(0xffffffffu, flag);
    int before = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
        int c = warp_tot[w];
        if (w < warp) off += c;
        tot += c;
    }
    __syncthreads();
    block_total = tot;
    return off + before;
}
}  // namespace

__global__ void __launch_bounds__(kThreads) topk_kernel(const int32_t *__restrict__ scores, int64_t n_targets,
                                                        const int32_t *__restrict__ min_score, int k, int64_t idx_base,
                                                        uint64_t *__restrict__ keys_out, int32_t *__restrict__ counts_out)
{
    __shared__ uint64_t s_keys[kMaxK];
    __shared__ int s_hist[1024];
    __shared__ int s_warp[kThreads / 32];
    __shared__ int s_sel_bin, s_sel_above, s_gt_count;

    const int m = blockIdx.x;
    const int32_t *row = scores + (size_t) m * n_targets;
    const int minsc = max(min_score[m], 1);

    // ---- radix select of the k-th largest count among entries with count >= minsc
    int prefix = 0;          // high bits fixed so far
    int prefix_mask = 0;
    int need = k;            // rank (1-based, from the top) still to locate inside the current prefix
    bool take_all = false;
    for (int pass = 0; pass < 3; pass++) {
        const int shift = 20 - 10 * pass;
        for (int i = threadIdx.x; i < 1024; i += kThreads) s_hist[i] = 0;
        __syncthreads();
        for (int64_t i = threadIdx.x; i < n_targets; i += kThreads) {
            int c = row[i] & ~CDS_SCORE_MIRROR_BIT;
            if (c >= minsc && (c & prefix_mask) == prefix) atomicAdd(&s_hist[(c >> shift) & 1023], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0, bin = -1, above = 0;
            for (int b = 1023; b >= 0; b--) {
                if (acc + s_hist[b] >= need) { bin = b; above = acc; break; }
                acc += s_hist[b];
            }
            if (bin < 0) above = acc;     // fewer than `need` entries under this prefix
            s_sel_bin = bin;
            s_sel_above = above;
        }
        __syncthreads();
        const int bin = s_sel_bin, above = s_sel_above;
        if (bin < 0) { take_all = true; break; }   // only possible in pass 0: fewer than k passing entries
        need -= above;
        prefix |= bin << shift;
        prefix_mask |= 1023 << shift;
        __syncthreads();
    }
    const int sK = take_all ? minsc : prefix;       // k-th largest count (or the filter floor)
    // entries with count > sK are all kept; of the ties (== sK) the first `need` in index order
    const int ties_wanted = take_all ? 0x7fffffff : need;

    if (threadIdx.x == 0) s_gt_count = 0;
    __syncthreads();
    // pass A: count entries strictly above sK so that ties can be placed after them
    {
        int local = 0;
        for (int64_t i = threadIdx.x; i < n_targets; i += kThreads) {
            int c = row[i] & ~CDS_SCORE_MIRROR_BIT;
            local += (c >= minsc && c > sK) ? 1 : 0;
        }
        local = __reduce_add_sync(0xffffffffu, local);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_gt_count, local);
    }
    __syncthreads();
    const int n_gt = take_all ? 0 : s_gt_count;
    __syncthreads();
    if (threadIdx.x == 0) s_gt_count = 0;           // reuse as the running slot for "above" entries
    __syncthreads();

    // pass B: ordered collection
    int ties_seen = 0;
    for (int64_t base = 0; base < n_targets; base += kThreads) {
        const int64_t i = base + threadIdx.x;
        int w = 0, c = 0;
        bool valid = i < n_targets;
        if (valid) { w = row[i]; c = w & ~CDS_SCORE_MIRROR_BIT; }
        const bool pass = valid && c >= minsc;
        const bool gt = pass && !take_all && c > sK;
        const bool eq = pass && (take_all ? true : c == sK);
        int eq_total;
        int eq_rank = block_exclusive_scan_flags(eq, s_warp, eq_total);
        if (gt) {
            int slot = atomicAdd(&s_gt_count, 1);
            s_keys[slot] = topk_make_key(c, idx_base + i, (w & CDS_SCORE_MIRROR_BIT) ? 1 : 0);
        }
        if (eq) {
            int r = ties_seen + eq_rank;
            if (r < ties_wanted && n_gt + r < k) s_keys[n_gt + r] = topk_make_key(c, idx_base + i, (w & CDS_SCORE_MIRROR_BIT) ? 1 : 0);
        }
        ties_seen += eq_total;
    }
    __syncthreads();
    int n_out = n_gt + min(ties_seen, ties_wanted);
    n_out = min(n_out, k);

    // ---- bitonic sort of s_keys[0..n_out) (padded with the largest key)
    int n_pow2 = 1;
    while (n_pow2 < n_out) n_pow2 <<= 1;
    for (int i = n_out + threadIdx.x; i < n_pow2; i += kThreads) s_keys[i] = ~0ull;
    __syncthreads();
    for (int size = 2; size <= n_pow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < n_pow2 / 2; t += kThreads) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool asc = ((lo & size) == 0);
                uint64_t a = s_keys[lo], b = s_keys[hi];
                if ((a > b) == asc) { s_keys[lo] = b; s_keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n_out; i += kThreads) keys_out[(size_t) m * k + i] = s_keys[i];
    if (threadIdx.x == 0) counts_out[m] = n_out;
}

void launch_topk(const int32_t *scores, int n_masks, int64_t n_targets, const int32_t *min_score, int k, int64_t idx_base,
                 uint64_t *keys_out, int32_t *counts_out, cudaStream_t s)
{
    if (n_masks == 0) return;
    topk_kernel<<<n_masks, kThreads, 0, s>>>(scores, n_targets, min_score, k, idx_base, keys_out, counts_out);
}

// run[m] <- the k smallest keys of run[m] U chunk[m] (both sorted ascending, keys are unique: they carry the target index).
// One CTA per mask: merge-path ranks, no sort -- element i of one list lands at i + (number of smaller keys in the other).
__global__ void __launch_bounds__(kThreads) topk_merge_kernel(uint64_t *__restrict__ run_keys, int32_t *__restrict__ run_counts,
                                                              const uint64_t *__restrict__ chunk_keys,
                                                              const int32_t *__restrict__ chunk_counts, int k)
{
    extern __shared__ uint64_t s_merge[];       // 2 * k keys
    uint64_t *s_a = s_merge, *s_b = s_merge + k;
    const int m = blockIdx.x;
    const int na = min(run_counts[m], k), nb = min(chunk_counts[m], k);
    if (nb == 0) return;
    uint64_t *a = run_keys + (size_t) m * k;
    const uint64_t *b = chunk_keys + (size_t) m * k;
    for (int i = threadIdx.x; i < na; i += kThreads) s_a[i] = a[i];
    for (int i = threadIdx.x; i < nb; i += kThreads) s_b[i] = b[i];
    __syncthreads();
    auto lower_bound = [](const uint64_t *v, int n, uint64_t key) {
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (v[mid] < key) lo = mid + 1; else hi = mid; }
        return lo;
    };
    for (int i = threadIdx.x; i < na; i += kThreads) {
        const int pos = i + lower_bound(s_b, nb, s_a[i]);
        if (pos < k) a[pos] = s_a[i];
    }
    for (int i = threadIdx.x; i < nb; i += kThreads) {
        const int pos = i + lower_bound(s_a, na, s_b[i]);
        if (pos < k) a[pos] = s_b[i];
    }
    if (threadIdx.x == 0) run_counts[m] = min(na + nb, k);
}

void launch_topk_merge(uint64_t *run_keys, int32_t *run_counts, const uint64_t *chunk_keys, const int32_t *chunk_counts,
                       int n_masks, int k, cudaStream_t s)
{
    if (n_masks == 0) return;
    const size_t smem = (size_t) 2 * k * sizeof(uint64_t);
    if (smem > 48 * 1024) cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    topk_merge_kernel<<<n_masks, kThreads, smem, s>>>(run_keys, run_counts, chunk_keys, chunk_counts, k);
}

// Every (mask, target) whose count reaches the mask's isMatch floor: key + mask index appended to an unordered device list.
// `counter` keeps counting past `capacity`, so the caller learns how much room a retry needs.
__global__ void __launch_bounds__(kThreads) collect_matches_kernel(const int32_t *__restrict__ scores, int64_t n_targets,
                                                                   const int32_t *__restrict__ min_score, int first_mask, int64_t idx_base,
                                                                   uint64_t *__restrict__ keys, int32_t *__restrict__ masks,
                                                                   unsigned long long *__restrict__ counter, unsigned long long capacity)
{
    const int64_t t = (int64_t) blockIdx.x * kThreads + threadIdx.x;
    const int m = blockIdx.y;
    const int lane = threadIdx.x & 31;
    int w = 0, c = 0;
    bool pass = false;
    if (t < n_targets) {
        w = scores[(size_t) m * n_targets + t];
        c = w & ~CDS_SCORE_MIRROR_BIT;
        pass = c >= max(min_score[m], 1);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (bal == 0) return;
    unsigned long long base = 0;
    if (lane == __ffs((int) bal) - 1) base = atomicAdd(counter, (unsigned long long) __popc(bal));
    base = __shfl_sync(0xffffffffu, base, __ffs((int) bal) - 1);
    if (pass) {
        const unsigned long long slot = base + (unsigned long long) __popc(bal & ((1u << lane) - 1u));
        if (slot < capacity) {
            keys[slot] = topk_make_key(c, idx_base + t, (w & CDS_SCORE_MIRROR_BIT) ? 1 : 0);
            masks[slot] = first_mask + m;
        }
    }
}

void launch_collect_matches(const int32_t *scores, int n_masks, int64_t n_targets, const int32_t *min_score, int first_mask, int64_t idx_base,
                            uint64_t *keys, int32_t *masks, unsigned long long *counter, unsigned long long capacity, cudaStream_t s)
{
    if (n_targets == 0) return;
    for (int m0 = 0; m0 < n_masks; m0 += 32768) {
        const int cnt = n_masks - m0 < 32768 ? n_masks - m0 : 32768;
        dim3 grid((unsigned) ((n_targets + kThreads - 1) / kThreads), (unsigned) cnt);
        collect_matches_kernel<<<grid, kThreads, 0, s>>>(scores + (size_t) m0 * n_targets, n_targets, min_score + m0, first_mask + m0, idx_base,
                                                        keys, masks, counter, capacity);
    }
}

}  // namespace cds
