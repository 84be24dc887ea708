// cds_band.cu -- batched pixel match: target row bands staged in shared memory by TMA bulk copies, a group of masks
// streamed against each band, match counts kept as packed 16-bit lanes in registers.
//
// What it computes is exactly PixelMatchColorDepthSearchAlgorithm.calculateMatchingScore
// (colormipsearch-api/src/main/java/org/janelia/colormipsearch/cds/PixelMatchColorDepthSearchAlgorithm.java:166-263) for every
// (mask, target) of the launch: per shift/mirror variant the number of mask pixels whose shifted target pixel is inside the
// image, above the data threshold and within zTolerance on the colour-depth scale; max over the unmirrored variants, max over
// the mirrored ones, mirrored wins only when strictly greater.
//
// How (B200):
//   * work item = (mask group of kGroup masks, target).  Items are ordered group-major so that all CTAs stream the SAME
//     group's pixel records (L2-resident, ~10 MB) while each target plane is read from HBM once per group.
//   * a persistent grid of one 512-thread CTA per SM pulls items from a global counter.  The target is consumed as
//     bands of R rows (+ s = xyShift halo rows on both sides).  A band of a plane is ONE contiguous span of HBM
//     (cds_kernels.cuh PlaneGeom), fetched by a single cp.async.bulk (TMA bulk copy, SASS UBLKCP) that completes on an
//     mbarrier; two stages double-buffer the copy of band b+2 (or of the next item's first bands) behind the compute of b, b+1.
//   * out-of-image pixels of shifted / mirrored variants need no bounds test: they land in the pad columns / guard rows of
//     the plane, whose code words can never match.
//   * inside a band, warps grab (mask, band) segments from a shared counter; a lane owns one mask pixel per iteration:
//     one LDG.128 for the record, one LDS per variant, 4 integer ops for the two-interval test, one predicated add into a
//     packed counter.  A segment ends with one REDUX per packed register and plain adds into per-(mask, variant) accumulators.
#include "cds_band.cuh"

#include <cstdlib>

namespace cds {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kGroup = 64;          // masks per work item
constexpr int kStages = 2;
constexpr int kMaxBands = 128;
constexpr int kPrePad = 4;          // words (16 bytes, keeps the bulk-copy destination aligned)

struct BandParams {
    const MaskDesc *masks;
    int n_masks;
    const uint32_t *planes;
    PlaneGeom g;
    int64_t n_targets;
    int32_t *scores;                // [n_masks][n_targets]
    unsigned long long *work_counter;
    int rows_per_band;              // R
    int n_bands;
    int stage_words;                // (R + 2s) * pitch
    int n_groups;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}

// shift offsets of ring structure NRINGS (0: none, 1: +-2, 2: +-2 and +-4), in the order of the maskset's ShiftSet
template <int NRINGS> struct Offsets;
template <> struct Offsets<0> { static constexpr int N = 1; };
template <> struct Offsets<1> { static constexpr int N = 9; };
template <> struct Offsets<2> { static constexpr int N = 17; };

template <int NRINGS>
__device__ __forceinline__ void offset_of(int v, int &dx, int &dy)
{
    // v is a compile-time constant after unrolling
    if (NRINGS == 0) { dx = 0; dy = 0; return; }
    if (v < 9) { dx = (v / 3 - 1) * 2; dy = (v % 3 - 1) * 2; return; }
    // ring 4 without its centre: xx in {-4,0,4} x yy in {-4,0,4} minus (0,0)
    int k = v - 9;
    if (k >= 4) k++;
    dx = (k / 3 - 1) * 4;
    dy = (k % 3 - 1) * 4;
}

template <int NRINGS, bool MIRROR>
__global__ void __launch_bounds__(kThreads, 1) pixelmatch_band_kernel(const BandParams p)
{
    constexpr int NS = Offsets<NRINGS>::N;            // variants per orientation
    constexpr int NV = MIRROR ? 2 * NS : NS;          // variants
    constexpr int NREG = (NV + 1) / 2;                // packed 16-bit counter registers
    constexpr int S = 2 * NRINGS;                     // halo rows = xyShift

    extern __shared__ __align__(128) unsigned char smem_raw[];
    // every stage is preceded by kPrePad never-matching words: a pixel in the first row of a band whose shifted / mirrored
    // column is -1 or -2 reads just below the stage
    uint32_t *s_stage = reinterpret_cast<uint32_t *>(smem_raw) + kPrePad;             // kStages * (stage_words + kPrePad)
    const int stage_stride = p.stage_words + kPrePad;
    int *s_acc = reinterpret_cast<int *>(s_stage - kPrePad + (size_t) kStages * stage_stride); // [kGroup][NV]
    uint32_t *s_seg = reinterpret_cast<uint32_t *>(s_acc + kGroup * NV);              // [kGroup][n_bands + 1]
    const cds_mask_record **s_rec = reinterpret_cast<const cds_mask_record **>(s_seg + kGroup * (p.n_bands + 1));
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_rec + kGroup);   // kStages mbarriers
    int *s_next = reinterpret_cast<int *>(s_bar + kStages);                           // [2] segment tickets
    long long *s_work = reinterpret_cast<long long *>(s_next + 2);                    // [2] current / next item

    const int tid = threadIdx.x, lane = tid & 31;
    const int pitch = p.g.pitch, W = p.g.W, H = p.g.H, R = p.rows_per_band;
    const long long n_items = (long long) p.n_groups * p.n_targets;

    if (tid == 0) {
        for (int s = 0; s < kStages; s++) mbar_init(smem_u32(s_bar + s), 1);
        s_next[0] = 0; s_next[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        s_work[0] = (long long) atomicAdd(p.work_counter, 1ull);
        s_work[1] = (long long) atomicAdd(p.work_counter, 1ull);
    }
    for (int i = tid; i < kGroup * NV; i += kThreads) s_acc[i] = 0;
    if (tid < kStages * kPrePad) s_stage[(tid / kPrePad) * stage_stride - kPrePad + (tid % kPrePad)] = CDS_CODE_PAD_WORD;
    __syncthreads();

    // issue the load of band `b` of item `w` into stage `stage` (thread 0 only)
    auto issue_load = [&](long long w, int b, int stage) {
        const int64_t t = w % p.n_targets;
        const int y0 = b * R;
        const int y1 = min(y0 + R, H);
        const uint32_t bytes = (uint32_t) ((y1 - y0 + 2 * S) * pitch) * 4u;
        const uint32_t *src = p.planes + p.g.row_offset(t, 0) + (long long) (y0 - S) * pitch;   // guard rows cover y0 - S < 0
        const uint32_t bar = smem_u32(s_bar + stage);
        mbar_expect_tx(bar, bytes);
        bulk_load(smem_u32(s_stage + (size_t) stage * stage_stride), src, bytes, bar);
    };

    long long w = s_work[0];
    long long w_next = s_work[1];
    uint32_t seq = 0;                     // running band number across items: stage = seq & 1, parity = (seq >> 1) & 1
    if (tid == 0 && w < n_items) {
        issue_load(w, 0, 0);
        if (p.n_bands > 1) issue_load(w, 1, 1);
        else if (w_next < n_items) issue_load(w_next, 0, 1);
    }

    while (w < n_items) {
        const int gi = (int) (w / p.n_targets);
        const int64_t t = w % p.n_targets;
        const int m0 = gi * kGroup;
        const int mb = min(kGroup, p.n_masks - m0);

        // per-item tables: record base pointers and band boundaries of every mask of the group
        for (int i = tid; i < mb; i += kThreads) s_rec[i] = p.masks[m0 + i].records;
        for (int i = tid; i < mb * (p.n_bands + 1); i += kThreads) {
            const int mi = i / (p.n_bands + 1), b = i % (p.n_bands + 1);
            s_seg[i] = __ldg(p.masks[m0 + mi].rowstart + min(b * R, H));
        }
        __syncthreads();

        for (int b = 0; b < p.n_bands; b++, seq++) {
            const int stage = seq & 1;
            mbar_wait(smem_u32(s_bar + stage), (seq >> 1) & 1);
            const uint32_t *band = s_stage + (size_t) stage * stage_stride;
            const int y0 = b * R;

            // warps pull (mask, band) segments
            for (;;) {
                int mi = 0;
                if (lane == 0) mi = atomicAdd(&s_next[b & 1], 1);
                mi = __shfl_sync(0xffffffffu, mi, 0);
                if (mi >= mb) break;
                const uint32_t seg0 = s_seg[mi * (p.n_bands + 1) + b];
                const uint32_t seg1 = s_seg[mi * (p.n_bands + 1) + b + 1];
                if (seg0 == seg1) continue;
                const cds_mask_record *rec = s_rec[mi];
                uint32_t cnt[NREG];
#pragma unroll
                for (int j = 0; j < NREG; j++) cnt[j] = 0;

                uint32_t i = seg0 + lane;
                uint4 q = make_uint4(0, 0, 0, 0);
                if (i < seg1) q = __ldg(reinterpret_cast<const uint4 *>(rec + i));
                while (i < seg1) {
                    const uint4 cur = q;
                    const uint32_t inext = i + 32;
                    if (inext < seg1) q = __ldg(reinterpret_cast<const uint4 *>(rec + inext));   // prefetch behind the compute
                    const int x = (int) (cur.x & 0xFFFFu);
                    const int row = (int) (cur.x >> 16) - y0 + S;
                    const uint32_t lo1 = cur.y, lo2 = cur.z;
                    const uint32_t len1 = ((cur.w & 0xFFFFu) << CDS_CODE_SR_SHIFT) | 0xFFu;
                    const uint32_t len2 = ((cur.w >> 16) << CDS_CODE_SR_SHIFT) | 0xFFu;
                    const uint32_t *pc = band + row * pitch + x;                 // unmirrored centre
                    const uint32_t *pm = band + row * pitch + (W - 1 - x);       // mirrored centre
#pragma unroll
                    for (int v = 0; v < NS; v++) {
                        int dx, dy;
                        offset_of<NRINGS>(v, dx, dy);
                        const uint32_t c = pc[dy * pitch + dx];
                        const bool hit = (c - lo1 <= len1) | (c - lo2 <= len2);
                        if (hit) cnt[v >> 1] += (v & 1) ? 0x10000u : 1u;
                    }
                    if (MIRROR) {
#pragma unroll
                        for (int v = 0; v < NS; v++) {
                            int dx, dy;
                            offset_of<NRINGS>(v, dx, dy);
                            const uint32_t c = pm[dy * pitch - dx];              // mirror of (x + dx) is (W-1-x) - dx
                            const bool hit = (c - lo1 <= len1) | (c - lo2 <= len2);
                            if (hit) cnt[(NS + v) >> 1] += ((NS + v) & 1) ? 0x10000u : 1u;
                        }
                    }
                    i = inext;
                }
                // segment done: warp totals (a segment has < 65536 pixels, so the packed halves cannot carry)
                uint32_t mine0 = 0, mine1 = 0;
#pragma unroll
                for (int j = 0; j < NREG; j++) {
                    const uint32_t tot = __reduce_add_sync(0xffffffffu, cnt[j]);
                    if ((lane >> 1) == j) mine0 = tot;
                    if (((lane + 32) >> 1) == j) mine1 = tot;
                }
                if (lane < NV) s_acc[mi * NV + lane] += (int) ((lane & 1) ? (mine0 >> 16) : (mine0 & 0xFFFFu));
                if (NV > 32 && lane + 32 < NV) s_acc[mi * NV + lane + 32] += (int) ((lane & 1) ? (mine1 >> 16) : (mine1 & 0xFFFFu));
            }
            __syncthreads();      // everyone is done with this stage and with this band's tickets
            if (tid == 0) {
                s_next[b & 1] = 0;
                // refill this stage with the band two ahead: of this item, or of the next one
                if (b + 2 < p.n_bands) issue_load(w, b + 2, stage);
                else if (w_next < n_items) {
                    const int nb = b + 2 - p.n_bands;
                    if (nb < p.n_bands) issue_load(w_next, nb, stage);
                }
            }
        }

        // item epilogue: max over variants per orientation; mirrored wins only when strictly greater
        for (int mi = tid; mi < mb; mi += kThreads) {
            int best = 0, bestm = 0;
#pragma unroll
            for (int v = 0; v < NS; v++) best = max(best, s_acc[mi * NV + v]);
            if (MIRROR) {
#pragma unroll
                for (int v = 0; v < NS; v++) bestm = max(bestm, s_acc[mi * NV + NS + v]);
            }
            int word = best;
            if (MIRROR && bestm > best) word = bestm | CDS_SCORE_MIRROR_BIT;
            p.scores[(size_t) (m0 + mi) * p.n_targets + t] = word;
#pragma unroll
            for (int v = 0; v < NV; v++) s_acc[mi * NV + v] = 0;
        }
        if (tid == 0) {
            s_work[0] = w_next;
            s_work[1] = (long long) atomicAdd(p.work_counter, 1ull);
        }
        __syncthreads();
        w = s_work[0];
        w_next = s_work[1];
        // a single-band image never prefetched the item after next: top it up
        if (p.n_bands == 1 && tid == 0 && w < n_items && w_next < n_items) issue_load(w_next, 0, (seq + 1) & 1);
        __syncthreads();
    }
}

struct BandConfig {
    int rows_per_band, n_bands, stage_words;
    size_t smem_bytes;
    bool ok;
};

BandConfig band_config(int xy_shift, bool mirror, const PlaneGeom &g)
{
    BandConfig c{};
    const int S = xy_shift;
    const int NS = xy_shift == 0 ? 1 : (xy_shift == 2 ? 9 : 17);
    const int NV = mirror ? 2 * NS : NS;
    const size_t budget = 227 * 1024;
    // fixed part: accumulators, record pointers, barriers, tickets, work slots (+ slack for the seg table, solved below)
    for (int R = (int) (65535 / g.W); R >= 1; R--) {
        int n_bands = (g.H + R - 1) / R;
        if (n_bands > kMaxBands) break;
        size_t stage_words = (size_t) (R + 2 * S) * g.pitch;
        size_t bytes = kStages * (stage_words + kPrePad) * 4 + (size_t) kGroup * NV * 4 + (size_t) kGroup * (n_bands + 1) * 4 +
                       (size_t) kGroup * 8 + kStages * 8 + 2 * 4 + 2 * 8 + 64;
        if (bytes <= budget && stage_words * 4 < (1u << 20)) {
            c.rows_per_band = R; c.n_bands = n_bands; c.stage_words = (int) stage_words; c.smem_bytes = bytes; c.ok = true;
            return c;
        }
    }
    c.ok = false;
    return c;
}

unsigned long long *g_work_counter[64] = {nullptr};

}  // namespace

bool band_kernel_supported(int xy_shift, const PlaneGeom &g)
{
    static const bool disabled = std::getenv("CDSGPU_DISABLE_BAND") != nullptr;
    if (disabled) return false;
    if (!(xy_shift == 0 || xy_shift == 2 || xy_shift == 4)) return false;
    if (xy_shift > g.guard || xy_shift > g.pitch - g.W) return false;
    return band_config(xy_shift, true, g).ok;
}

int band_min_masks()
{
    static const int v = [] { const char *e = std::getenv("CDSGPU_BAND_MIN_MASKS"); return e ? std::atoi(e) : 16; }();
    return v;
}

int launch_pixelmatch_band(const MaskDesc *masks, int n_masks, const uint32_t *planes, PlaneGeom g, int64_t n_targets,
                           int xy_shift, bool mirror, int32_t *scores, cudaStream_t s)
{
    if (n_masks == 0 || n_targets == 0) return 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64) return 0;
    if (!g_work_counter[dev]) {
        if (cudaMalloc(&g_work_counter[dev], sizeof(unsigned long long)) != cudaSuccess) return 0;
    }
    cudaMemsetAsync(g_work_counter[dev], 0, sizeof(unsigned long long), s);
    BandConfig c = band_config(xy_shift, mirror, g);
    if (!c.ok) return 0;
    BandParams p;
    p.masks = masks; p.n_masks = n_masks; p.planes = planes; p.g = g; p.n_targets = n_targets; p.scores = scores;
    p.work_counter = g_work_counter[dev];
    p.rows_per_band = c.rows_per_band; p.n_bands = c.n_bands; p.stage_words = c.stage_words;
    p.n_groups = (n_masks + kGroup - 1) / kGroup;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    long long n_items = (long long) p.n_groups * n_targets;
    int grid = (int) std::min<long long>(n_sm, n_items);
    void (*kern)(const BandParams) = nullptr;
    const int rings = xy_shift / 2;
    if (rings == 0) kern = mirror ? pixelmatch_band_kernel<0, true> : pixelmatch_band_kernel<0, false>;
    else if (rings == 1) kern = mirror ? pixelmatch_band_kernel<1, true> : pixelmatch_band_kernel<1, false>;
    else kern = mirror ? pixelmatch_band_kernel<2, true> : pixelmatch_band_kernel<2, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c.smem_bytes);
    kern<<<grid, kThreads, c.smem_bytes, s>>>(p);
    return 1;
}

}  // namespace cds
