// cds_ingest.cu -- device side of the image ingest: PackBits / stored TIFF strips -> RGB pixels.
//
// Replaces, for targets that arrive as TIFF files, the JVM-side decode of the reference:
// ImageArrayUtils.readImageArrayRangeWithTiffReader + packBitsUncompress
// (colormipsearch-api/src/main/java/org/janelia/colormipsearch/imageprocessing/ImageArrayUtils.java:184-258).
// The host reads only the tags (cds_tiff.cpp); the strips cross PCIe as stored -- colour-depth MIPs are mostly black, so
// PackBits shrinks the 2 MB of a 1210x566 image to 65-255 kB -- and are expanded here, next to the encoder that consumes them.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "cds_runtime.h"
#include "cds_tiff.h"

using namespace cds;

#define CDS_TRY(expr) do { cds_status _s = (expr); if (_s != CDS_OK) return _s; } while (0)
#define CDS_CUDA(ctx, expr) CDS_TRY((ctx)->check((expr), #expr))

namespace {

constexpr int kDecodeWarps = 8;
constexpr int kDecodeBuf = 2048;       // bytes of decoded output a warp collects in shared memory before it writes them out
constexpr int kDecodeSlack = 128;      // the longest run: a run that starts inside the buffer always fits

// Writes the first `n_buf` bytes of a warp's buffer to global memory.  Buffer byte b is strip byte `first + b` (first may be
// negative: the buffer window is aligned to 16 bytes in GLOBAL memory, so up to 15 bytes in front of the strip's first byte
// and behind its last one belong to the neighbouring strips and must not be touched).  Whole 16-byte vectors inside the strip
// go out as one 128-bit store per lane, fully coalesced; the ragged ends byte by byte.
__device__ __forceinline__ void flush_window(const uint8_t *sb, uint8_t *gout, int first, uint32_t n_buf, uint32_t out_len, uint32_t lane)
{
    __syncwarp();
    for (uint32_t v = lane; v * 16 < n_buf; v += 32) {
        const int lo = first + (int) (v * 16);
        if (lo >= 0 && lo + 16 <= (int) out_len && v * 16 + 16 <= n_buf) {
            *reinterpret_cast<uint4 *>(gout + lo) = *reinterpret_cast<const uint4 *>(sb + v * 16);
        } else {
            for (uint32_t b = 0; b < 16 && v * 16 + b < n_buf; b++) {
                const int o = lo + (int) b;
                if (o >= 0 && o < (int) out_len) gout[o] = sb[v * 16 + b];
            }
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }

// One warp per strip.  PackBits is a chain of runs whose positions depend on every control byte before them, so a strip is
// inherently serial in its control bytes -- but a chunk of 1 024 images has ~73 000 strips, enough to keep every SM's warp
// slots full, and a colour-depth MIP is mostly black: of its ~30 000 runs most are 128-byte fills with the same value that
// follow each other.  What counts is the number of instructions per run, so every iteration of the loop below takes as many
// CONSECUTIVE EQUAL-VALUE FILL RUNS as it can see at once: lane i reads the byte pair at input position idx + 2 i, a ballot
// finds how many leading pairs are fill runs of the first pair's value, one warp reduction adds their lengths, and the
// whole stretch (a black image row is 29 runs) is written as a single fill.  Literal runs are copied 32 bytes per step.
// Output is collected in a per-warp shared-memory window aligned with global memory and leaves the SM as full 128-bit
// stores (byte stores straight to global memory made every run a handful of partial-sector writes).
// (A lane-per-strip variant, one run state machine per lane, was 2.4 x slower: a third of all 4-byte output words contain
// a run boundary of SOME lane, so the warp lived in the divergent slow path with ~8 of 32 lanes active.)
// Like the reference's loop (ImageArrayUtils.java:229-258): 0..127 = that many + 1 literal bytes, -127..-1 = repeat the next
// byte 1 - n times, -128 = no-op; decoding stops at the end of the strip's input or of its rows, and whatever was not
// produced stays 0 (the Java array is zero-initialised).  Stored (uncompressed) pieces take the literal path.
__global__ void __launch_bounds__(kDecodeWarps * 32)
tiff_decode_kernel(const uint8_t *__restrict__ src, const TiffStrip *__restrict__ strips, int64_t n_strips, uint8_t *__restrict__ dst)
{
    __shared__ __align__(16) uint8_t s_buf[kDecodeWarps][kDecodeBuf + kDecodeSlack];
    const uint32_t lane = threadIdx.x & 31;
    const int64_t w = (int64_t) blockIdx.x * kDecodeWarps + (threadIdx.x >> 5);
    if (w >= n_strips) return;
    uint8_t *sb = s_buf[threadIdx.x >> 5];
    const uint32_t sba = (uint32_t) __cvta_generic_to_shared(sb);
    const TiffStrip st = strips[w];
    const uint8_t *__restrict__ in = src + st.src;
    const uint32_t in_len = st.src_len, out_len = st.dst_len & ~kTiffStripPacked;
    const bool packed = (st.dst_len & kTiffStripPacked) != 0;
    uint8_t *gout = dst + st.dst;
    const uint32_t mis = (uint32_t) (reinterpret_cast<uintptr_t>(gout) & 15);     // the window starts at the 16-byte boundary below gout

    uint32_t pos = 0, idx = 0;         // strip-relative output / input positions
    uint32_t flushed = 0;              // window bytes already written out (a multiple of kDecodeBuf)
    while (pos < out_len) {
        uint32_t cnt, fill = 0, from = 0;
        bool copy = false;
        if (idx >= in_len) {
            cnt = out_len - pos;                                   // input used up: zeros to the end of the strip
        } else if (!packed) {
            cnt = min((uint32_t) kDecodeSlack, in_len - idx);
            copy = true; from = idx; idx += cnt;
        } else {
            // lane i looks at the pair (control, value) that starts at idx + 2 i
            const uint32_t q = idx + 2 * lane;
            const uint32_t c = q < in_len ? (uint32_t) in[q] : 0u;
            const uint32_t v = q + 1 < in_len ? (uint32_t) in[q + 1] : 0u;
            const uint32_t c0 = __shfl_sync(0xffffffffu, c, 0), v0 = __shfl_sync(0xffffffffu, v, 0);
            if (c0 < 128) { cnt = c0 + 1; copy = true; from = idx + 1; idx += 1 + cnt; }
            else if (c0 == 128) { idx += 1; continue; }
            else {
                const uint32_t same = __ballot_sync(0xffffffffu, c > 128 && v == v0 && q + 1 < in_len);
                const uint32_t k = same == 0xffffffffu ? 32u : (uint32_t) __ffs((int) ~same) - 1u;      // >= 1 unless the value byte is missing
                fill = v0;
                if (k == 0) { cnt = 257 - c0; fill = 0; idx += 2; }                                       // a fill whose value byte lies beyond the input
                else { cnt = __reduce_add_sync(0xffffffffu, lane < k ? 257 - c : 0u); idx += 2 * k; }
            }
        }
        uint32_t todo = min(cnt, out_len - pos);
        if (copy) {
            // a literal of at most 128 bytes: it fits behind any position inside the buffer
            const uint32_t bi = pos + mis - flushed;
            for (uint32_t i = lane; i < todo; i += 32) sts_u8(sba + bi + i, from + i < in_len ? (uint32_t) in[from + i] : 0u);
            pos += todo;
            todo = 0;
        }
        const uint32_t vw = fill * 0x01010101u;
        for (;;) {
            // a fill of `todo` bytes, as much as the buffer takes: aligned words [a, e), head bytes [bi, 4a), tail bytes [4e, end)
            if (todo) {
                const uint32_t bi = pos + mis - flushed;
                const uint32_t n = min(todo, (uint32_t) (kDecodeBuf + kDecodeSlack) - bi);
                const uint32_t a = (bi + 3) >> 2, e = (bi + n) >> 2;
                if (a < e) {
                    for (uint32_t x = a + lane; x < e; x += 32) sts_u32(sba + 4 * x, vw);
                    if (lane < 4 * a - bi) sts_u8(sba + bi + lane, fill);
                    if (lane < bi + n - 4 * e) sts_u8(sba + 4 * e + lane, fill);
                } else if (lane < n) {
                    sts_u8(sba + bi + lane, fill);
                }
                pos += n; todo -= n;
            }
            if (pos + mis - flushed < (uint32_t) kDecodeBuf) break;
            flush_window(sb, gout, (int) flushed - (int) mis, kDecodeBuf, out_len, lane);
            // carry the bytes beyond the window (at most kDecodeSlack) to its start
            const uint32_t over = pos + mis - flushed - kDecodeBuf;
            const uint32_t wv = lane * 4 < over ? *reinterpret_cast<const uint32_t *>(sb + kDecodeBuf + lane * 4) : 0u;
            __syncwarp();
            if (lane * 4 < over) *reinterpret_cast<uint32_t *>(sb + lane * 4) = wv;
            __syncwarp();              // the carried words may reach past `over`, where the next run writes
            flushed += kDecodeBuf;
        }
    }
    flush_window(sb, gout, (int) flushed - (int) mis, pos + mis - flushed, out_len, lane);
}

}  // namespace

void cds::launch_tiff_decode(const uint8_t *src, const TiffStrip *strips, int64_t n_strips, uint8_t *dst_rgb, cudaStream_t s)
{
    if (n_strips <= 0) return;
    const int64_t blocks = (n_strips + kDecodeWarps - 1) / kDecodeWarps;
    tiff_decode_kernel<<<(unsigned) blocks, kDecodeWarps * 32, 0, s>>>(src, strips, n_strips, dst_rgb);
}

namespace {

// Strip table of files [i0, i0 + cnt): sources relative to the first byte of file i0, destinations relative to image i0.
cds_status collect_chunk(cds_ctx *ctx, const char *who, const uint8_t *blob, const int64_t *offsets, int64_t i0, int64_t cnt,
                         int W, int H, std::vector<TiffStrip> &strips)
{
    strips.clear();
    const size_t img_bytes = (size_t) W * H * 3;
    const int64_t base = offsets[i0];
    std::string err;
    for (int64_t i = 0; i < cnt; i++) {
        const int64_t a = offsets[i0 + i], b = offsets[i0 + i + 1];
        if (a < 0 || b < a) return ctx->fail(CDS_ERR_BAD_ARG, std::string(who) + ": offsets must be non-decreasing");
        cds_status s = tiff_collect_strips(blob + a, (size_t) (b - a), W, H, (uint64_t) (a - base), (uint64_t) i * img_bytes, strips, err);
        if (s != CDS_OK) return ctx->fail(s, std::string(who) + ": file " + std::to_string(i0 + i) + ": " + err);
    }
    return CDS_OK;
}

}  // namespace

namespace cds {

// Uploads files [i0, i0 + cnt) and decodes them into d_rgb, everything on stream `s` (used by the library and the one-shot
// decoder; the streaming search has its own double-buffered version).  d_comp / d_strips must hold the chunk.
cds_status ingest_chunk(cds_ctx *ctx, const char *who, const uint8_t *blob, const int64_t *offsets, int64_t i0, int64_t cnt, int W, int H,
                        uint8_t *d_comp, size_t comp_cap, TiffStrip *d_strips, size_t strips_cap, uint8_t *d_rgb, cudaStream_t s,
                        std::vector<TiffStrip> &strips)
{
    CDS_TRY(collect_chunk(ctx, who, blob, offsets, i0, cnt, W, H, strips));
    const size_t bytes = (size_t) (offsets[i0 + cnt] - offsets[i0]);
    if (bytes > comp_cap || strips.size() > strips_cap) return ctx->fail(CDS_ERR_CAPACITY, std::string(who) + ": internal staging too small");
    CDS_CUDA(ctx, cudaMemcpyAsync(d_comp, blob + offsets[i0], bytes, cudaMemcpyHostToDevice, s));
    // the table is pageable host memory: the runtime stages it before the call returns, so `strips` may be reused
    CDS_CUDA(ctx, cudaMemcpyAsync(d_strips, strips.data(), strips.size() * sizeof(TiffStrip), cudaMemcpyHostToDevice, s));
    ctx->stats.h2d_bytes += (int64_t) bytes + (int64_t) (strips.size() * sizeof(TiffStrip));
    launch_tiff_decode(d_comp, d_strips, (int64_t) strips.size(), d_rgb, s);
    ctx->stats.kernel_launches++;
    CDS_CUDA(ctx, cudaGetLastError());
    return CDS_OK;
}

// upper bounds for the staging of up to `cnt` consecutive files of the blob
void ingest_bounds(const int64_t *offsets, int64_t n, int64_t cnt, int W, int H, size_t &comp_cap, size_t &strips_cap)
{
    int64_t worst = 0;
    for (int64_t i = 0; i < n; i += cnt) worst = std::max(worst, offsets[std::min(n, i + cnt)] - offsets[i]);
    comp_cap = (size_t) std::max<int64_t>(worst, 0) + 64;
    strips_cap = (size_t) cnt * tiff_strips_bound(W, H);
}

}  // namespace cds

extern "C" cds_status cds_tiff_decode_rgb(cds_ctx *ctx, const uint8_t *blob, const int64_t *offsets, int64_t n,
                                          int32_t width, int32_t height, uint8_t *out_rgb)
{
    return cds::abi_guard("cds_tiff_decode_rgb", [&]() -> cds_status {
        if (!ctx) { set_tls_error("cds_tiff_decode_rgb: NULL context"); return CDS_ERR_BAD_ARG; }
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        if (n < 0 || width <= 0 || height <= 0 || width > 16384 || height > 16384) return ctx->fail(CDS_ERR_BAD_ARG, "cds_tiff_decode_rgb: bad size");
        if (n == 0) return CDS_OK;
        if (!blob || !offsets || !out_rgb) return ctx->fail(CDS_ERR_BAD_ARG, "cds_tiff_decode_rgb: NULL argument");
        DevState &ds = ctx->devs[0];
        CDS_CUDA(ctx, cudaSetDevice(ds.dev));
        const size_t img_bytes = (size_t) width * height * 3;
        const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(64, (int64_t) ((size_t) 1 << 30) / (int64_t) img_bytes));
        size_t comp_cap, strips_cap;
        ingest_bounds(offsets, n, chunk, width, height, comp_cap, strips_cap);
        uint8_t *d_comp = nullptr, *d_rgb = nullptr;
        TiffStrip *d_strips = nullptr;
        auto release = [&]() { cudaStreamSynchronize(ds.stream); ds.pool.free(d_comp); ds.pool.free(d_rgb); ds.pool.free(d_strips); };
        struct Guard { std::function<void()> f; ~Guard() { f(); } } guard{release};
        CDS_CUDA(ctx, ds.pool.alloc((void **) &d_comp, comp_cap));
        CDS_CUDA(ctx, ds.pool.alloc((void **) &d_strips, strips_cap * sizeof(TiffStrip)));
        CDS_CUDA(ctx, ds.pool.alloc((void **) &d_rgb, (size_t) chunk * img_bytes));
        std::vector<TiffStrip> strips;
        for (int64_t i0 = 0; i0 < n; i0 += chunk) {
            const int64_t cnt = std::min(chunk, n - i0);
            CDS_TRY(ingest_chunk(ctx, "cds_tiff_decode_rgb", blob, offsets, i0, cnt, width, height, d_comp, comp_cap, d_strips, strips_cap, d_rgb, ds.stream, strips));
            CDS_CUDA(ctx, cudaMemcpyAsync(out_rgb + (size_t) i0 * img_bytes, d_rgb, (size_t) cnt * img_bytes, cudaMemcpyDeviceToHost, ds.stream));
            CDS_CUDA(ctx, cudaStreamSynchronize(ds.stream));
            ctx->stats.d2h_bytes += (int64_t) cnt * (int64_t) img_bytes;
        }
        return CDS_OK;
    });
}

extern "C" cds_status cds_library_add_tiff(cds_library *lib, const uint8_t *blob, const int64_t *offsets, int64_t n, int64_t *first_index)
{
    return cds::abi_guard("cds_library_add_tiff", [&]() -> cds_status {
        if (!lib) { set_tls_error("cds_library_add_tiff: NULL library"); return CDS_ERR_BAD_ARG; }
        cds_ctx *ctx = lib->ctx;
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        if (n < 0) return ctx->fail(CDS_ERR_BAD_ARG, "negative image count");
        if (n > 0 && (!blob || !offsets)) return ctx->fail(CDS_ERR_BAD_ARG, "cds_library_add_tiff: NULL argument");
        if (n == 0) return library_append(lib, 0, nullptr, first_index);
        const int W = lib->g.W, H = lib->g.H;
        // library_append hands out runs of at most kLibBlock images; bound the staging by any kLibBlock consecutive files
        size_t comp_cap = 0, strips_cap = 0;
        {
            int64_t worst = 0;
            for (int64_t i = 0; i < n; i++) worst = std::max(worst, offsets[std::min(n, i + kLibBlock)] - offsets[i]);
            comp_cap = (size_t) std::max<int64_t>(worst, 0) + 64;
            strips_cap = (size_t) kLibBlock * tiff_strips_bound(W, H);
        }
        const size_t D = ctx->devs.size();
        std::vector<uint8_t *> d_comp(D, nullptr);
        std::vector<TiffStrip *> d_strips(D, nullptr);
        auto release = [&]() {
            for (size_t d = 0; d < D; d++) {
                if (!d_comp[d] && !d_strips[d]) continue;
                cudaSetDevice(ctx->devs[d].dev);
                cudaStreamSynchronize(ctx->devs[d].stream);
                ctx->devs[d].pool.free(d_comp[d]);
                ctx->devs[d].pool.free(d_strips[d]);
            }
        };
        struct Guard { std::function<void()> f; ~Guard() { f(); } } guard{release};
        std::vector<TiffStrip> strips;
        return library_append(lib, n, [&](DevState &ds, int64_t i0, int64_t cnt, uint8_t *d_rgb) -> cds_status {
            const size_t d = (size_t) (&ds - ctx->devs.data());
            if (!d_comp[d]) {
                CDS_CUDA(ctx, ds.pool.alloc((void **) &d_comp[d], comp_cap));
                CDS_CUDA(ctx, ds.pool.alloc((void **) &d_strips[d], strips_cap * sizeof(TiffStrip)));
            }
            return ingest_chunk(ctx, "cds_library_add_tiff", blob, offsets, i0, cnt, W, H, d_comp[d], comp_cap, d_strips[d], strips_cap, d_rgb, ds.stream, strips);
        }, first_index);
    });
}

extern "C" cds_status cds_maskset_add_tiff(cds_maskset *ms, const uint8_t *blob, const int64_t *offsets, int32_t n, int32_t *mask_size_out)
{
    return cds::abi_guard("cds_maskset_add_tiff", [&]() -> cds_status {
        if (!ms) { set_tls_error("cds_maskset_add_tiff: NULL mask set"); return CDS_ERR_BAD_ARG; }
        cds_ctx *ctx = ms->ctx;
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        if (n < 0 || (n > 0 && (!blob || !offsets))) return ctx->fail(CDS_ERR_BAD_ARG, "cds_maskset_add_tiff: bad arguments");
        if (n == 0) return CDS_OK;
        const int W = ms->W, H = ms->H;
        // maskset_append asks for runs of at most 64 masks; uploads and decodes are ordered on one stream, so one buffer does
        size_t comp_cap = 0, strips_cap = 0;
        ingest_bounds(offsets, n, 64, W, H, comp_cap, strips_cap);
        for (int64_t i = 0; i < n; i++) comp_cap = std::max(comp_cap, (size_t) std::max<int64_t>(offsets[std::min<int64_t>(n, i + 64)] - offsets[i], 0) + 64);
        DevState &d0 = ctx->devs[0];
        CDS_CUDA(ctx, cudaSetDevice(d0.dev));
        uint8_t *d_comp = nullptr;
        TiffStrip *d_strips = nullptr;
        auto release = [&]() { cudaStreamSynchronize(d0.copy_stream); cudaStreamSynchronize(d0.stream); d0.pool.free(d_comp); d0.pool.free(d_strips); };
        struct Guard { std::function<void()> f; ~Guard() { f(); } } guard{release};
        CDS_CUDA(ctx, d0.pool.alloc((void **) &d_comp, comp_cap));
        CDS_CUDA(ctx, d0.pool.alloc((void **) &d_strips, strips_cap * sizeof(TiffStrip)));
        std::vector<TiffStrip> strips;
        return maskset_append(ms, n, mask_size_out, [&](int i0, int cnt, uint8_t *stage, cudaStream_t stream) -> cds_status {
            return ingest_chunk(ctx, "cds_maskset_add_tiff", blob, offsets, i0, cnt, W, H, d_comp, comp_cap, d_strips, strips_cap, stage, stream, strips);
        });
    });
}
