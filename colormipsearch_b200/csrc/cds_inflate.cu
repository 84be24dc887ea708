// cds_inflate.cu -- the zlib streams of a window of PNG files inflated on the device, one warp per stream (cds_inflate.h).
// Input: the files' IDAT payloads, concatenated per file by the host and uploaded as stored; output: the filtered scanlines that
// png_unfilter_kernel (cds_ingest.cu) turns into pixels.  A stream the decoder refuses, or one that does not hold exactly the
// image's bytes, gets a non-zero status; the caller inflates that file on the host (zlib) and decides there.
#include "cds_inflate.h"
#include "cds_tiff.h"

namespace cds {
namespace {

constexpr int kInflateWarps = 4;      // streams per CTA: few, so that the warps of a window spread over all SMs

__global__ void __launch_bounds__(kInflateWarps * 32) png_inflate_kernel(const uint8_t *__restrict__ comp, const InflateJob *__restrict__ jobs, int64_t n,
                                                                         uint8_t *out, size_t stride, const uint8_t *__restrict__ bytes_per_sample,
                                                                         int W, int H, int32_t *__restrict__ status)
{
    __shared__ InflateTables tables[kInflateWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t) blockIdx.x * kInflateWarps + warp;
    if (i >= n) return;                                       // whole warps leave
    const InflateJob job = jobs[i];
    const size_t expect = (size_t) H * (1 + (size_t) W * bytes_per_sample[i]);
    size_t produced = 0;
    int st = job.src_len ? inflate_stream<32>(comp + job.src, job.src_len, out + (size_t) i * stride, expect, tables[warp], lane, &produced) : (int) kInfInputShort;
    // the image is the first `expect` bytes of the stream, as for a reader that stops when it has its rows (the host path does)
    if (st == kInfOutputFull && produced == expect) st = kInfOk;
    if (st == kInfOk && produced != expect) st = 16;          // a shorter image than the header states
    if (lane == 0) status[i] = st;
}

}  // namespace

void launch_png_inflate(const uint8_t *comp, const InflateJob *jobs, int64_t n, uint8_t *out, size_t stride, const uint8_t *bytes_per_sample,
                        int W, int H, int32_t *status, cudaStream_t s)
{
    if (n <= 0) return;
    png_inflate_kernel<<<(unsigned) ((n + kInflateWarps - 1) / kInflateWarps), kInflateWarps * 32, 0, s>>>(comp, jobs, n, out, stride, bytes_per_sample, W, H, status);
}

}  // namespace cds
