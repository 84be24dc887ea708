// cds_tiff.h -- TIFF container parsing on the host and the strip table the device decoder consumes.
#ifndef CDS_TIFF_H
#define CDS_TIFF_H

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/cdsgpu.h"
#include "cds_kernels.cuh"

namespace cds {

// One strip of one image, as the decode kernel sees it: `src` = first byte of the strip relative to the start of the uploaded
// byte range, `dst` = first decoded byte relative to the start of the chunk's RGB area (tiff_collect_strips with whole_rows, the fused
// ingest: first decoded image ROW, image index * height + row, so that a chunk may hold more than 4 GB of pixels).  Bit 31 of dst_len: the strip is
// PackBits-compressed (otherwise stored bytes are copied).
struct TiffStrip {
    uint32_t src, src_len, dst, dst_len;
};
constexpr uint32_t kTiffStripPacked = 0x80000000u;
constexpr uint32_t kTiffStoredPiece = 8192;      // stored (uncompressed) strips are cut into pieces of this many bytes
// most table entries one W x H image can need: a PackBits strip holds at least one row
inline size_t tiff_strips_bound(int W, int H) { return (size_t) H + (size_t) W * H * 3 / kTiffStoredPiece + 2; }

// Reads the header and the first IFD of `file` (classic TIFF, either byte order).  Fills `info`; when strip_off / strip_len are
// given, also the strip table (file-relative offsets, byte counts).  Returns CDS_OK, or CDS_ERR_BAD_ARG with `err` set for a
// buffer that is not a well-formed TIFF.  Whether the device decoder takes the file is info.decodable (`why` says why not).
cds_status tiff_parse(const uint8_t *file, size_t len, cds_tiff_info &info, std::vector<uint64_t> *strip_off,
                      std::vector<uint64_t> *strip_len, std::string &err, std::string *why = nullptr);

// Appends the strips of file `file` (located `src_base` bytes into the uploaded range, decoding to `dst_base`) to `out`;
// checks that the image is width x height and decodable.  On failure `err` describes the problem.
// whole_rows: stored (uncompressed) strips are cut into pieces of whole rows instead of kTiffStoredPiece bytes.
cds_status tiff_collect_strips(const uint8_t *file, size_t len, int width, int height, uint64_t src_base, uint64_t dst_base,
                               std::vector<TiffStrip> &out, std::string &err, bool whole_rows = false);

}  // namespace cds
struct cds_ctx;
namespace cds {
// host decoders (cds_formats.cpp)
cds_status tiff_decode_host(const uint8_t *file, size_t len, int width, int height, uint8_t *out_rgb, std::string &err);
cds_status png_inflate(const uint8_t *file, size_t len, int width, int height, int *bit_depth_out, uint8_t *out, size_t out_cap, std::string &err);
// PNG scanline filters undone on the device: image i's inflated stream at filtered + i * stride, bytes_per_sample[i] in {1, 2};
// out = uint16[n][height][width]
cds_status png_inflate_many(cds_ctx *ctx, const char *who, const uint8_t *blob, const int64_t *offsets, const int64_t *which, int64_t cnt,
                            int W, int H, uint8_t *h_filtered, size_t stride, uint8_t *bps);
// Device inflate (cds_inflate.cu): job i = the deflate data of image i (zlib header skipped) inside the uploaded byte range; out as
// for png_inflate_many; status[i] = 0, or why the stream was refused (InflateStatus, 16 = not exactly the image's bytes).
struct InflateJob { uint32_t src, src_len; };
void launch_png_inflate(const uint8_t *comp, const InflateJob *jobs, int64_t n, uint8_t *out, size_t stride, const uint8_t *bytes_per_sample,
                        int W, int H, int32_t *status, cudaStream_t s);
// Host half of it: checks file `file` (a W x H grayscale PNG), appends its IDAT payloads to dst (capacity cap) and fills the job;
// *bps = bytes per sample.  CDS_ERR_* with `err` set for a file that is no such PNG.
cds_status png_collect_idat(const uint8_t *file, size_t len, int W, int H, uint8_t *dst, size_t cap, size_t base, size_t *used, InflateJob *job,
                            uint8_t *bps, std::string &err);
void launch_png_unfilter(const uint8_t *filtered, size_t stride, const uint8_t *bytes_per_sample, int64_t n, int width, int height,
                         uint16_t *out, cudaStream_t s);

// Uploads files [i0, i0 + cnt) of the blob and decodes them into d_rgb, everything on stream `s` (cds_ingest.cu); d_comp / d_strips
// must hold the chunk (ingest_bounds gives sizes that suffice for any `cnt` consecutive files).  `strips` is host scratch.
cds_status ingest_chunk(cds_ctx *ctx, const char *who, const uint8_t *blob, const int64_t *offsets, int64_t i0, int64_t cnt, int W, int H,
                        uint8_t *d_comp, size_t comp_cap, TiffStrip *d_strips, size_t strips_cap, uint8_t *d_rgb, cudaStream_t s,
                        std::vector<TiffStrip> &strips);
void ingest_bounds(const int64_t *offsets, int64_t n, int64_t cnt, int W, int H, size_t &comp_cap, size_t &strips_cap);

void launch_tiff_decode(const uint8_t *src, const TiffStrip *strips, int64_t n_strips, uint8_t *dst_rgb, cudaStream_t s);
// Fused ingest (cds_ingest.cu): the same strips straight to code words (planes, slot first_slot + image) and, when `valid` is
// given, the per-sector "can match" bits of every row (launch_occupancy with valid_ready).  Needs a strip table whose stored
// pieces are whole rows (tiff_collect_strips with whole_rows = true); every row of every image must be covered by a strip.
// `work_counter`: 8 bytes of device memory for the kernel's strip counter (cleared on the stream by the launcher)
void launch_tiff_encode(const uint8_t *src, const TiffStrip *strips, int64_t n_strips, uint32_t *planes, PlaneGeom g, int64_t first_slot,
                        const uint16_t *rank_tab, int data_threshold, uint32_t *valid, unsigned long long *work_counter, cudaStream_t s);

}  // namespace cds
#endif
