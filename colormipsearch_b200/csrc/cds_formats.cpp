// cds_formats.cpp -- host side of the remaining image ingest: what the reference reads that is not a PackBits
// or stored RGB TIFF.
//
//   * LZW TIFF.  The reference hands every TIFF that is not PackBits to ImageJ's Opener
//     (colormipsearch-api/src/main/java/org/janelia/colormipsearch/imageprocessing/ImageArrayUtils.java:176-182, 196-198): a decode
//     on the JVM.  cds_tiff_decode_rgb_host decodes any 8-bit RGB strip TIFF the parser understands (none, PackBits, LZW with or
//     without the horizontal predictor) on the host, and cds_tiff_to_packbits rewrites it as the PackBits TIFF the device path takes.
//   * 16-bit PNG.  Gradient images are 16-bit grayscale PNG files read through ImageIO.read (ImageArrayUtils.java:98-121,176-178).
//     A PNG is a zlib stream of filtered scanlines: the stream is inflated here (zlib, one image per host thread), the filter
//     reconstruction and the byte swap happen on the device (png_unfilter_kernel in cds_ingest.cu).
//   * zip archives.  Libraries are routinely zip files whose entries are the MIPs
//     (colormipsearch-api/src/main/java/org/janelia/colormipsearch/mips/NeuronMIPUtils.java:124-129, 177-227): the central directory
//     is read here; stored entries are used where they lie, deflated ones are inflated.  Entry lookup follows the reference: the
//     exact name, else the first entry with the same file name (:193-208).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <thread>

#include <zlib.h>

#include <memory>

#include "cds_inflate.h"
#include "cds_runtime.h"
#include "cds_tiff.h"

using namespace cds;

namespace cds {

// ------------------------------------------------------------------------------------------------------------------ TIFF LZW
// TIFF 6.0 section 13: MSB-first codes of 9..12 bits, ClearCode 256, EndOfInformation 257, "early change" (the code width grows one
// code before the table is full).  Decodes one strip into out[0 .. out_len); returns the number of bytes produced.
static size_t lzw_decode_strip(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
{
    struct Entry { uint16_t prefix; uint8_t ch; uint16_t len; };
    std::vector<Entry> tab(4096);
    for (int i = 0; i < 256; i++) tab[i] = Entry{0xFFFF, (uint8_t) i, 1};
    int next_code = 258, width = 9;
    uint32_t acc = 0;
    int bits = 0;
    size_t ip = 0, op = 0;
    int prev = -1;
    auto emit = [&](int code) -> uint8_t {            // writes the string of `code` at op, returns its first character
        const int len = tab[code].len;
        uint8_t first = 0;
        int c = code;
        for (int k = len - 1; k >= 0; k--) {
            if (op + (size_t) k < out_len) out[op + k] = tab[c].ch;
            first = tab[c].ch;
            c = tab[c].prefix;
        }
        op += (size_t) len;
        return first;
    };
    while (op < out_len) {
        while (bits < width && ip < in_len) { acc = (acc << 8) | in[ip++]; bits += 8; }
        if (bits < width) break;
        const int code = (int) ((acc >> (bits - width)) & ((1u << width) - 1u));
        bits -= width;
        if (code == 257) break;
        if (code == 256) { next_code = 258; width = 9; prev = -1; continue; }
        if (prev < 0) {
            if (code >= 256) break;                   // malformed: the first code after a clear must be a literal
            emit(code);
            prev = code;
            continue;
        }
        if (code < next_code) {
            const uint8_t first = emit(code);
            if (next_code < 4096) tab[next_code++] = Entry{(uint16_t) prev, first, (uint16_t) (tab[prev].len + 1)};
        } else if (code == next_code && next_code < 4096) {
            // the string that is being defined: prev + first(prev)
            int c = prev;
            while (tab[c].prefix != 0xFFFF) c = tab[c].prefix;
            tab[next_code] = Entry{(uint16_t) prev, tab[c].ch, (uint16_t) (tab[prev].len + 1)};
            next_code++;
            emit(code);
        } else {
            break;                                    // malformed
        }
        prev = code;
        if (next_code + 1 >= (1 << width) && width < 12) width++;      // early change
    }
    return std::min(op, out_len);
}

static size_t packbits_decode_strip(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
{
    // the same clamped rules as the device decoders (cds_ingest.cu) and the reference's packBitsUncompress (ImageArrayUtils.java:229-258)
    size_t idx = 0, pos = 0;
    while (idx < in_len && pos < out_len) {
        const uint8_t c = in[idx];
        if (c < 128) {
            const size_t cnt = (size_t) c + 1;
            for (size_t i = 0; i < cnt && pos + i < out_len; i++) out[pos + i] = idx + 1 + i < in_len ? in[idx + 1 + i] : 0;
            idx += 1 + cnt; pos += cnt;
        } else if (c != 128) {
            const size_t cnt = 257 - (size_t) c;
            const uint8_t v = idx + 1 < in_len ? in[idx + 1] : 0;
            for (size_t i = 0; i < cnt && pos + i < out_len; i++) out[pos + i] = v;
            idx += 2; pos += cnt;
        } else {
            idx += 1;
        }
    }
    return std::min(pos, out_len);
}

// reads tag 317 (Predictor) of the first IFD; 1 when absent
static int tiff_predictor(const uint8_t *file, size_t len)
{
    if (len < 8) return 1;
    const bool be = file[0] == 'M';
    auto u16 = [&](size_t o) -> uint32_t { return be ? (uint32_t) file[o] << 8 | file[o + 1] : (uint32_t) file[o + 1] << 8 | file[o]; };
    auto u32 = [&](size_t o) -> uint32_t { return be ? u16(o) << 16 | u16(o + 2) : u16(o + 2) << 16 | u16(o); };
    const size_t ifd = u32(4);
    if (ifd + 2 > len) return 1;
    const uint32_t n = u16(ifd);
    for (uint32_t e = 0; e < n; e++) {
        const size_t at = ifd + 2 + (size_t) e * 12;
        if (at + 12 > len) break;
        if (u16(at) == 317) return (int) (u16(at + 2) == 3 ? u16(at + 8) : u32(at + 8));
    }
    return 1;
}

cds_status tiff_decode_host(const uint8_t *file, size_t len, int width, int height, uint8_t *out_rgb, std::string &err)
{
    cds_tiff_info info;
    std::vector<uint64_t> offs, lens;
    std::string why;
    cds_status s = tiff_parse(file, len, info, &offs, &lens, err, &why);
    if (s != CDS_OK) return s;
    if (info.width != width || info.height != height) {
        char buf[160];
        snprintf(buf, sizeof buf, "Invalid image size - TIFF image size (%d, %d) must match (%d, %d)", info.width, info.height, width, height);
        err = buf;
        return CDS_ERR_SIZE_MISMATCH;
    }
    const bool lzw = info.compression == 5;
    if (!info.decodable && !(lzw && why.rfind("compression", 0) == 0)) { err = "TIFF not decodable: " + why; return CDS_ERR_UNSUPPORTED; }
    if (lzw) {
        // the checks tiff_parse skips once it has seen an unsupported compression
        if (info.samples_per_pixel != 3 || info.bits_per_sample != 8 || info.planar_config != 1 || info.photometric != 2) { err = "LZW TIFF is not 8-bit chunky RGB"; return CDS_ERR_UNSUPPORTED; }
        const int64_t expect = ((int64_t) info.height + info.rows_per_strip - 1) / info.rows_per_strip;
        if ((int64_t) info.n_strips != expect) { err = "strip count does not match RowsPerStrip"; return CDS_ERR_UNSUPPORTED; }
    }
    const int predictor = tiff_predictor(file, len);
    if (predictor != 1 && predictor != 2) { err = "TIFF predictor " + std::to_string(predictor) + " is not supported"; return CDS_ERR_UNSUPPORTED; }
    const size_t row_bytes = (size_t) width * 3;
    memset(out_rgb, 0, row_bytes * height);
    for (int i = 0; i < info.n_strips; i++) {
        const size_t row0 = (size_t) i * info.rows_per_strip;
        const size_t rows = std::min<size_t>(info.rows_per_strip, (size_t) height - row0);
        uint8_t *dst = out_rgb + row0 * row_bytes;
        const size_t dlen = rows * row_bytes;
        const uint8_t *src = file + offs[i];
        if (info.compression == 1) memcpy(dst, src, std::min<size_t>(lens[i], dlen));
        else if (info.compression == 32773) packbits_decode_strip(src, lens[i], dst, dlen);
        else lzw_decode_strip(src, lens[i], dst, dlen);
        if (predictor == 2)                           // horizontal differencing, per sample, restarting at every row (TIFF 6.0 section 14)
            for (size_t r = 0; r < rows; r++) {
                uint8_t *row = dst + r * row_bytes;
                for (size_t b = 3; b < row_bytes; b++) row[b] = (uint8_t) (row[b] + row[b - 3]);
            }
    }
    return CDS_OK;
}

// ------------------------------------------------------------------------------------------------------------------ PNG
static uint32_t be32(const uint8_t *p) { return (uint32_t) p[0] << 24 | (uint32_t) p[1] << 16 | (uint32_t) p[2] << 8 | p[3]; }

cds_status png_parse(const uint8_t *file, size_t len, cds_png_info &info, std::vector<std::pair<size_t, size_t>> *idat, std::string &err)
{
    memset(&info, 0, sizeof info);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (!file || len < 8 + 25 || memcmp(file, sig, 8) != 0) { err = "not a PNG file"; return CDS_ERR_BAD_ARG; }
    size_t pos = 8;
    bool have_hdr = false;
    while (pos + 12 <= len) {
        const size_t n = be32(file + pos);
        const uint8_t *type = file + pos + 4;
        if (n > len - pos - 12) { err = "PNG: chunk runs past the end of the file"; return CDS_ERR_BAD_ARG; }
        if (memcmp(type, "IHDR", 4) == 0 && n == 13) {
            info.width = (int32_t) std::min<uint32_t>(be32(file + pos + 8), INT32_MAX);
            info.height = (int32_t) std::min<uint32_t>(be32(file + pos + 12), INT32_MAX);
            info.bit_depth = file[pos + 16]; info.color_type = file[pos + 17];
            info.interlace = file[pos + 20];
            have_hdr = true;
        } else if (memcmp(type, "IDAT", 4) == 0) {
            if (idat) idat->push_back({pos + 8, n});
            info.data_bytes += (int64_t) n;
        } else if (memcmp(type, "IEND", 4) == 0) {
            break;
        }
        pos += 12 + n;
    }
    if (!have_hdr || info.width <= 0 || info.height <= 0) { err = "PNG: no header"; return CDS_ERR_BAD_ARG; }
    // grayscale, 8 or 16 bits, not interlaced, compression / filter method 0
    info.decodable = (info.color_type == 0 && (info.bit_depth == 8 || info.bit_depth == 16) && info.interlace == 0) ? 1 : 0;
    return CDS_OK;
}

// inflates the IDAT stream of a grayscale PNG into filtered scanlines: height * (1 + width * bytes per pixel) bytes
cds_status png_inflate(const uint8_t *file, size_t len, int width, int height, int *bit_depth_out, uint8_t *out, size_t out_cap, std::string &err)
{
    cds_png_info info;
    std::vector<std::pair<size_t, size_t>> idat;
    cds_status s = png_parse(file, len, info, &idat, err);
    if (s != CDS_OK) return s;
    if (!info.decodable) { err = "PNG is not a non-interlaced 8- or 16-bit grayscale image"; return CDS_ERR_UNSUPPORTED; }
    if (info.width != width || info.height != height) {
        char buf[160];
        snprintf(buf, sizeof buf, "Invalid image size - PNG image size (%d, %d) must match (%d, %d)", info.width, info.height, width, height);
        err = buf;
        return CDS_ERR_SIZE_MISMATCH;
    }
    const size_t need = (size_t) height * (1 + (size_t) width * (info.bit_depth / 8));
    if (need > out_cap) { err = "PNG: output buffer too small"; return CDS_ERR_CAPACITY; }
    if (bit_depth_out) *bit_depth_out = info.bit_depth;
    z_stream z;
    memset(&z, 0, sizeof z);
    if (inflateInit(&z) != Z_OK) { err = "zlib inflateInit failed"; return CDS_ERR_OOM; }
    z.next_out = out;
    z.avail_out = (uInt) std::min<size_t>(need, 0xFFFFFFFFu);
    int zr = Z_OK;
    for (size_t k = 0; k < idat.size() && zr == Z_OK; k++) {
        z.next_in = const_cast<Bytef *>(file + idat[k].first);
        z.avail_in = (uInt) idat[k].second;
        zr = inflate(&z, Z_NO_FLUSH);
    }
    const size_t got = need - z.avail_out;
    inflateEnd(&z);
    if (zr != Z_OK && zr != Z_STREAM_END) { err = "PNG: corrupt zlib stream"; return CDS_ERR_BAD_ARG; }
    if (got != need) { err = "PNG: the image data is shorter than the image"; return CDS_ERR_BAD_ARG; }
    return CDS_OK;
}

// the IDAT payloads of a PNG, back to back: its zlib stream, for the device inflate (cds_inflate.cu)
cds_status png_collect_idat(const uint8_t *file, size_t len, int W, int H, uint8_t *dst, size_t cap, size_t base, size_t *used, InflateJob *job,
                            uint8_t *bps, std::string &err)
{
    cds_png_info info;
    std::vector<std::pair<size_t, size_t>> idat;
    cds_status s = png_parse(file, len, info, &idat, err);
    if (s != CDS_OK) return s;
    if (!info.decodable) { err = "PNG is not a non-interlaced 8- or 16-bit grayscale image"; return CDS_ERR_UNSUPPORTED; }
    if (info.width != W || info.height != H) {
        char buf[160];
        snprintf(buf, sizeof buf, "Invalid image size - PNG image size (%d, %d) must match (%d, %d)", info.width, info.height, W, H);
        err = buf;
        return CDS_ERR_SIZE_MISMATCH;
    }
    size_t total = 0;
    for (const auto &c : idat) total += c.second;
    if (total > cap || base + total > 0xFFFFFFFFull) { err = "PNG: internal staging too small"; return CDS_ERR_CAPACITY; }
    size_t o = 0;
    for (const auto &c : idat) { memcpy(dst + o, file + c.first, c.second); o += c.second; }
    *bps = (uint8_t) (info.bit_depth / 8);
    *used = total;
    // zlib header (RFC 1950): deflate, window <= 32 kB, no preset dictionary, check bits; anything else goes to the host path (src_len 0)
    const bool hdr_ok = total >= 2 && (dst[0] & 0x0F) == 8 && (dst[0] >> 4) <= 7 && (dst[1] & 0x20) == 0 && (((uint32_t) dst[0] << 8) | dst[1]) % 31 == 0;
    job->src = (uint32_t) (base + 2);
    job->src_len = hdr_ok ? (uint32_t) (total - 2) : 0u;
    return CDS_OK;
}

}  // namespace cds

// raw DEFLATE through the decoder the device runs, built for one lane (cds_inflate.h): what the CPU tests pin against zlib
extern "C" cds_status cds_debug_inflate_host(const uint8_t *in, int64_t len, uint8_t *out, int64_t capacity, int64_t *out_len, int32_t *reason)
{
    return cds::abi_guard("cds_debug_inflate_host", [&]() -> cds_status {
        if (!in || len < 0 || capacity < 0 || (capacity > 0 && !out) || !out_len) { set_tls_error("cds_debug_inflate_host: bad argument"); return CDS_ERR_BAD_ARG; }
        std::unique_ptr<cds::InflateTables> t(new cds::InflateTables);
        size_t produced = 0;
        const int st = cds::inflate_stream<1>(in, (size_t) len, out, (size_t) capacity, *t, 0, &produced);
        *out_len = (int64_t) produced;
        if (reason) *reason = st;
        if (st != cds::kInfOk) { set_tls_error("cds_debug_inflate_host: stream refused (reason " + std::to_string(st) + ")"); return st == cds::kInfOutputFull ? CDS_ERR_CAPACITY : CDS_ERR_BAD_ARG; }
        return CDS_OK;
    });
}

// ------------------------------------------------------------------------------------------------------------------ C ABI (host only)
extern "C" cds_status cds_tiff_decode_rgb_host(const uint8_t *file, int64_t len, int32_t width, int32_t height, uint8_t *out_rgb)
{
    return cds::abi_guard("cds_tiff_decode_rgb_host", [&]() -> cds_status {
        if (!file || len < 0 || width <= 0 || height <= 0 || !out_rgb) { set_tls_error("cds_tiff_decode_rgb_host: bad argument"); return CDS_ERR_BAD_ARG; }
        std::string err;
        cds_status s = tiff_decode_host(file, (size_t) len, width, height, out_rgb, err);
        if (s != CDS_OK) set_tls_error("cds_tiff_decode_rgb_host: " + err);
        return s;
    });
}

extern "C" cds_status cds_tiff_to_packbits(const uint8_t *file, int64_t len, uint8_t *out, int64_t capacity, int64_t *out_len)
{
    return cds::abi_guard("cds_tiff_to_packbits", [&]() -> cds_status {
        if (!file || len < 0 || !out || !out_len) { set_tls_error("cds_tiff_to_packbits: bad argument"); return CDS_ERR_BAD_ARG; }
        cds_tiff_info info;
        std::string err;
        cds_status s = tiff_parse(file, (size_t) len, info, nullptr, nullptr, err);
        if (s != CDS_OK) { set_tls_error("cds_tiff_to_packbits: " + err); return s; }
        // PackBits shrinks a run of 128 bytes to 2: an image that cannot fit `capacity` even then is refused before anything of the
        // size the FILE states is allocated (a corrupted size tag must not become a terabyte request)
        const uint64_t rgb_bytes = (uint64_t) info.width * (uint64_t) info.height * 3;
        if (capacity < 0 || rgb_bytes / 64 > (uint64_t) capacity) {
            set_tls_error("cds_tiff_to_packbits: the output buffer cannot hold an image of the size the file states");
            return CDS_ERR_BAD_ARG;
        }
        std::vector<uint8_t> rgb((size_t) rgb_bytes);
        s = tiff_decode_host(file, (size_t) len, info.width, info.height, rgb.data(), err);
        if (s != CDS_OK) { set_tls_error("cds_tiff_to_packbits: " + err); return s; }
        return cds_tiff_encode_rgb(rgb.data(), info.width, info.height, 8, 32773, out, capacity, out_len);
    });
}

extern "C" cds_status cds_png_probe(const uint8_t *file, int64_t len, cds_png_info *info)
{
    return cds::abi_guard("cds_png_probe", [&]() -> cds_status {
        if (!file || !info || len < 0) { set_tls_error("cds_png_probe: bad argument"); return CDS_ERR_BAD_ARG; }
        std::string err;
        cds_status s = png_parse(file, (size_t) len, *info, nullptr, err);
        if (s != CDS_OK) set_tls_error("cds_png_probe: " + err);
        return s;
    });
}

// A 16-bit grayscale PNG writer (filter type per row chosen among None / Sub / Up / Average / Paeth by the usual minimum-sum-of-
// absolute-differences heuristic, so that the reader's five reconstruction paths all get exercised); tests and the bench only.
extern "C" int64_t cds_png_encode_bound(int32_t width, int32_t height)
{
    if (width <= 0 || height <= 0) return 0;
    const uint64_t raw = (uint64_t) height * (1 + (uint64_t) width * 2);
    return (int64_t) (compressBound((uLong) raw) + 8 + 25 + 12 + 12 + 64);
}

static void png_chunk(std::vector<uint8_t> &o, const char *type, const uint8_t *data, size_t n)
{
    const uint8_t l[4] = {(uint8_t) (n >> 24), (uint8_t) (n >> 16), (uint8_t) (n >> 8), (uint8_t) n};
    o.insert(o.end(), l, l + 4);
    const size_t at = o.size();
    o.insert(o.end(), type, type + 4);
    if (n) o.insert(o.end(), data, data + n);
    const uint32_t crc = (uint32_t) crc32(0L, o.data() + at, (uInt) (4 + n));
    const uint8_t c[4] = {(uint8_t) (crc >> 24), (uint8_t) (crc >> 16), (uint8_t) (crc >> 8), (uint8_t) crc};
    o.insert(o.end(), c, c + 4);
}

extern "C" cds_status cds_png_encode_gray16(const uint16_t *pixels, int32_t width, int32_t height, int32_t filter_mode,
                                            uint8_t *out, int64_t capacity, int64_t *out_len)
{
    return cds::abi_guard("cds_png_encode_gray16", [&]() -> cds_status {
        if (!pixels || !out || !out_len || width <= 0 || height <= 0) { set_tls_error("cds_png_encode_gray16: bad argument"); return CDS_ERR_BAD_ARG; }
        const size_t rb = (size_t) width * 2;
        std::vector<uint8_t> raw((size_t) height * (1 + rb)), cur(rb), prev(rb, 0), cand(rb);
        for (int y = 0; y < height; y++) {
            for (int x = 0; x < width; x++) { cur[2 * x] = (uint8_t) (pixels[(size_t) y * width + x] >> 8); cur[2 * x + 1] = (uint8_t) pixels[(size_t) y * width + x]; }
            long best_cost = -1;
            uint8_t *dst = raw.data() + (size_t) y * (1 + rb);
            for (int f = 0; f < 5; f++) {
                if (filter_mode >= 0 && f != filter_mode) continue;
                long cost = 0;
                for (size_t i = 0; i < rb; i++) {
                    const int a = i >= 2 ? cur[i - 2] : 0, b = prev[i], c = i >= 2 ? prev[i - 2] : 0;
                    int pred = 0;
                    if (f == 1) pred = a;
                    else if (f == 2) pred = b;
                    else if (f == 3) pred = (a + b) >> 1;
                    else if (f == 4) { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
                    cand[i] = (uint8_t) (cur[i] - pred);
                    cost += std::abs((int) (int8_t) cand[i]);
                }
                if (best_cost < 0 || cost < best_cost) { best_cost = cost; dst[0] = (uint8_t) f; memcpy(dst + 1, cand.data(), rb); }
            }
            prev = cur;
        }
        uLongf zlen = compressBound((uLong) raw.size());
        std::vector<uint8_t> z(zlen);
        if (compress2(z.data(), &zlen, raw.data(), (uLong) raw.size(), 6) != Z_OK) { set_tls_error("cds_png_encode_gray16: zlib compress failed"); return CDS_ERR_OOM; }
        std::vector<uint8_t> o = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
        uint8_t hdr[13] = {(uint8_t) (width >> 24), (uint8_t) (width >> 16), (uint8_t) (width >> 8), (uint8_t) width,
                           (uint8_t) (height >> 24), (uint8_t) (height >> 16), (uint8_t) (height >> 8), (uint8_t) height, 16, 0, 0, 0, 0};
        png_chunk(o, "IHDR", hdr, 13);
        // several IDAT chunks, like the reference's files
        for (size_t at = 0; at < zlen; at += 32768) png_chunk(o, "IDAT", z.data() + at, std::min<size_t>(32768, zlen - at));
        png_chunk(o, "IEND", nullptr, 0);
        if ((int64_t) o.size() > capacity) { set_tls_error("cds_png_encode_gray16: capacity below cds_png_encode_bound"); return CDS_ERR_CAPACITY; }
        memcpy(out, o.data(), o.size());
        *out_len = (int64_t) o.size();
        return CDS_OK;
    });
}

// ------------------------------------------------------------------------------------------------------------------ zip
static uint32_t le16(const uint8_t *p) { return (uint32_t) p[0] | (uint32_t) p[1] << 8; }
static uint32_t le32(const uint8_t *p) { return le16(p) | le16(p + 2) << 16; }

extern "C" cds_status cds_zip_index(const uint8_t *archive, int64_t len, cds_zip_entry *entries, int64_t capacity, int64_t *n_entries)
{
    return cds::abi_guard("cds_zip_index", [&]() -> cds_status {
        if (!archive || len < 22 || !n_entries || capacity < 0 || (capacity > 0 && !entries)) { set_tls_error("cds_zip_index: bad argument"); return CDS_ERR_BAD_ARG; }
        *n_entries = 0;
        // end-of-central-directory record: the last 22 .. 22 + 65535 bytes
        int64_t eocd = -1;
        for (int64_t p = len - 22; p >= 0 && p >= len - 22 - 65535; p--)
            if (le32(archive + p) == 0x06054b50u) { eocd = p; break; }
        if (eocd < 0) { set_tls_error("cds_zip_index: not a zip archive (no end-of-central-directory record)"); return CDS_ERR_BAD_ARG; }
        const uint32_t total = le16(archive + eocd + 10);
        const uint64_t cd_size = le32(archive + eocd + 12), cd_off = le32(archive + eocd + 16);
        if (total == 0xFFFF || cd_off == 0xFFFFFFFFu) { set_tls_error("cds_zip_index: zip64 archives are not supported"); return CDS_ERR_UNSUPPORTED; }
        if (cd_off + cd_size > (uint64_t) len) { set_tls_error("cds_zip_index: central directory outside the archive"); return CDS_ERR_BAD_ARG; }
        uint64_t p = cd_off;
        int64_t n = 0;
        for (uint32_t e = 0; e < total; e++) {
            if (p + 46 > (uint64_t) len || le32(archive + p) != 0x02014b50u) { set_tls_error("cds_zip_index: corrupt central directory"); return CDS_ERR_BAD_ARG; }
            const uint32_t method = le16(archive + p + 10), csize = le32(archive + p + 20), usize = le32(archive + p + 24);
            const uint32_t nlen = le16(archive + p + 28), xlen = le16(archive + p + 30), clen = le16(archive + p + 32);
            const uint64_t lho = le32(archive + p + 42);
            if (p + 46 + nlen > (uint64_t) len || lho + 30 > (uint64_t) len || le32(archive + lho) != 0x04034b50u) { set_tls_error("cds_zip_index: corrupt entry"); return CDS_ERR_BAD_ARG; }
            const uint64_t data = lho + 30 + le16(archive + lho + 26) + le16(archive + lho + 28);
            if (data + csize > (uint64_t) len) { set_tls_error("cds_zip_index: entry data outside the archive"); return CDS_ERR_BAD_ARG; }
            // a stored entry IS the file (callers point into the archive with `size`): its two sizes must agree
            if (method == 0 && csize != usize) { set_tls_error("cds_zip_index: stored entry whose sizes differ"); return CDS_ERR_BAD_ARG; }
            if (n < capacity) {
                cds_zip_entry &z = entries[n];
                z.name_offset = (int64_t) (p + 46); z.name_len = (int32_t) nlen; z.method = (int32_t) method;
                z.data_offset = (int64_t) data; z.compressed_size = csize; z.size = usize; z.crc32 = le32(archive + p + 16);
                z.is_directory = (nlen > 0 && archive[p + 46 + nlen - 1] == '/') ? 1 : 0;
            }
            n++;
            p += 46 + (uint64_t) nlen + xlen + clen;
        }
        *n_entries = n;
        if (n > capacity && capacity > 0) { set_tls_error("cds_zip_index: more entries than capacity"); return CDS_ERR_CAPACITY; }
        return CDS_OK;
    });
}

extern "C" int64_t cds_zip_find(const uint8_t *archive, const cds_zip_entry *entries, int64_t n_entries, const char *name)
{
    if (!archive || !entries || !name) return -1;
    const size_t nl = strlen(name);
    for (int64_t i = 0; i < n_entries; i++)
        if ((size_t) entries[i].name_len == nl && memcmp(archive + entries[i].name_offset, name, nl) == 0) return i;
    // NeuronMIPUtils.openZipEntryStream :193-208: the first non-directory entry whose FILE NAME equals the file name of `name`
    const char *base = strrchr(name, '/');
    base = base ? base + 1 : name;
    const size_t bl = strlen(base);
    for (int64_t i = 0; i < n_entries; i++) {
        if (entries[i].is_directory) continue;
        const char *en = (const char *) archive + entries[i].name_offset;
        size_t start = 0;
        for (size_t k = 0; k < (size_t) entries[i].name_len; k++) if (en[k] == '/') start = k + 1;
        if ((size_t) entries[i].name_len - start == bl && memcmp(en + start, base, bl) == 0) return i;
    }
    return -1;
}

extern "C" cds_status cds_zip_read(const uint8_t *archive, int64_t len, const cds_zip_entry *entry, uint8_t *out, int64_t capacity)
{
    return cds::abi_guard("cds_zip_read", [&]() -> cds_status {
        if (!archive || !entry || !out || len < 0 || entry->data_offset < 0 || entry->compressed_size < 0 || entry->size < 0 ||
            entry->data_offset > len || entry->compressed_size > len - entry->data_offset || entry->size > (int64_t) 0xFFFFFFFFll ||
            (entry->method == 0 && entry->size != entry->compressed_size)) {
            set_tls_error("cds_zip_read: bad argument");
            return CDS_ERR_BAD_ARG;
        }
        if (capacity < entry->size) { set_tls_error("cds_zip_read: capacity below the entry's size"); return CDS_ERR_CAPACITY; }
        const uint8_t *src = archive + entry->data_offset;
        if (entry->method == 0) {
            memcpy(out, src, (size_t) entry->size);
        } else if (entry->method == 8) {
            z_stream z;
            memset(&z, 0, sizeof z);
            if (inflateInit2(&z, -15) != Z_OK) { set_tls_error("cds_zip_read: zlib inflateInit failed"); return CDS_ERR_OOM; }
            z.next_in = const_cast<Bytef *>(src); z.avail_in = (uInt) entry->compressed_size;
            z.next_out = out; z.avail_out = (uInt) entry->size;
            const int zr = inflate(&z, Z_FINISH);
            const bool ok = zr == Z_STREAM_END && z.avail_out == 0;
            inflateEnd(&z);
            if (!ok) { set_tls_error("cds_zip_read: corrupt deflate stream"); return CDS_ERR_BAD_ARG; }
        } else {
            set_tls_error("cds_zip_read: compression method " + std::to_string(entry->method) + " is not supported (stored and deflate are)");
            return CDS_ERR_UNSUPPORTED;
        }
        if ((uint32_t) crc32(0L, out, (uInt) entry->size) != entry->crc32) { set_tls_error("cds_zip_read: CRC mismatch"); return CDS_ERR_BAD_ARG; }
        return CDS_OK;
    });
}
