// cds_tiff.cpp -- host side of the image ingest: TIFF tag parsing and a PackBits TIFF writer.
//
// What the reference does for the same files: LocalTiffDecoder.getTiffInfo() (a fork of ImageJ's TiffDecoder,
// colormipsearch-api/src/main/java/org/janelia/colormipsearch/imageprocessing/LocalTiffDecoder.java) collects width, height,
// compression, strip offsets and strip lengths of the first image; ImageArrayUtils.readImageArrayRangeWithTiffReader
// (.../imageprocessing/ImageArrayUtils.java:184-227) then walks the strips.  Only the tags are read here; the pixels are
// decoded on the device (cds_ingest.cu).
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "cds_tiff.h"
#include "cds_runtime.h"

namespace cds {

namespace {

struct Reader {
    const uint8_t *p;
    size_t len;
    bool be;
    bool ok(uint64_t off, uint64_t n) const { return off <= len && n <= len - off; }
    uint32_t u16(uint64_t off) const { return be ? (uint32_t) p[off] << 8 | p[off + 1] : (uint32_t) p[off + 1] << 8 | p[off]; }
    uint32_t u32(uint64_t off) const
    {
        return be ? (uint32_t) p[off] << 24 | (uint32_t) p[off + 1] << 16 | (uint32_t) p[off + 2] << 8 | p[off + 3]
                  : (uint32_t) p[off + 3] << 24 | (uint32_t) p[off + 2] << 16 | (uint32_t) p[off + 1] << 8 | p[off];
    }
};

int type_size(uint32_t type)
{
    switch (type) {
        case 1: case 2: case 6: case 7: return 1;      // BYTE, ASCII, SBYTE, UNDEFINED
        case 3: case 8: return 2;                      // SHORT, SSHORT
        case 4: case 9: case 11: return 4;             // LONG, SLONG, FLOAT
        case 5: case 10: case 12: return 8;            // RATIONAL, SRATIONAL, DOUBLE
        default: return 0;
    }
}

// the values of an IFD entry of type BYTE / SHORT / LONG
bool entry_values(const Reader &r, uint64_t entry, std::vector<uint64_t> *out, uint64_t *first)
{
    const uint32_t type = r.u16(entry + 2), count = r.u32(entry + 4);
    const int ts = type_size(type);
    if ((type != 1 && type != 3 && type != 4) || count == 0) return false;
    const uint64_t bytes = (uint64_t) ts * count;
    const uint64_t at = bytes <= 4 ? entry + 8 : r.u32(entry + 8);
    if (!r.ok(at, bytes)) return false;
    auto get = [&](uint64_t i) -> uint64_t {
        return type == 1 ? r.p[at + i] : type == 3 ? r.u16(at + 2 * i) : r.u32(at + 4 * i);
    };
    if (first) *first = get(0);
    if (out) {
        out->resize(count);
        for (uint64_t i = 0; i < count; i++) (*out)[i] = get(i);
    }
    return true;
}

}  // namespace

cds_status tiff_parse(const uint8_t *file, size_t len, cds_tiff_info &info, std::vector<uint64_t> *strip_off,
                      std::vector<uint64_t> *strip_len, std::string &err, std::string *why)
{
    memset(&info, 0, sizeof info);
    if (!file || len < 8) { err = "not a TIFF file: shorter than a header"; return CDS_ERR_BAD_ARG; }
    Reader r{file, len, false};
    if (file[0] == 'M' && file[1] == 'M') r.be = true;
    else if (!(file[0] == 'I' && file[1] == 'I')) { err = "not a TIFF file: bad byte-order mark"; return CDS_ERR_BAD_ARG; }
    const uint32_t magic = r.u16(2);
    if (magic == 43) { err = "BigTIFF is not supported"; return CDS_ERR_UNSUPPORTED; }
    if (magic != 42) { err = "not a TIFF file: bad magic number"; return CDS_ERR_BAD_ARG; }
    const uint64_t ifd = r.u32(4);
    if (!r.ok(ifd, 2)) { err = "TIFF: image file directory outside the file"; return CDS_ERR_BAD_ARG; }
    const uint32_t n_entries = r.u16(ifd);
    if (!r.ok(ifd + 2, (uint64_t) n_entries * 12)) { err = "TIFF: truncated image file directory"; return CDS_ERR_BAD_ARG; }

    info.big_endian = r.be ? 1 : 0;
    info.compression = 1;
    info.samples_per_pixel = 1;
    info.bits_per_sample = 1;
    info.planar_config = 1;
    bool tiled = false, have_rps = false, same_bits = true;
    std::vector<uint64_t> offs, lens, bits;
    for (uint32_t e = 0; e < n_entries; e++) {
        const uint64_t at = ifd + 2 + (uint64_t) e * 12;
        const uint32_t tag = r.u16(at);
        uint64_t v = 0;
        switch (tag) {
            case 256: if (entry_values(r, at, nullptr, &v)) info.width = (int32_t) std::min<uint64_t>(v, INT32_MAX); break;
            case 257: if (entry_values(r, at, nullptr, &v)) info.height = (int32_t) std::min<uint64_t>(v, INT32_MAX); break;
            case 258:
                if (entry_values(r, at, &bits, &v)) {
                    info.bits_per_sample = (int32_t) v;
                    for (uint64_t b : bits) same_bits = same_bits && b == v;
                }
                break;
            case 259: if (entry_values(r, at, nullptr, &v)) info.compression = (int32_t) v; break;
            case 262: if (entry_values(r, at, nullptr, &v)) info.photometric = (int32_t) v; break;
            case 273: entry_values(r, at, &offs, nullptr); break;
            case 277: if (entry_values(r, at, nullptr, &v)) info.samples_per_pixel = (int32_t) v; break;
            case 278: if (entry_values(r, at, nullptr, &v)) { info.rows_per_strip = (int32_t) std::min<uint64_t>(v, INT32_MAX); have_rps = true; } break;
            case 279: entry_values(r, at, &lens, nullptr); break;
            case 284: if (entry_values(r, at, nullptr, &v)) info.planar_config = (int32_t) v; break;
            case 322: case 323: case 324: case 325: tiled = true; break;
            default: break;
        }
    }
    if (info.width <= 0 || info.height <= 0) { err = "TIFF: no image size"; return CDS_ERR_BAD_ARG; }
    if (!have_rps || info.rows_per_strip <= 0 || info.rows_per_strip > info.height) info.rows_per_strip = info.height;
    if (offs.empty() && !tiled) { err = "TIFF: no strip offsets"; return CDS_ERR_BAD_ARG; }
    if (lens.empty() && offs.size() == 1 && info.compression == 1) {
        // some writers leave StripByteCounts out for a single stored strip
        const uint64_t want = (uint64_t) info.width * info.height * info.samples_per_pixel * ((info.bits_per_sample + 7) / 8);
        lens.push_back(offs[0] <= len ? std::min<uint64_t>(want, len - offs[0]) : 0);
    }
    // a tiled file is reported (decodable = 0) but its strip tags, if it carries any, mean nothing: they are dropped, so nothing
    // below (or in a caller) ever indexes one table with the other's length
    if (tiled) { offs.clear(); lens.clear(); }
    if (lens.size() != offs.size()) { err = "TIFF: strip offsets and byte counts differ in number"; return CDS_ERR_BAD_ARG; }
    info.n_strips = (int32_t) offs.size();
    for (size_t i = 0; i < offs.size(); i++) {
        if (!r.ok(offs[i], lens[i])) { err = "TIFF: strip outside the file"; return CDS_ERR_BAD_ARG; }
        info.data_bytes += (int64_t) lens[i];
    }
    const int64_t expect_strips = ((int64_t) info.height + info.rows_per_strip - 1) / info.rows_per_strip;
    std::string reason;
    if (tiled) reason = "tiled TIFF";
    else if (info.compression != 1 && info.compression != 32773) reason = "compression " + std::to_string(info.compression) + " (only none and PackBits)";
    else if (info.samples_per_pixel != 3 || info.bits_per_sample != 8 || !same_bits) reason = "not 8-bit RGB";
    else if (info.planar_config != 1) reason = "planar sample layout";
    else if (info.photometric != 2) reason = "photometric interpretation is not RGB";
    else if ((int64_t) info.n_strips != expect_strips) reason = "strip count does not match RowsPerStrip";
    info.decodable = reason.empty() ? 1 : 0;
    if (why) *why = reason;
    if (strip_off) strip_off->swap(offs);
    if (strip_len) strip_len->swap(lens);
    return CDS_OK;
}

cds_status tiff_collect_strips(const uint8_t *file, size_t len, int width, int height, uint64_t src_base, uint64_t dst_base,
                               std::vector<TiffStrip> &out, std::string &err, bool whole_rows)
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
    if (!info.decodable) { err = "TIFF not decodable on the device: " + why; return CDS_ERR_UNSUPPORTED; }
    const uint64_t row_bytes = (uint64_t) width * 3;
    const uint32_t flag = info.compression == 32773 ? kTiffStripPacked : 0u;
    for (int i = 0; i < info.n_strips; i++) {
        const uint64_t row0 = (uint64_t) i * info.rows_per_strip;
        const uint64_t rows = std::min<uint64_t>(info.rows_per_strip, (uint64_t) height - row0);
        // whole_rows (the fused ingest): `dst` counts image ROWS from the start of the chunk, so a chunk is not bounded by 4 GB of pixels
        const uint64_t src = src_base + offs[i], dst = whole_rows ? dst_base + row0 : dst_base + row0 * row_bytes, dlen = rows * row_bytes;
        if (src + lens[i] > 0xFFFFFFFFull || dst + (whole_rows ? rows : dlen) > 0xFFFFFFFFull || dlen >= kTiffStripPacked) {
            err = "TIFF: chunk too large for the strip table (lower stream_chunk)";
            return CDS_ERR_UNSUPPORTED;
        }
        if (flag) { out.push_back(TiffStrip{(uint32_t) src, (uint32_t) lens[i], (uint32_t) dst, (uint32_t) dlen | flag}); continue; }
        // stored bytes need no order: cut them into pieces so that one-strip files still spread over many warps
        const uint64_t have = std::min<uint64_t>(lens[i], dlen);
        const uint64_t step = whole_rows ? std::max<uint64_t>(1, kTiffStoredPiece / row_bytes) * row_bytes : kTiffStoredPiece;
        for (uint64_t o = 0; o < dlen; o += step) {
            const uint64_t piece = std::min<uint64_t>(step, dlen - o);
            const uint64_t avail = o < have ? std::min<uint64_t>(piece, have - o) : 0;
            out.push_back(TiffStrip{(uint32_t) (src + o), (uint32_t) avail, (uint32_t) (whole_rows ? dst + o / row_bytes : dst + o), (uint32_t) piece});
        }
    }
    return CDS_OK;
}

}  // namespace cds

extern "C" cds_status cds_tiff_probe(const uint8_t *file, int64_t len, cds_tiff_info *info)
{
    return cds::abi_guard("cds_tiff_probe", [&]() -> cds_status {
        if (!file || !info || len < 0) { cds::set_tls_error("cds_tiff_probe: bad argument"); return CDS_ERR_BAD_ARG; }
        std::string err;
        cds_status s = cds::tiff_parse(file, (size_t) len, *info, nullptr, nullptr, err);
        if (s != CDS_OK) cds::set_tls_error("cds_tiff_probe: " + err);
        return s;
    });
}

// ------------------------------------------------------------------------------------------------------------------ writer
namespace {

// PackBits of one row (TIFF 6.0, section 9): a control byte n in 0..127 is followed by n + 1 literal bytes, n in -127..-1 by
// one byte to repeat 1 - n times; -128 is never written.  Runs of two are kept inside literals unless they stand alone.
size_t packbits_row(const uint8_t *in, size_t n, uint8_t *out)
{
    size_t o = 0, i = 0;
    while (i < n) {
        size_t run = 1;
        while (i + run < n && run < 128 && in[i + run] == in[i]) run++;
        if (run >= 3 || (run == 2 && (i + 2 == n))) {
            out[o++] = (uint8_t) (257 - run);
            out[o++] = in[i];
            i += run;
            continue;
        }
        // literal: up to the next run of three or more
        size_t start = i, lit = 0;
        while (i < n && lit < 128) {
            size_t r2 = 1;
            while (i + r2 < n && r2 < 3 && in[i + r2] == in[i]) r2++;
            if (r2 >= 3) break;
            i++; lit++;
        }
        out[o++] = (uint8_t) (lit - 1);
        memcpy(out + o, in + start, lit);
        o += lit;
    }
    return o;
}

void put16(uint8_t *p, uint32_t v) { p[0] = (uint8_t) v; p[1] = (uint8_t) (v >> 8); }
void put32(uint8_t *p, uint32_t v) { p[0] = (uint8_t) v; p[1] = (uint8_t) (v >> 8); p[2] = (uint8_t) (v >> 16); p[3] = (uint8_t) (v >> 24); }

}  // namespace

extern "C" int64_t cds_tiff_encode_bound(int32_t width, int32_t height, int32_t rows_per_strip)
{
    if (width <= 0 || height <= 0) return 0;
    if (rows_per_strip <= 0 || rows_per_strip > height) rows_per_strip = height;
    const int64_t row = (int64_t) width * 3;
    const int64_t strips = ((int64_t) height + rows_per_strip - 1) / rows_per_strip;
    return 8 + (row + row / 128 + 2) * height + strips * 8 + 2 + 12 * 12 + 4 + 6 + 16;
}

extern "C" cds_status cds_tiff_encode_rgb(const uint8_t *rgb, int32_t width, int32_t height, int32_t rows_per_strip, int32_t compression,
                                          uint8_t *out, int64_t capacity, int64_t *out_len)
{
    return cds::abi_guard("cds_tiff_encode_rgb", [&]() -> cds_status {
        if (!rgb || !out || !out_len || width <= 0 || height <= 0) { cds::set_tls_error("cds_tiff_encode_rgb: bad argument"); return CDS_ERR_BAD_ARG; }
        if (compression != 1 && compression != 32773) { cds::set_tls_error("cds_tiff_encode_rgb: compression must be 1 or 32773"); return CDS_ERR_UNSUPPORTED; }
        if (rows_per_strip <= 0 || rows_per_strip > height) rows_per_strip = height;
        if (capacity < cds_tiff_encode_bound(width, height, rows_per_strip)) { cds::set_tls_error("cds_tiff_encode_rgb: capacity below cds_tiff_encode_bound"); return CDS_ERR_CAPACITY; }
        const size_t row = (size_t) width * 3;
        const int strips = (height + rows_per_strip - 1) / rows_per_strip;
        std::vector<uint32_t> offs(strips), lens(strips);
        size_t o = 8;
        for (int s = 0; s < strips; s++) {
            offs[s] = (uint32_t) o;
            const int y1 = std::min(height, (s + 1) * rows_per_strip);
            for (int y = s * rows_per_strip; y < y1; y++) {
                if (compression == 1) { memcpy(out + o, rgb + (size_t) y * row, row); o += row; }
                else o += packbits_row(rgb + (size_t) y * row, row, out + o);
            }
            lens[s] = (uint32_t) (o - offs[s]);
        }
        if (o & 1) out[o++] = 0;
        // out-of-line values: BitsPerSample, then the two strip tables when there is more than one strip
        const size_t bits_at = o;
        put16(out + o, 8); put16(out + o + 2, 8); put16(out + o + 4, 8); o += 6;
        size_t offs_at = 0, lens_at = 0;
        if (strips > 1) {
            offs_at = o;
            for (int s = 0; s < strips; s++, o += 4) put32(out + o, offs[s]);
            lens_at = o;
            for (int s = 0; s < strips; s++, o += 4) put32(out + o, lens[s]);
        }
        const size_t ifd = o;
        struct Entry { uint16_t tag, type; uint32_t count, value; };
        const Entry entries[] = {
            {256, 4, 1, (uint32_t) width}, {257, 4, 1, (uint32_t) height}, {258, 3, 3, (uint32_t) bits_at},
            {259, 3, 1, (uint32_t) compression}, {262, 3, 1, 2}, {273, 4, (uint32_t) strips, strips > 1 ? (uint32_t) offs_at : offs[0]},
            {277, 3, 1, 3}, {278, 4, 1, (uint32_t) rows_per_strip}, {279, 4, (uint32_t) strips, strips > 1 ? (uint32_t) lens_at : lens[0]},
            {284, 3, 1, 1},
        };
        const int n_entries = (int) (sizeof entries / sizeof entries[0]);
        put16(out + o, n_entries); o += 2;
        for (const Entry &e : entries) {
            put16(out + o, e.tag); put16(out + o + 2, e.type); put32(out + o + 4, e.count);
            if (e.type == 3 && e.count == 1) { put16(out + o + 8, e.value); put16(out + o + 10, 0); }
            else put32(out + o + 8, e.value);
            o += 12;
        }
        put32(out + o, 0); o += 4;
        out[0] = 'I'; out[1] = 'I'; put16(out + 2, 42); put32(out + 4, (uint32_t) ifd);
        *out_len = (int64_t) o;
        return CDS_OK;
    });
}
