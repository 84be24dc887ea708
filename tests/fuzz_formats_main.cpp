// Corruption harness for the host-side readers of files that come from outside (TIFF tags and strips, LZW / PackBits strips,
// PNG chunks + zlib streams, zip directories).  tests/test_fuzz_cpu.py compiles this file together with csrc/cds_tiff.cpp and
// csrc/cds_formats.cpp under -fsanitize=address,undefined and runs it on the reference's own test files: every mutated file must
// end in a status (CDS_OK or an error), never in a memory error, and a call that reports CDS_OK must have filled its output.
//
//     fuzz_formats <iterations> <seed> <file> [<file> ...]          (file kind by content: TIFF, PNG or zip)
//
// Test infrastructure: nothing here ships.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../colormipsearch_b200/csrc/cds_tiff.h"

// the library keeps the last error text per thread (cds_api.cu); here it is dropped
namespace cds { void set_tls_error(const std::string &) {} }

namespace {

struct Rng {
    uint64_t s;
    uint64_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
    uint64_t below(uint64_t n) { return n ? next() % n : 0; }
};

std::vector<uint8_t> read_file(const char *path)
{
    std::vector<uint8_t> v;
    if (FILE *f = std::fopen(path, "rb")) {
        std::fseek(f, 0, SEEK_END);
        const long n = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        v.resize(n > 0 ? (size_t) n : 0);
        if (n > 0 && std::fread(v.data(), 1, v.size(), f) != v.size()) v.clear();
        std::fclose(f);
    }
    return v;
}

// A handful of byte edits; most of them where the structure lives (`hot` = a region of `hot_len` bytes), some anywhere, sometimes a
// truncation, sometimes a 16 / 32-bit field set to an extreme value.
void mutate(std::vector<uint8_t> &d, Rng &r, size_t hot, size_t hot_len)
{
    if (d.empty()) return;
    const int edits = 1 + (int) r.below(6);
    for (int e = 0; e < edits; e++) {
        size_t pos = r.below(10) < 7 ? hot + r.below(hot_len) : r.below(d.size());
        if (pos >= d.size()) pos = r.below(d.size());
        switch (r.below(4)) {
        case 0: d[pos] = (uint8_t) r.next(); break;
        case 1: d[pos] ^= (uint8_t) (1u << r.below(8)); break;
        case 2: {
            static const uint32_t extreme[] = {0u, 1u, 0x7FFFFFFFu, 0x80000000u, 0xFFFFFFFFu, 0xFFFFu, 0x10000u, 0xFFFFFFF0u};
            const uint32_t v = extreme[r.below(8)];
            for (int k = 0; k < 4 && pos + k < d.size(); k++) d[pos + k] = (uint8_t) (v >> (8 * (r.below(2) ? k : 3 - k)));
            break;
        }
        default: d[pos] = r.below(2) ? 0x00 : 0xFF; break;
        }
    }
    if (r.below(6) == 0) d.resize(r.below(d.size() + 1));
}

int failures = 0;
void expect(bool ok, const char *what, uint64_t it)
{
    if (!ok) { std::fprintf(stderr, "iteration %llu: %s\n", (unsigned long long) it, what); failures++; }
}

void run_tiff(const std::vector<uint8_t> &seed, std::vector<uint8_t> d, Rng &r, uint64_t it)
{
    uint32_t ifd = 8;
    if (seed.size() >= 8) ifd = seed[0] == 'I' ? (uint32_t) seed[4] | seed[5] << 8 | seed[6] << 16 | (uint32_t) seed[7] << 24
                                                : (uint32_t) seed[7] | seed[6] << 8 | seed[5] << 16 | (uint32_t) seed[4] << 24;
    mutate(d, r, r.below(5) == 0 ? 0 : ifd, 256);
    cds_tiff_info info;
    std::memset(&info, 0, sizeof info);
    const cds_status st = cds_tiff_probe(d.data(), (int64_t) d.size(), &info);
    int W = 0, H = 0;
    if (st == CDS_OK) {
        expect(info.width > 0 && info.height > 0, "probe: OK without a size", it);
        W = info.width; H = info.height;
    }
    // the strip table the device decoders are driven by: every strip it lists must lie inside the file
    if (st == CDS_OK && (int64_t) W * H <= (int64_t) 1 << 24) {
        std::vector<cds::TiffStrip> strips;
        std::string err;
        for (int whole = 0; whole < 2; whole++) {
            strips.clear();
            if (cds::tiff_collect_strips(d.data(), d.size(), W, H, 0, 0, strips, err, whole != 0) == CDS_OK)
                for (const cds::TiffStrip &s : strips)
                    expect(s.src_len == 0 || (uint64_t) s.src + s.src_len <= d.size(), "collect_strips: a strip outside the file", it);
        }
    }
    // host decoders: with the size the file states (bounded), and with a size it does not have
    const int dims[2][2] = {{W, H}, {64, 48}};
    for (const auto &wh : dims) {
        if (wh[0] <= 0 || wh[1] <= 0 || (int64_t) wh[0] * wh[1] > (int64_t) 1 << 22) continue;
        std::vector<uint8_t> rgb((size_t) wh[0] * wh[1] * 3);
        (void) cds_tiff_decode_rgb_host(d.data(), (int64_t) d.size(), wh[0], wh[1], rgb.data());
        const int64_t cap = cds_tiff_encode_bound(wh[0], wh[1], 8);
        std::vector<uint8_t> out((size_t) cap);
        int64_t out_len = -1;
        if (cds_tiff_to_packbits(d.data(), (int64_t) d.size(), out.data(), cap, &out_len) == CDS_OK)
            expect(out_len > 0 && out_len <= cap, "to_packbits: OK with a length outside the buffer", it);
        if (cap > 64) {   // a buffer that is too small must be refused, not overrun
            out_len = -1;
            (void) cds_tiff_to_packbits(d.data(), (int64_t) d.size(), out.data(), 64, &out_len);
        }
    }
}

void run_png(std::vector<uint8_t> d, Rng &r, uint64_t it)
{
    mutate(d, r, r.below(3) == 0 ? 0 : r.below(d.size() + 1), 64);
    cds_png_info info;
    std::memset(&info, 0, sizeof info);
    const cds_status st = cds_png_probe(d.data(), (int64_t) d.size(), &info);
    if (st == CDS_OK) expect(info.width > 0 && info.height > 0, "png probe: OK without a size", it);
    const int dims[2][2] = {{st == CDS_OK ? info.width : 0, st == CDS_OK ? info.height : 0}, {1210, 566}};
    for (const auto &wh : dims) {
        if (wh[0] <= 0 || wh[1] <= 0 || (int64_t) wh[0] * wh[1] > (int64_t) 1 << 22) continue;
        const size_t cap = (size_t) wh[1] * ((size_t) wh[0] * 2 + 1);
        std::vector<uint8_t> out(cap);
        std::string err;
        int depth = 0;
        if (cds::png_inflate(d.data(), d.size(), wh[0], wh[1], &depth, out.data(), cap, err) == CDS_OK)
            expect(depth == 8 || depth == 16, "png inflate: OK with an unknown bit depth", it);
        // and into a buffer that is too small for the stream
        (void) cds::png_inflate(d.data(), d.size(), wh[0], wh[1], &depth, out.data(), cap / 3, err);
        // the device's decoder, built for one lane, on the same stream: it must agree with zlib whenever zlib accepts the stream,
        // and end in a status otherwise
        // staged the way the library does (deflate data on a 4-byte boundary) or one byte off it: both paths of the bit reader
        const size_t lead = (it & 1) ? 2 : 3;
        std::vector<uint8_t> stage(d.size() + 8);
        size_t used = 0;
        cds::InflateJob job{0, 0};
        uint8_t bps = 0;
        if (cds::png_collect_idat(d.data(), d.size(), wh[0], wh[1], stage.data() + lead, stage.size() - lead, lead, &used, &job, &bps, err) == CDS_OK && job.src_len) {
            expect((size_t) job.src + job.src_len <= lead + used, "collect_idat: job outside the staged bytes", it);
            const size_t need = (size_t) wh[1] * (1 + (size_t) wh[0] * bps);
            std::vector<uint8_t> mine(need + 1);
            int64_t got = -1;
            int32_t why = -1;
            (void) cds_debug_inflate_host(stage.data() + job.src, job.src_len, mine.data(), (int64_t) need, &got, &why);
            expect(got >= 0 && (size_t) got <= need, "inflate: produced count outside the buffer", it);
            std::string e2;
            int d2 = 0;
            const bool zlib_ok = cds::png_inflate(d.data(), d.size(), wh[0], wh[1], &d2, out.data(), cap, e2) == CDS_OK;
            if (zlib_ok) expect((why == 0 || why == 6) && (size_t) got == need && std::memcmp(mine.data(), out.data(), need) == 0, "inflate: differs from zlib on a stream zlib accepts", it);
            (void) cds_debug_inflate_host(stage.data() + job.src, job.src_len, mine.data(), (int64_t) need / 3, &got, &why);
        }
    }
}

void run_zip(std::vector<uint8_t> d, Rng &r, uint64_t it)
{
    // the directory sits at the end of the archive
    const size_t tail = d.size() < 512 ? d.size() : 512;
    mutate(d, r, r.below(4) == 0 ? 0 : d.size() - tail, tail);
    int64_t n = -1;
    if (cds_zip_index(d.data(), (int64_t) d.size(), nullptr, 0, &n) != CDS_OK) return;
    expect(n >= 0, "zip index: OK with a negative count", it);
    if (n > 4096) return;
    std::vector<cds_zip_entry> entries((size_t) n + 1);
    int64_t n2 = -1;
    if (cds_zip_index(d.data(), (int64_t) d.size(), entries.data(), n, &n2) != CDS_OK) return;
    expect(n2 == n, "zip index: the two passes disagree", it);
    (void) cds_zip_find(d.data(), entries.data(), n, "a/b/img_1.tif");
    (void) cds_zip_find(d.data(), entries.data(), n, "");
    for (int64_t i = 0; i < n; i++) {
        const cds_zip_entry &e = entries[(size_t) i];
        const uint64_t want = (uint64_t) e.size;
        if (want > (uint64_t) 1 << 24) continue;
        std::vector<uint8_t> out((size_t) want + 1);
        (void) cds_zip_read(d.data(), (int64_t) d.size(), &e, out.data(), (int64_t) want);
        if (want > 8) (void) cds_zip_read(d.data(), (int64_t) d.size(), &e, out.data(), (int64_t) want / 2);      // too small: refused
    }
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: %s <iterations> <seed> <file>...\n", argv[0]); return 2; }
    const uint64_t iterations = std::strtoull(argv[1], nullptr, 10);
    Rng r{std::strtoull(argv[2], nullptr, 10) * 0x9E3779B97F4A7C15ull + 1};
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 3; i < argc; i++) {
        seeds.push_back(read_file(argv[i]));
        if (seeds.back().size() < 8) { std::fprintf(stderr, "cannot read %s\n", argv[i]); return 2; }
    }
    uint64_t by_kind[3] = {0, 0, 0};
    for (uint64_t it = 0; it < iterations; it++) {
        const std::vector<uint8_t> &s = seeds[it % seeds.size()];
        if (s[0] == 0x89 && s[1] == 'P') { run_png(s, r, it); by_kind[1]++; }
        else if (s[0] == 'P' && s[1] == 'K') { run_zip(s, r, it); by_kind[2]++; }
        else { run_tiff(s, s, r, it); by_kind[0]++; }
    }
    std::printf("fuzz_formats: %llu tiff, %llu png, %llu zip inputs, %d failed expectations\n", (unsigned long long) by_kind[0],
                (unsigned long long) by_kind[1], (unsigned long long) by_kind[2], failures);
    return failures ? 1 : 0;
}
