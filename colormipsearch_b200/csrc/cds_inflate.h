// cds_inflate.h -- a DEFLATE (RFC 1951) decoder laid out for ONE WARP PER STREAM, so that the zlib streams of a window of PNG
// gradient images (ImageArrayUtils.java:98-121 reads them through ImageIO, i.e. java.util.zip.Inflater) are inflated on the device
// and cross PCIe compressed (~55 kB per 1210 x 566 16-bit image instead of 1.37 MB).
//
// Everything that decides WHAT the stream says -- the bit reader, the Huffman decodes, the block headers -- is uniform scalar code:
// on the device all 32 lanes execute it redundantly with identical registers (one warp instruction either way), so no lane ever
// needs another lane's value.  Three things are lane-dependent: the decode tables of a block are built by lane 0 (shared memory),
// literals are stored by lane 0, and a match (or a stored block) is copied by all lanes, byte i by lane i % LANES; a warp-level
// barrier orders the output bytes before every copy that may read them.  With LANES = 1 the same template is plain sequential
// code: the host build of it (cds_debug_inflate_host) is what the CPU tests pin against zlib, corrupted streams included.
//
// Not here: the zlib / gzip wrappers (the caller skips the 2-byte zlib header; the Adler-32 trailer is not checked -- the stream's
// length is: the caller knows how many bytes an image has) and preset dictionaries.
#ifndef CDS_INFLATE_H
#define CDS_INFLATE_H

#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
#define CDS_INF_HD __host__ __device__ __forceinline__
#else
#define CDS_INF_HD inline
#endif

namespace cds {

constexpr int kInfFastLit = 10;        // bits of the first-level literal/length table
constexpr int kInfFastDist = 8;        // ... of the distance table
constexpr uint32_t kInfRing = 8192;    // bytes of recent output kept next to the tables (a power of two)
constexpr uint32_t kInfNear = kInfRing - 258;      // matches no further back than this are served from there

enum InflateStatus : int {
    kInfOk = 0,
    kInfInputShort = 1,        // the stream needs bits past the end of the input
    kInfBadBlock = 2,          // reserved block type / stored length check
    kInfBadLengths = 3,        // code lengths that form no prefix code
    kInfBadSymbol = 4,         // a bit pattern that is no code, or a length / distance symbol outside the alphabet
    kInfBadDistance = 5,       // a match that starts before the output does
    kInfOutputFull = 6,        // the stream holds more than `out_cap` bytes (the first out_cap of them are written)
};

// Decode state of one stream (one warp): 3 584 bytes of tables and the last 8 kB of output, shared memory on the device.
// The output itself goes to global memory and is not read back for a match within kInfNear bytes -- a store does not allocate in
// L1, so a match that read its source from global memory would wait for an L2 round trip, one after the other (measured: 85 ms per
// 1.4 MB image that way).  PNG rows are a few kB long, so matches against the previous rows stay inside the ring.
struct InflateTables {
    uint8_t ring[kInfRing];                   // output byte p lives at ring[p % kInfRing]
    uint16_t lit_fast[1 << kInfFastLit];      // (symbol << 4) | code length for codes of <= kInfFastLit bits, else 0
    uint16_t dist_fast[1 << kInfFastDist];
    uint16_t lit_sym[288];                    // symbols in canonical order (by length, then value)
    uint16_t dist_sym[32];
    uint16_t lit_count[16], dist_count[16];   // codes per length
    uint8_t lens[320];                        // code lengths of the block being set up (288 + 32)
};

struct InflateBits {
    const uint8_t *in;
    size_t len, pos;           // pos: next byte to load (may run past len: zeros are fed, and the overrun is noticed)
    uint64_t buf;
    int cnt;                   // valid bits in buf
};

// at least 33 valid bits afterwards
CDS_INF_HD void inf_refill(InflateBits &b)
{
    if (b.cnt <= 32) {
        uint32_t w = 0;
        const uint8_t *p = b.in + b.pos;
        if (b.pos + 4 <= b.len && (reinterpret_cast<uintptr_t>(p) & 3u) == 0) {
            // the usual case: callers place the data on a 4-byte boundary, and the position then moves in steps of four
            w = *reinterpret_cast<const uint32_t *>(p);            // little-endian hosts and devices only
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (b.pos + k < b.len) w |= (uint32_t) p[k] << (8 * k);
        }
        b.buf |= (uint64_t) w << b.cnt;
        b.cnt += 32;
        b.pos += 4;
    }
}
CDS_INF_HD uint32_t inf_take(InflateBits &b, int n)       // n <= 16, after a refill
{
    const uint32_t v = (uint32_t) b.buf & ((1u << n) - 1u);
    b.buf >>= n;
    b.cnt -= n;
    return v;
}
CDS_INF_HD bool inf_overrun(const InflateBits &b) { return b.pos * 8 - (size_t) b.cnt > b.len * 8; }

// Code lengths -> canonical tables.  Returns 0 for a complete code, > 0 for an incomplete one (the caller decides), < 0 for an
// over-subscribed one.  `fast` may be null (the 19-symbol code-length code is decoded bit by bit).
CDS_INF_HD int inf_build(const uint8_t *lens, int n, uint16_t *count, uint16_t *sym, uint16_t *fast, int fast_bits)
{
    for (int l = 0; l < 16; l++) count[l] = 0;
    for (int s = 0; s < n; s++) count[lens[s]]++;
    if (fast) for (int i = 0; i < (1 << fast_bits); i++) fast[i] = 0;
    if (count[0] == n) return 0;                                  // no codes at all: every decode fails, which is right
    int left = 1;
    for (int l = 1; l < 16; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return -1;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t) (offs[l] + count[l]);
    for (int s = 0; s < n; s++)
        if (lens[s]) sym[offs[lens[s]]++] = (uint16_t) s;
    if (fast) {
        // a code of l bits arrives least-significant bit first: every table index whose low l bits are the reversed code
        uint32_t code = 0;
        int idx = 0;
        for (int l = 1; l <= fast_bits; l++) {
            for (int k = 0; k < count[l]; k++, code++) {
                uint32_t rev = 0;
                for (int j = 0; j < l; j++) rev |= ((code >> j) & 1u) << (l - 1 - j);
                const uint16_t e = (uint16_t) ((sym[idx++] << 4) | l);
                for (uint32_t j = rev; j < (1u << fast_bits); j += 1u << l) fast[j] = e;
            }
            code <<= 1;
        }
    }
    return left;
}

// One symbol by the canonical walk (codes longer than the fast table, and the code-length code): -1 when the bits are no code.
CDS_INF_HD int inf_decode_slow(uint64_t bits, const uint16_t *count, const uint16_t *sym, int &len_out)
{
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; l++) {
        code |= (int) (bits & 1u);
        bits >>= 1;
        const int c = count[l];
        if (code - c < first) { len_out = l; return sym[index + (code - first)]; }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

CDS_INF_HD int inf_decode(InflateBits &b, const uint16_t *fast, int fast_bits, const uint16_t *count, const uint16_t *sym)
{
    const uint32_t e = fast[(uint32_t) b.buf & ((1u << fast_bits) - 1u)];
    int l = (int) (e & 15u), s = (int) (e >> 4);
    if (l == 0) {
        s = inf_decode_slow(b.buf, count, sym, l);
        if (s < 0) return -1;
    }
    b.buf >>= l;
    b.cnt -= l;
    return s;
}

template <int LANES>
CDS_INF_HD void inf_sync()
{
#ifdef __CUDA_ARCH__
    if (LANES > 1) __syncwarp();
#elif defined(CDS_INF_HOST_SYNC)
    if (LANES > 1) CDS_INF_HOST_SYNC();      // test builds: LANES host threads meet at a barrier (tests/inflate_lanes_main.cpp, under ThreadSanitizer)
#endif
}

// Inflates in[0, in_len) into out[0, out_cap).  *produced = bytes written (valid for lane 0 .. all lanes: uniform).
// `lane` in [0, LANES); on the device LANES = 32 and the whole warp must call this together.
template <int LANES>
CDS_INF_HD int inflate_stream(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap, InflateTables &t, int lane, size_t *produced)
{
    InflateBits b{in, in_len, 0, 0, 0};
    size_t op = 0;
    int status = kInfOk;
    bool last = false;
    while (!last && status == kInfOk) {
        inf_refill(b);
        if (inf_overrun(b)) { status = kInfInputShort; break; }
        last = inf_take(b, 1) != 0;
        const uint32_t type = inf_take(b, 2);
        if (type == 3) { status = kInfBadBlock; break; }
        if (type == 0) {
            // stored: back to a byte boundary, LEN, ~LEN, the bytes
            inf_take(b, b.cnt & 7);
            inf_refill(b);
            const uint32_t n = inf_take(b, 16), nn = inf_take(b, 16);
            if ((n ^ nn) != 0xFFFFu) { status = kInfBadBlock; break; }
            const size_t at = b.pos - (size_t) (b.cnt >> 3);      // the buffer holds whole bytes now
            if (at > in_len || n > in_len - at) { status = kInfInputShort; break; }
            const uint32_t fit = n > out_cap - op ? (uint32_t) (out_cap - op) : n;      // what does not fit is dropped, and reported
            for (uint32_t i = (uint32_t) lane; i < fit; i += LANES) {
                const uint8_t v = in[at + i];
                out[op + i] = v;
                t.ring[(op + i) & (kInfRing - 1)] = v;
            }
            op += fit;
            if (fit < n) { status = kInfOutputFull; break; }
            b.pos = at + n;
            b.buf = 0;
            b.cnt = 0;
            continue;
        }
        // ---- the block's two codes (built by lane 0 once no lane decodes with the previous block's tables any more)
        inf_sync<LANES>();
        int bad = 0;
        if (type == 1) {
            if (lane == 0) {
                for (int s = 0; s < 288; s++) t.lens[s] = (uint8_t) (s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
                for (int s = 0; s < 32; s++) t.lens[288 + s] = (uint8_t) (s < 30 ? 5 : 0);
                inf_build(t.lens, 288, t.lit_count, t.lit_sym, t.lit_fast, kInfFastLit);
                inf_build(t.lens + 288, 32, t.dist_count, t.dist_sym, t.dist_fast, kInfFastDist);
            }
        } else {
            inf_refill(b);
            const int nlen = (int) inf_take(b, 5) + 257, ndist = (int) inf_take(b, 5) + 1, ncode = (int) inf_take(b, 4) + 4;
            if (nlen > 286 || ndist > 30) { status = kInfBadLengths; break; }
            // the code-length code: 3 bits per length, in the order 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
            if (lane == 0) for (int i = 0; i < 19; i++) t.lens[i] = 0;
            for (int i = 0; i < ncode; i++) {
                inf_refill(b);
                const uint32_t v = inf_take(b, 3);
                const int k = i - 4;
                const int where = i < 3 ? 16 + i : i == 3 ? 0 : (k & 1) ? 7 - (k >> 1) : 8 + (k >> 1);
                if (lane == 0) t.lens[where] = (uint8_t) v;
            }
            if (lane == 0 && inf_build(t.lens, 19, t.dist_count, t.dist_sym, nullptr, 0) != 0) t.dist_count[0] = 0xFFFF;      // it must be complete
            inf_sync<LANES>();
            if (t.dist_count[0] == 0xFFFF) { status = kInfBadLengths; break; }
            // the nlen + ndist code lengths, run-length coded: every lane decodes, lane 0 stores them (t.lens, from index 0)
            int i = 0;
            while (i < nlen + ndist) {
                inf_refill(b);
                int l = 0;
                const int s = inf_decode_slow(b.buf, t.dist_count, t.dist_sym, l);
                if (s < 0) { bad = kInfBadSymbol; break; }
                b.buf >>= l;
                b.cnt -= l;
                if (s < 16) {
                    if (lane == 0) t.lens[i] = (uint8_t) s;
                    i++;
                    continue;
                }
                int rep;
                if (s == 16) rep = 3 + (int) inf_take(b, 2);
                else if (s == 17) rep = 3 + (int) inf_take(b, 3);
                else rep = 11 + (int) inf_take(b, 7);
                if ((s == 16 && i == 0) || i + rep > nlen + ndist) { bad = kInfBadLengths; break; }
                if (lane == 0) {
                    const uint8_t v = s == 16 ? t.lens[i - 1] : (uint8_t) 0;
                    for (int k = 0; k < rep; k++) t.lens[i + k] = v;
                }
                i += rep;
            }
            if (bad) { status = bad; break; }
            if (inf_overrun(b)) { status = kInfInputShort; break; }
            inf_sync<LANES>();                                     // every lane is done with the code-length code (it sits in the distance tables)
            if (lane == 0) {
                // distance lengths to their place (288 ..., top down: the ranges may overlap), the rest zero, then the tables
                for (int k = ndist - 1; k >= 0; k--) t.lens[288 + k] = t.lens[nlen + k];
                for (int k = nlen; k < 288; k++) t.lens[k] = 0;
                for (int k = ndist; k < 32; k++) t.lens[288 + k] = 0;
                int ok = t.lens[256] != 0;                                                                  // a block needs its end code
                // an incomplete code is legal only when it is a single code of one bit (zlib, puff)
                const int rl = inf_build(t.lens, 288, t.lit_count, t.lit_sym, t.lit_fast, kInfFastLit);
                if (rl < 0 || (rl > 0 && 288 != t.lit_count[0] + t.lit_count[1])) ok = 0;
                const int rd = inf_build(t.lens + 288, 32, t.dist_count, t.dist_sym, t.dist_fast, kInfFastDist);
                if (rd < 0 || (rd > 0 && 32 != t.dist_count[0] + t.dist_count[1])) ok = 0;
                if (!ok) t.lit_count[0] = 0xFFFF;
            }
        }
        inf_sync<LANES>();
        if (t.lit_count[0] == 0xFFFF) { status = kInfBadLengths; break; }
        // ---- symbols
        for (;;) {
            inf_refill(b);
            int s = inf_decode(b, t.lit_fast, kInfFastLit, t.lit_count, t.lit_sym);
            if (s < 0) { status = kInfBadSymbol; break; }
            if (s < 256) {
                if (op >= out_cap) { status = kInfOutputFull; break; }
                if (lane == 0) {
                    out[op] = (uint8_t) s;
                    t.ring[op & (kInfRing - 1)] = (uint8_t) s;
                }
                op++;
                continue;
            }
            if (s == 256) break;
            s -= 257;
            if (s >= 29) { status = kInfBadSymbol; break; }
            uint32_t mlen;
            if (s < 8) mlen = (uint32_t) s + 3u;
            else if (s == 28) mlen = 258u;
            else {
                const int eb = (s - 4) >> 2;
                mlen = 3u + ((4u + ((uint32_t) s & 3u)) << eb) + inf_take(b, eb);
            }
            inf_refill(b);
            const int d = inf_decode(b, t.dist_fast, kInfFastDist, t.dist_count, t.dist_sym);
            if (d < 0 || d >= 30) { status = kInfBadSymbol; break; }
            uint32_t dist;
            if (d < 4) dist = (uint32_t) d + 1u;
            else {
                const int eb = (d >> 1) - 1;
                dist = 1u + ((2u + ((uint32_t) d & 1u)) << eb) + inf_take(b, eb);
            }
            if (dist > op) { status = kInfBadDistance; break; }
            const bool cut = mlen > out_cap - op;                  // the part of the match that fits is still written
            if (cut) mlen = (uint32_t) (out_cap - op);
            inf_sync<LANES>();                                     // the bytes this match reads are written
            uint8_t *dst = out + op;
            // byte i of the match repeats source byte i mod dist (a match may overlap its own output; i mod dist = i when it does not)
            const bool pow2 = (dist & (dist - 1u)) == 0u;
            if (dist <= kInfNear) {
                // the source is among the last 8 kB: read from the ring; reads are below `op`, writes at or above it, and
                // dist + mlen <= kInfRing keeps the two apart in the ring as well
                const uint32_t base = (uint32_t) (op - dist), top = (uint32_t) op;
                if (dist >= mlen || pow2) {
                    const uint32_t jm = dist >= mlen ? 0xFFFFFFFFu : dist - 1u;      // i mod dist without a division
                    for (uint32_t i = (uint32_t) lane; i < mlen; i += LANES) {
                        const uint8_t v = t.ring[(base + (i & jm)) & (kInfRing - 1)];
                        t.ring[(top + i) & (kInfRing - 1)] = v;
                        dst[i] = v;
                    }
                } else {
                    for (uint32_t i = (uint32_t) lane; i < mlen; i += LANES) {
                        const uint8_t v = t.ring[(base + i % dist) & (kInfRing - 1)];
                        t.ring[(top + i) & (kInfRing - 1)] = v;
                        dst[i] = v;
                    }
                }
                inf_sync<LANES>();                                 // literals that follow (lane 0) reuse ring slots a slower lane may still be reading
            } else {
                const uint8_t *src = out + op - dist;              // dist > kInfNear >= mlen: no overlap
                for (uint32_t i = (uint32_t) lane; i < mlen; i += LANES) {
                    const uint8_t v = src[i];
                    t.ring[((uint32_t) op + i) & (kInfRing - 1)] = v;
                    dst[i] = v;
                }
            }
            op += mlen;
            if (cut) { status = kInfOutputFull; break; }
        }
        if (status == kInfOk && inf_overrun(b)) status = kInfInputShort;
    }
    // bits taken from beyond the input (zeros) can look like anything, a full output buffer included: that outranks every other verdict
    if (inf_overrun(b)) status = kInfInputShort;
    inf_sync<LANES>();
    *produced = op;
    return status;
}

}  // namespace cds
#endif
