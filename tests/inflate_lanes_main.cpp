// The warp-level protocol of the device's DEFLATE decoder (csrc/cds_inflate.h) checked without a GPU: the template is built for 32
// lanes, every lane is a host thread, the warp barrier is a std::barrier -- and the whole thing runs under ThreadSanitizer.  Lanes of
// a warp may interleave freely between barriers (independent thread scheduling), which is exactly what 32 threads do; a read of the
// tables, the ring or the output that is not ordered behind its write by a barrier is a data race TSan reports, and a wrong byte
// shows in the comparison with zlib.  tests/test_inflate_cpu.py builds and runs this on streams of every kind.
//
//     inflate_lanes <rounds> <zlib stream file> ...
//
// Test infrastructure: nothing here ships.
#include <barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <zlib.h>

static std::barrier<> *g_barrier = nullptr;
static thread_local uint32_t t_rng = 1;
static void lane_barrier()
{
    // some lanes dawdle before they arrive, so that the others run ahead as far as the protocol lets them
    t_rng = t_rng * 1664525u + 1013904223u;
    if ((t_rng >> 24) < 8) std::this_thread::yield();
    g_barrier->arrive_and_wait();
}
#define CDS_INF_HOST_SYNC lane_barrier
#include "../colormipsearch_b200/csrc/cds_inflate.h"

static std::vector<uint8_t> read_file(const char *path)
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

int main(int argc, char **argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s <rounds> <zlib stream>...\n", argv[0]); return 2; }
    const int rounds = std::atoi(argv[1]);
    constexpr int kLanes = 32;
    int bad = 0;
    for (int a = 2; a < argc; a++) {
        const std::vector<uint8_t> z = read_file(argv[a]);
        if (z.size() < 6) { std::fprintf(stderr, "cannot read %s\n", argv[a]); return 2; }
        std::vector<uint8_t> want(64u << 20);
        uLongf want_len = (uLongf) want.size();
        if (uncompress(want.data(), &want_len, z.data(), (uLong) z.size()) != Z_OK) { std::fprintf(stderr, "%s: zlib refuses it\n", argv[a]); return 2; }
        for (int round = 0; round < rounds; round++) {
            // the deflate data on a 4-byte boundary (as the library stages it) in even rounds, off it in odd ones
            std::vector<uint8_t> in(z.size() + 16);
            const size_t lead = (round & 1) ? 3 : 4;
            std::memcpy(in.data() + lead, z.data() + 2, z.size() - 2);
            // one round with exactly the room the stream needs, one with less (the cut-match path)
            const size_t cap = (round & 2) ? want_len - want_len / 3 : want_len;
            std::vector<uint8_t> out(cap + 64, 0xEE);
            cds::InflateTables tables;
            std::barrier<> bar(kLanes);
            g_barrier = &bar;
            int status[kLanes];
            size_t produced[kLanes];
            std::vector<std::thread> lanes;
            for (int l = 0; l < kLanes; l++)
                lanes.emplace_back([&, l]() {
                    t_rng = 12345u + 977u * (uint32_t) l + 31u * (uint32_t) round;
                    status[l] = cds::inflate_stream<kLanes>(in.data() + lead, z.size() - 2, out.data(), cap, tables, l, &produced[l]);
                });
            for (auto &t : lanes) t.join();
            const int expect_status = cap < want_len ? cds::kInfOutputFull : cds::kInfOk;
            for (int l = 0; l < kLanes; l++)
                if (status[l] != expect_status || produced[l] != cap) { std::fprintf(stderr, "%s round %d lane %d: status %d, %zu bytes\n", argv[a], round, l, status[l], produced[l]); bad++; }
            if (std::memcmp(out.data(), want.data(), cap) != 0) { std::fprintf(stderr, "%s round %d: output differs from zlib\n", argv[a], round); bad++; }
            for (size_t i = cap; i < cap + 64; i++) if (out[i] != 0xEE) { std::fprintf(stderr, "%s round %d: wrote past the buffer\n", argv[a], round); bad++; break; }
        }
    }
    std::printf("inflate_lanes: %d stream(s) x %d round(s) on %d lanes, %d problem(s)\n", argc - 2, rounds, kLanes, bad);
    return bad ? 1 : 0;
}
