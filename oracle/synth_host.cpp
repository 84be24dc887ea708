// synth_host.cpp -- host-only build of the synthetic-image generator (colormipsearch_b200/csrc/cds_synth.h), for the legs of
// bench.py that must not map the GPU library: `--impl reference` times the CPU port of the reference on inputs that come from
// here.  TEST / BENCH INFRASTRUCTURE like the rest of oracle/; the generator is input data, not part of the matching algorithm,
// and it is bit-identical to the device renderer of libcdsgpu (tests/test_abi_cpu.py compares the two).
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "../colormipsearch_b200/csrc/cds_lut.h"
#include "../colormipsearch_b200/csrc/cds_synth.h"

using namespace cds;

#define CDSS_API extern "C" __attribute__((visibility("default")))

// kind 0 = EM-like mask, 1 = LM-like target; rgb_out = uint8[n][H][W][3].  Returns 0, or 1 for bad arguments.
CDSS_API int cdss_synth_rgb(int kind, uint64_t seed, int64_t first_index, int64_t n, int W, int H, uint8_t *rgb_out)
{
    if (n < 0 || (n > 0 && !rgb_out) || W <= 0 || H <= 0 || W > 16000 || H > 16000 || (kind != 0 && kind != 1)) return 1;
    std::vector<SynthSpec> spec(1);
    std::vector<int16_t> list(CDS_SYNTH_MAX_CAPS);
    for (int64_t i = 0; i < n; i++) {
        const SynthSpec &sp = spec[0];
        synth_make_spec(kind, seed, first_index + i, W, H, spec[0]);
        uint8_t *rgb = rgb_out + (size_t) i * W * H * 3;
        for (int y = 0; y < H; y++) {
            int m = 0;
            for (int k = 0; k < sp.n; k++) {
                const SynthCapsule &c = sp.caps[k];
                const int ymin = std::min(c.y0, c.y1) - c.r, ymax = std::max(c.y0, c.y1) + c.r;
                if (y >= ymin && y <= ymax) list[m++] = (int16_t) k;
            }
            uint8_t *row = rgb + (size_t) y * W * 3;
            for (int x = 0; x < W; x++) synth_pixel(sp, kColorDepthLut, list.data(), m, x, y, row[3 * x], row[3 * x + 1], row[3 * x + 2]);
        }
    }
    return 0;
}

// gray16 gradient image of synthetic target `index`: capped distance to the nearest generated neurite
CDSS_API int cdss_synth_gradient(uint64_t seed, int64_t first_index, int64_t n, int W, int H, uint16_t *grad_out)
{
    if (n < 0 || (n > 0 && !grad_out) || W <= 0 || H <= 0 || W > 16000 || H > 16000) return 1;
    std::vector<SynthSpec> spec(1);
    for (int64_t i = 0; i < n; i++) {
        synth_make_spec(1, seed, first_index + i, W, H, spec[0]);
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) grad_out[((size_t) i * H + y) * W + x] = synth_gradient_pixel(spec[0], x, y);
    }
    return 0;
}
