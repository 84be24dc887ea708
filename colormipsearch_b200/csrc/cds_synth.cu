// cds_synth.cu -- device and host renderers of the synthetic MIPs of cds_synth.h, and the C ABI entry points that use
// them (cds_synth_rgb, cds_synth_gradient, cds_library_generate_synthetic).
#include <algorithm>
#include <functional>
#include <vector>

#include "cds_lut.h"
#include "cds_runtime.h"
#include "cds_synth.h"

using namespace cds;

namespace cds {

__constant__ uint8_t c_lut[256 * 3];
static bool g_lut_uploaded[64] = {false};

static cudaError_t ensure_lut(int dev)
{
    if (dev < 64 && g_lut_uploaded[dev]) return cudaSuccess;
    cudaError_t e = cudaMemcpyToSymbol(c_lut, kColorDepthLut, sizeof(kColorDepthLut));
    if (e == cudaSuccess && dev < 64) g_lut_uploaded[dev] = true;
    return e;
}

// one CTA per (row, image)
__global__ void __launch_bounds__(256) synth_render_kernel(const SynthSpec *__restrict__ specs, uint8_t *__restrict__ rgb)
{
    __shared__ int16_t s_list[CDS_SYNTH_MAX_CAPS];
    __shared__ int s_n;
    extern __shared__ uint8_t s_row[];
    const SynthSpec &sp = specs[blockIdx.y];
    const int y = blockIdx.x, W = sp.W, H = sp.H;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < sp.n; k += blockDim.x) {
        const SynthCapsule &c = sp.caps[k];
        int ymin = min(c.y0, c.y1) - c.r, ymax = max(c.y0, c.y1) + c.r;
        if (y >= ymin && y <= ymax) s_list[atomicAdd(&s_n, 1)] = (int16_t) k;
    }
    __syncthreads();
    const int n_list = s_n;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        uint8_t r, g, b;
        synth_pixel(sp, c_lut, s_list, n_list, x, y, r, g, b);
        s_row[3 * x] = r; s_row[3 * x + 1] = g; s_row[3 * x + 2] = b;
    }
    __syncthreads();
    uint8_t *dst = rgb + ((size_t) blockIdx.y * H + y) * W * 3;
    for (int k = threadIdx.x; k < 3 * W; k += blockDim.x) dst[k] = s_row[k];
}

__global__ void __launch_bounds__(256) synth_gradient_kernel(const SynthSpec *__restrict__ specs, uint16_t *__restrict__ grad)
{
    const SynthSpec &sp = specs[blockIdx.y];
    const int y = blockIdx.x, W = sp.W, H = sp.H;
    uint16_t *dst = grad + ((size_t) blockIdx.y * H + y) * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) dst[x] = synth_gradient_pixel(sp, x, y);
}

void synth_render_host(const SynthSpec &sp, uint8_t *rgb)
{
    std::vector<int16_t> list(CDS_SYNTH_MAX_CAPS);
    for (int y = 0; y < sp.H; y++) {
        int n = 0;
        for (int k = 0; k < sp.n; k++) {
            const SynthCapsule &c = sp.caps[k];
            int ymin = std::min(c.y0, c.y1) - c.r, ymax = std::max(c.y0, c.y1) + c.r;
            if (y >= ymin && y <= ymax) list[n++] = (int16_t) k;
        }
        uint8_t *row = rgb + (size_t) y * sp.W * 3;
        for (int x = 0; x < sp.W; x++) synth_pixel(sp, kColorDepthLut, list.data(), n, x, y, row[3 * x], row[3 * x + 1], row[3 * x + 2]);
    }
}

void synth_gradient_host(const SynthSpec &sp, uint16_t *grad)
{
    for (int y = 0; y < sp.H; y++)
        for (int x = 0; x < sp.W; x++) grad[(size_t) y * sp.W + x] = synth_gradient_pixel(sp, x, y);
}

// Renders `cnt` images [index0, index0 + cnt) of `kind` into a device buffer (uint8 rgb) on ds.stream.
cds_status synth_render_device(cds_ctx *ctx, DevState &ds, int kind, uint64_t seed, int64_t index0, int64_t cnt, int W, int H,
                               SynthSpec *d_specs, uint8_t *d_rgb)
{
    std::vector<SynthSpec> specs((size_t) cnt);
    for (int64_t i = 0; i < cnt; i++) synth_make_spec(kind, seed, index0 + i, W, H, specs[(size_t) i]);
    cds_status s = ctx->check(ensure_lut(ds.dev), "lut upload");
    if (s != CDS_OK) return s;
    s = ctx->check(cudaMemcpyAsync(d_specs, specs.data(), (size_t) cnt * sizeof(SynthSpec), cudaMemcpyHostToDevice, ds.stream), "spec H2D");
    if (s != CDS_OK) return s;
    dim3 grid(H, (unsigned) cnt);
    synth_render_kernel<<<grid, 256, (size_t) W * 3, ds.stream>>>(d_specs, d_rgb);
    ctx->stats.kernel_launches++;
    return ctx->check(cudaGetLastError(), "synth_render_kernel");
}

}  // namespace cds

extern "C" cds_status cds_library_generate_synthetic(cds_library *lib, uint64_t seed, int64_t first_synth_index, int64_t n, int64_t *first_index)
{
    return cds::abi_guard("cds_library_generate_synthetic", [&]() -> cds_status {
        if (!lib) { set_tls_error("cds_library_generate_synthetic: NULL library"); return CDS_ERR_BAD_ARG; }
        cds_ctx *ctx = lib->ctx;
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        if (lib->g.W > 16000 || lib->g.H > 16000) return ctx->fail(CDS_ERR_UNSUPPORTED, "synthetic images are limited to 16000 x 16000");
        const int D = lib->n_dev();
        std::vector<SynthSpec *> d_specs(D, nullptr);
        cds_status st = CDS_OK;
        for (int d = 0; d < D && st == CDS_OK; d++) {
            st = ctx->check(cudaSetDevice(ctx->devs[d].dev), "cudaSetDevice");
            if (st == CDS_OK) st = ctx->check(cudaMalloc(&d_specs[d], (size_t) kLibBlock * sizeof(SynthSpec)), "cudaMalloc(specs)");
        }
        if (st == CDS_OK) {
            st = library_append(lib, n, [&](DevState &ds, int64_t i0, int64_t cnt, uint8_t *d_rgb) -> cds_status {
                int di = 0;
                for (int d = 0; d < D; d++) if (ctx->devs[d].dev == ds.dev) di = d;
                return synth_render_device(ctx, ds, 1, seed, first_synth_index + i0, cnt, lib->g.W, lib->g.H, d_specs[di], d_rgb);
            }, first_index);
        }
        for (int d = 0; d < D; d++) if (d_specs[d]) { cudaSetDevice(ctx->devs[d].dev); cudaFree(d_specs[d]); }
        return st;
    });
}

extern "C" cds_status cds_synth_rgb(cds_ctx *ctx, int32_t kind, uint64_t seed, int64_t first_index, int64_t n,
                                    int32_t width, int32_t height, int32_t on_device, uint8_t *rgb_out)
{
    return cds::abi_guard("cds_synth_rgb", [&]() -> cds_status {
        if (n < 0 || (n > 0 && !rgb_out) || width <= 0 || height <= 0 || width > 16000 || height > 16000 || (kind != 0 && kind != 1)) {
            set_tls_error("cds_synth_rgb: bad arguments");
            return CDS_ERR_BAD_ARG;
        }
        const size_t img_bytes = (size_t) width * height * 3;
        if (!on_device) {
            std::vector<SynthSpec> spec(1);
            for (int64_t i = 0; i < n; i++) {
                synth_make_spec(kind, seed, first_index + i, width, height, spec[0]);
                synth_render_host(spec[0], rgb_out + (size_t) i * img_bytes);
            }
            return CDS_OK;
        }
        if (!ctx) { set_tls_error("cds_synth_rgb: device generation needs a context"); return CDS_ERR_BAD_ARG; }
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        DevState &d0 = ctx->devs[0];
        {
            cds_status st = ctx->check(cudaSetDevice(d0.dev), "cudaSetDevice");
            if (st != CDS_OK) return st;
            SynthSpec *d_specs = nullptr;
            st = ctx->check(cudaMalloc(&d_specs, (size_t) kLibBlock * sizeof(SynthSpec)), "cudaMalloc(specs)");
            if (st != CDS_OK) return st;
            st = ctx->ensure_staging(d0, (size_t) kLibBlock * img_bytes);
            for (int64_t i0 = 0; i0 < n && st == CDS_OK; i0 += kLibBlock) {
                int64_t cnt = std::min<int64_t>(kLibBlock, n - i0);
                st = synth_render_device(ctx, d0, kind, seed, first_index + i0, cnt, width, height, d_specs, (uint8_t *) d0.staging);
                if (st == CDS_OK) st = ctx->check(cudaMemcpyAsync(rgb_out + (size_t) i0 * img_bytes, d0.staging, (size_t) cnt * img_bytes, cudaMemcpyDeviceToHost, d0.stream), "synth D2H");
                if (st == CDS_OK) st = ctx->check(cudaStreamSynchronize(d0.stream), "synth sync");
            }
            cudaFree(d_specs);
            return st;
        }
    });
}

extern "C" cds_status cds_synth_gradient(cds_ctx *ctx, uint64_t seed, int64_t first_index, int64_t n,
                                         int32_t width, int32_t height, int32_t on_device, uint16_t *grad_out)
{
    return cds::abi_guard("cds_synth_gradient", [&]() -> cds_status {
        if (n < 0 || (n > 0 && !grad_out) || width <= 0 || height <= 0 || width > 16000 || height > 16000) {
            set_tls_error("cds_synth_gradient: bad arguments");
            return CDS_ERR_BAD_ARG;
        }
        const size_t img_px = (size_t) width * height;
        std::vector<SynthSpec> spec(1);
        if (!on_device) {
            for (int64_t i = 0; i < n; i++) {
                synth_make_spec(1, seed, first_index + i, width, height, spec[0]);
                synth_gradient_host(spec[0], grad_out + (size_t) i * img_px);
            }
            return CDS_OK;
        }
        if (!ctx) { set_tls_error("cds_synth_gradient: device generation needs a context"); return CDS_ERR_BAD_ARG; }
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        DevState &d0 = ctx->devs[0];
        cds_status st = ctx->check(cudaSetDevice(d0.dev), "cudaSetDevice");
        if (st != CDS_OK) return st;
        SynthSpec *d_spec = nullptr;
        uint16_t *d_grad = nullptr;
        st = ctx->check(cudaMalloc(&d_spec, sizeof(SynthSpec)), "cudaMalloc(spec)");
        if (st == CDS_OK) st = ctx->check(cudaMalloc(&d_grad, img_px * sizeof(uint16_t)), "cudaMalloc(grad)");
        for (int64_t i = 0; i < n && st == CDS_OK; i++) {
            synth_make_spec(1, seed, first_index + i, width, height, spec[0]);
            st = ctx->check(cudaMemcpyAsync(d_spec, spec.data(), sizeof(SynthSpec), cudaMemcpyHostToDevice, d0.stream), "spec H2D");
            if (st != CDS_OK) break;
            dim3 grid(height, 1);
            synth_gradient_kernel<<<grid, 256, 0, d0.stream>>>(d_spec, d_grad);
            ctx->stats.kernel_launches++;
            st = ctx->check(cudaGetLastError(), "synth_gradient_kernel");
            if (st == CDS_OK) st = ctx->check(cudaMemcpyAsync(grad_out + (size_t) i * img_px, d_grad, img_px * sizeof(uint16_t), cudaMemcpyDeviceToHost, d0.stream), "grad D2H");
            if (st == CDS_OK) st = ctx->check(cudaStreamSynchronize(d0.stream), "grad sync");
        }
        if (d_spec) cudaFree(d_spec);
        if (d_grad) cudaFree(d_grad);
        return st;
    });
}
