// cds_shape.cu -- shape (gradient area gap) scoring.  Placeholder entry points until the kernels land.
#include "cds_runtime.h"

using namespace cds;

struct cds_shape_maskset { cds_ctx *ctx; };

extern "C" cds_status cds_shape_maskset_create(cds_ctx *ctx, int32_t, int32_t, int32_t, int32_t, int32_t, const cds_rect *, int32_t,
                                               const uint8_t *, cds_shape_maskset **out)
{
    if (out) *out = nullptr;
    if (!ctx) { set_tls_error("cds_shape_maskset_create: NULL ctx"); return CDS_ERR_BAD_ARG; }
    return ctx->fail(CDS_ERR_UNSUPPORTED, "shape scoring is not built yet");
}
extern "C" void cds_shape_maskset_destroy(cds_shape_maskset *sms) { delete sms; }
extern "C" cds_status cds_shape_maskset_add_rgb(cds_shape_maskset *, const uint8_t *, int32_t, int64_t *, int64_t *)
{
    set_tls_error("shape scoring is not built yet");
    return CDS_ERR_UNSUPPORTED;
}
extern "C" int32_t cds_shape_maskset_size(const cds_shape_maskset *) { return 0; }
extern "C" cds_status cds_shape_score_pairs(cds_ctx *, const cds_shape_maskset *, const uint8_t *, const uint16_t *, const uint8_t *,
                                            const uint8_t *, int64_t, const int32_t *, const int64_t *, int64_t, int64_t *, int64_t *, uint8_t *)
{
    set_tls_error("shape scoring is not built yet");
    return CDS_ERR_UNSUPPORTED;
}
extern "C" cds_status cds_make_zgap(cds_ctx *, const uint8_t *, int64_t, int32_t, int32_t, int32_t, double, const cds_rect *, int32_t, uint8_t *)
{
    set_tls_error("shape scoring is not built yet");
    return CDS_ERR_UNSUPPORTED;
}
