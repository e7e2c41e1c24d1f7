// vr_fused.cu -- placeholder until the fused tile kernel lands (see flow.cu).
#include "common.cuh"
int k_vr_fused(mr_context *ctx, const uint8_t *, const uint8_t *, float *)
{
    return mr_fail(ctx, MR_EINVAL, "k_vr_fused", "fused VR kernel not built");
}
