// rm_render.cu — G-buffer, sample loops, resolve and post pass (C ABI).
#include "rm_context.cuh"

using namespace rm;

void rm_render_state_free(RmContext *ctx) { (void)ctx; }

extern "C" {

int rm_gbuffer(RmContext *, const RmRenderArgs *, RmHitInfo *) { return rm_fail(RM_ERR_STATE, "rm_gbuffer: not built yet"); }
int rm_render_samples(RmContext *, const RmRenderArgs *, int32_t, int32_t, uint64_t, int32_t) { return rm_fail(RM_ERR_STATE, "rm_render_samples: not built yet"); }
int rm_accum_view(RmContext *, float **, int64_t *, float **, int64_t *) { return rm_fail(RM_ERR_STATE, "rm_accum_view: not built yet"); }
int rm_accum_after_reduce(RmContext *, int32_t, int32_t) { return rm_fail(RM_ERR_STATE, "rm_accum_after_reduce: not built yet"); }
int rm_resolve(RmContext *, const RmRenderArgs *, RmRadiance *, RmRadiance *, RmRadiance *, RmRadiance *) { return rm_fail(RM_ERR_STATE, "rm_resolve: not built yet"); }
int rm_render(RmContext *, const RmRenderArgs *, uint64_t, RmHitInfo *, RmRadiance *, RmRadiance *, RmRadiance *, RmRadiance *) { return rm_fail(RM_ERR_STATE, "rm_render: not built yet"); }
int rm_fxaa(RmContext *, const float *, float *, int32_t, int32_t) { return rm_fail(RM_ERR_STATE, "rm_fxaa: not built yet"); }
int rm_fxaa_device(RmContext *, const float *, float *, int32_t, int32_t) { return rm_fail(RM_ERR_STATE, "rm_fxaa_device: not built yet"); }
int rm_postprocess(RmContext *, const RmRenderArgs *, int32_t, float *) { return rm_fail(RM_ERR_STATE, "rm_postprocess: not built yet"); }

} // extern "C"
