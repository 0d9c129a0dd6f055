// mock_device.cpp — TEST-ONLY stand-ins for the device entry points of the C ABI, linked into a test build of the console
// so that the host logic above the ABI (render_multiThread's call order, the export list, file names, PNG writing, the lens
// parameters of the depth-of-field exports) can be exercised without a GPU.  It renders nothing: every call is appended to
// the file named by RM_MOCK_LOG and the image calls return fixed patterns.  The host-only entry points (rm_prepare_scene,
// rm_prepared_*, rm_last_error, rm_version) still come from the real library.  Never linked into the product.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

#include "raym0nade_b200.h"

namespace {
void note(const char *fmt, ...) {
    const char *path = std::getenv("RM_MOCK_LOG");
    if (!path) return;
    FILE *f = std::fopen(path, "a");
    if (!f) return;
    va_list ap;
    va_start(ap, fmt);
    std::vfprintf(f, fmt, ap);
    va_end(ap);
    std::fputc('\n', f);
    std::fclose(f);
}
int g_npix = 0;
}  // namespace

struct RmContext { int device; };

extern "C" {
int rm_context_create(int device, void *, RmContext **out) { static RmContext c; c.device = device; *out = &c; note("context_create %d", device); return RM_OK; }
int rm_stats_reset(RmContext *) { note("stats_reset"); return RM_OK; }
int rm_stats_read(RmContext *, uint64_t out[4]) { out[0] = 1000; out[1] = out[2] = 0; out[3] = 7; return RM_OK; }
int rm_scene_upload(RmContext *, const RmSceneDesc *s) { note("scene_upload faces=%d nodes=%d lights=%d", s->n_faces, s->n_nodes, s->n_lights); return rm_scene_validate(s); }
int rm_render(RmContext *, const RmRenderArgs *a, uint64_t seed, RmHitInfo *g, RmRadiance *Dd, RmRadiance *, RmRadiance *, RmRadiance *) {
    note("render %dx%d spp=%d seed=%llu", a->width, a->height, a->spp, (unsigned long long)seed);
    g_npix = a->width * a->height;
    if (g) g[0].id = 42;
    if (Dd) Dd[0].Var = 0.5f;
    return RM_OK;
}
void rm_context_destroy(RmContext *) { note("context_destroy"); }
int rm_fxaa(RmContext *, const float *in, float *out, int32_t w, int32_t h) {
    note("fxaa %dx%d first=%.9g,%.9g,%.9g last=%.9g", w, h, in[0], in[1], in[2], in[size_t(w) * h * 3 - 1]);
    std::memcpy(out, in, size_t(w) * h * 12);
    return RM_OK;
}
int rm_upload_resolved(RmContext *, const RmRenderArgs *a, const RmHitInfo *g, const RmRadiance *Dd, const RmRadiance *Ds, const RmRadiance *, const RmRadiance *) {
    g_npix = a->width * a->height;
    note("upload_resolved %dx%d Dd0=%.9g,%.9g,%.9g Ds0=%g base0=%g", a->width, a->height, Dd[0].radiance[0], Dd[0].radiance[1], Dd[0].radiance[2],
         Ds[0].radiance[0], g[0].baseColor[0]);
    return RM_OK;
}
int rm_context_synchronize(RmContext *) { note("synchronize"); return RM_OK; }
int rm_comm_unique_id(uint8_t id[128]) { for (int i = 0; i < 128; i++) id[i] = uint8_t(i * 7 + 3); note("comm_unique_id"); return RM_OK; }
int rm_comm_init(RmContext *, const uint8_t id[128], int32_t rank, int32_t world) {
    unsigned sum = 0;
    for (int i = 0; i < 128; i++) sum += id[i];
    note("comm_init rank=%d world=%d idsum=%u", rank, world, sum);
    // ncclCommInitRank is a collective: nobody returns before everybody has arrived.  Emulated with one marker file per
    // rank in the directory RM_MOCK_BARRIER names (the host relies on it: rank 0 removes the id file right after).
    if (const char *dir = std::getenv("RM_MOCK_BARRIER")) {
        char path[512];
        std::snprintf(path, sizeof path, "%s/rank%d", dir, rank);
        if (FILE *f = std::fopen(path, "w")) std::fclose(f);
        for (int waited = 0; waited < 6000; waited++) {
            int present = 0;
            for (int r = 0; r < world; r++) {
                std::snprintf(path, sizeof path, "%s/rank%d", dir, r);
                if (FILE *f = std::fopen(path, "r")) { std::fclose(f); present++; }
            }
            if (present == world) return RM_OK;
            struct timespec ts = {0, 10 * 1000 * 1000};
            nanosleep(&ts, nullptr);
        }
        return RM_ERR_STATE;
    }
    return RM_OK;
}
int rm_trace_primary(RmContext *, const RmRenderArgs *a, int32_t *tri, float *t) { note("trace_primary %dx%d host=%d", a->width, a->height, int(tri || t)); g_npix = a->width * a->height; return RM_OK; }
int rm_gbuffer(RmContext *, const RmRenderArgs *, RmHitInfo *g) { note("gbuffer host=%d", int(g != nullptr)); return RM_OK; }
int rm_render_samples(RmContext *, const RmRenderArgs *a, int32_t begin, int32_t stride, uint64_t seed, int32_t reset) {
    note("render_samples begin=%d stride=%d spp=%d seed=%llu reset=%d", begin, stride, a->spp, (unsigned long long)seed, reset);
    return RM_OK;
}
int rm_reduce(RmContext *, int32_t root) { note("reduce root=%d", root); return RM_OK; }
int rm_resolve(RmContext *, const RmRenderArgs *, RmRadiance *Dd, RmRadiance *Ds, RmRadiance *Id, RmRadiance *Is) {
    note("resolve host=%d", int(Dd && Ds && Id && Is));
    return RM_OK;
}
int rm_download_resolved(RmContext *, RmHitInfo *g, RmRadiance *Dd, RmRadiance *, RmRadiance *, RmRadiance *) {
    note(Dd ? "download_resolved" : "download_resolved gbuffer_only");
    if (g) g[0].id = 43;
    return RM_OK;
}
int rm_spatial_clamp(RmContext *, const RmRenderArgs *) { note("spatial_clamp"); return RM_OK; }
int rm_filter(RmContext *, const RmRenderArgs *) { note("filter"); return RM_OK; }
int rm_postprocess(RmContext *, const RmRenderArgs *a, int32_t options, float *rgb) {
    note("postprocess options=%d focus=%g CoC=%g pos=%g,%g,%g exposure=%g", options, a->focus, a->CoC, a->position[0], a->position[1], a->position[2], a->exposure);
    // a pattern that depends on the options, inside and outside [0, 1): red = options / 2048, green = x / width, blue = 1
    for (int i = 0; i < g_npix; i++) {
        rgb[i * 3] = float(options) / 2048.0f;
        rgb[i * 3 + 1] = float(i % a->width) / float(a->width);
        rgb[i * 3 + 2] = 1.0f;
    }
    return RM_OK;
}
}
