// bridge_harness.cpp — TEST INFRASTRUCTURE.  Exposes the reference-tree binding
// (raym0nade_b200/host/reference_tree/gpu_bridge.cpp) over a reference-built Model to the tests: the descriptor it fills is
// compared with the one the product's rm_prepare_scene derives from the same raw scene, and - on a GPU box - the
// reference's own Photo is rendered into through render_multiThread_b200 and shaded by the reference's own code.
// Built by `make -C oracle bridge` into _ref/libraym_bridge.so = the unmodified reference objects + ref_harness.cpp + the
// binding, linked against the product library exactly as a maintainer's Demo target would be.
#include <cstring>

#include "gpu_bridge.h"

extern "C" {

const void *ref_scene_model(void *scene);                     // ref_harness.cpp

void *ref_bridge_create(void *scene) {
    auto *b = new B200Bridge();
    b->fill(*static_cast<const Model *>(ref_scene_model(scene)));
    return b;
}
const RmSceneDesc *ref_bridge_desc(void *b) { return &static_cast<B200Bridge *>(b)->desc; }
void ref_bridge_destroy(void *b) { delete static_cast<B200Bridge *>(b); }

// render_multiThread_b200 into a reference Photo, then the reference's own Photo::postProcessing (src/image.cpp:470-479)
// on the CPU: rgb_out [h][w][3] is what Photo::save would quantise.  Returns 0, or -1 when the render failed (no GPU).
int ref_bridge_render(void *scene, const RmRenderArgs *a, int shade_options, float *rgb_out, RmRadiance *Dd_out) {
    Model &m = *const_cast<Model *>(static_cast<const Model *>(ref_scene_model(scene)));
    RenderArgs args;
    for (int k = 0; k < 3; k++) { args.position[k] = a->position[k]; args.direction[k] = a->direction[k]; args.up[k] = a->up[k]; args.right[k] = a->right[k]; }
    args.accuracy = a->accuracy; args.focus = a->focus; args.CoC = a->CoC; args.exposure = a->exposure; args.P_Direct = a->P_Direct;
    args.width = a->width; args.height = a->height; args.spp = a->spp; args.threads = 1;
    Photo photo(a->width, a->height);
    photo.Gbuffer[0].id = -12345;                             // render_multiThread_b200 overwrites it when it runs
    render_multiThread_b200(m, args, photo);
    if (photo.Gbuffer[0].id == -12345) return -1;
    photo.postProcessing(shade_options);
    const size_t n = size_t(a->width) * a->height;
    std::memcpy(rgb_out, photo.pixelarray, n * sizeof(vec3));
    if (Dd_out) std::memcpy(Dd_out, photo.radiance_Dd, n * sizeof(RadianceData));
    return 0;
}

}
