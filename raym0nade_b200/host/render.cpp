// render.cpp — render_multiThread (include/render.h:43, src/render.cpp:593-676): render one frame of `model` with `args`
// and export the same set of PNGs, in the same order, under the same names.  The pixel loop the reference deals out to
// `threads` CPU workers (605-626) is one rm_render call on a B200; the image-space passes between the exports run on the
// device as well.  Errors are printed and the call returns, as everywhere in the reference.
#include <chrono>
#include <cstdlib>
#include <iostream>

#include "host.hpp"

namespace {

// One context per process and device (include/raym0nade_b200.h); RM_DEVICE picks the device, RM_SEED the random stream.
RmContext *context() {
    static RmContext *ctx = nullptr;
    if (ctx) return ctx;
    const char *dev = std::getenv("RM_DEVICE");
    if (rm_context_create(dev ? std::atoi(dev) : 0, nullptr, &ctx) != RM_OK) {
        std::cerr << "No CUDA context: " << rm_last_error() << std::endl;
        ctx = nullptr;
    }
    return ctx;
}

double msSince(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

void render_multiThread(Model &model, const RenderArgs &args) {
    const auto startTime = std::chrono::steady_clock::now();
    const int width = args.width, height = args.height;
    if (width <= 0 || height <= 0 || args.spp < 0) {
        std::cerr << "Invalid RenderArgs: width, height must be positive and spp non-negative." << std::endl;
        return;
    }
    RmContext *ctx = context();
    if (!ctx) return;                                      // no B200, no render: there is no CPU path behind this call

    std::cout << "Rendering started on CUDA device (" << rm_version() << "); the threads argument (" << args.threads
              << ") does not apply." << std::endl;

    Photo photo(width, height);
    photo.exposure = args.exposure;

    const char *seedEnv = std::getenv("RM_SEED");
    const uint64_t seed = seedEnv ? std::strtoull(seedEnv, nullptr, 10) : 0;
    rm_stats_reset(ctx);
    if (!photo.render(ctx, model, args, seed)) return;

    uint64_t stats[4] = {0, 0, 0, 0};
    rm_stats_read(ctx, stats);
    const double renderMs = msSince(startTime);
    std::cout << "Rendering completed in " << renderMs << " ms." << std::endl;
    std::cout << "Rays traced: " << stats[0] << " (" << double(stats[0]) / (renderMs * 1e3) << " Mrays/s incl. scene upload and download)" << std::endl;

    auto exportImage = [&](const std::string &tag, int shadeOptions) {
        photo.postProcessing(shadeOptions);
        photo.save((args.savePath + "(" + tag + ").png").c_str());
    };

    // same exports, same order (src/render.cpp:635-674)
    exportImage("DiffuseColor", Photo::BaseColor);
    exportImage("DiffuseColor_FXAA", Photo::BaseColor | Photo::DoFXAA);
    exportImage("shapeNormal", Photo::shapeNormal);
    exportImage("surfaceNormal", Photo::surfaceNormal);

    photo.spatialClamp();
    exportImage("Direct_Diffuse", Photo::Direct_Diffuse);
    exportImage("Direct_Specular", Photo::Direct_Specular);
    exportImage("Indirect_Diffuse", Photo::Indirect_Diffuse);
    exportImage("Indirect_Specular", Photo::Indirect_Specular);
    exportImage("Raw", Photo::Full);
    exportImage("Raw_Bloom", Photo::Full | Photo::DoBloom);
    exportImage("Raw_FXAA", Photo::Full | Photo::DoFXAA);
    exportImage("Raw_Bloom_FXAA", Photo::Full | Photo::DoBloom | Photo::DoFXAA);
    photo.filter();
    exportImage("Direct_Diffuse_Filter", Photo::Direct_Diffuse);
    exportImage("Direct_Specular_Filter", Photo::Direct_Specular);
    exportImage("Indirect_Diffuse_Filter", Photo::Indirect_Diffuse);
    exportImage("Indirect_Specular_Filter", Photo::Indirect_Specular);
    exportImage("Filter", Photo::Full);
    exportImage("Filter_Bloom", Photo::Full | Photo::DoBloom);
    exportImage("Filter_FXAA", Photo::Full | Photo::DoFXAA);
    exportImage("Filter_Bloom_FXAA", Photo::Full | Photo::DoBloom | Photo::DoFXAA);

    if (args.CoC > 1e-4f) {                                 // eps_zero, include/geometry.h:13
        photo.focus = args.focus;
        photo.CoC = args.CoC;
        photo.cameraPosition = args.position;
        exportImage("BaseColor_DepthFieldBlur", Photo::BaseColor | Photo::DoDepthFieldBlur);
        exportImage("Filter_DepthFieldBlur", Photo::Full | Photo::DoDepthFieldBlur);
        exportImage("Filter_DepthFieldBlur_Bloom", Photo::Full | Photo::DoDepthFieldBlur | Photo::DoBloom);
        exportImage("Filter_DepthFieldBlur_FXAA", Photo::Full | Photo::DoDepthFieldBlur | Photo::DoFXAA);
        exportImage("Filter_DepthFieldBlur_Bloom_FXAA", Photo::Full | Photo::DoDepthFieldBlur | Photo::DoBloom | Photo::DoFXAA);
    }

    std::cout << "Post processing finished. Total: " << msSince(startTime) << " ms." << std::endl;
}
