// render.cpp — render_multiThread (include/render.h:43, src/render.cpp:593-676): render one frame of `model` with `args`
// and export the same set of PNGs, in the same order, under the same names.  The pixel loop the reference deals out to
// `threads` CPU workers (605-626) is one rm_render call on a B200; the image-space passes between the exports run on the
// device as well.  Errors are printed and the call returns, as everywhere in the reference.
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <iterator>

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

    // The exports, as data: {file tag, shade options}, grouped by the state of the radiance planes they show - as
    // rendered, after the firefly clamp, after the denoiser - and the depth-of-field set when the lens has a blur circle
    // (the reference writes the same files in the same order, src/render.cpp:635-674).
    struct Export { const char *tag; int options; };
    constexpr int DD = Photo::Direct_Diffuse, DS = Photo::Direct_Specular, ID = Photo::Indirect_Diffuse, IS = Photo::Indirect_Specular;
    constexpr int Full = Photo::Full, Bloom = Photo::DoBloom, Fxaa = Photo::DoFXAA, Dof = Photo::DoDepthFieldBlur;
    static const Export gbufferViews[] = {{"DiffuseColor", Photo::BaseColor}, {"DiffuseColor_FXAA", Photo::BaseColor | Fxaa},
                                          {"shapeNormal", Photo::shapeNormal}, {"surfaceNormal", Photo::surfaceNormal}};
    static const Export clamped[] = {{"Direct_Diffuse", DD}, {"Direct_Specular", DS}, {"Indirect_Diffuse", ID}, {"Indirect_Specular", IS},
                                     {"Raw", Full}, {"Raw_Bloom", Full | Bloom}, {"Raw_FXAA", Full | Fxaa}, {"Raw_Bloom_FXAA", Full | Bloom | Fxaa}};
    static const Export filtered[] = {{"Direct_Diffuse_Filter", DD}, {"Direct_Specular_Filter", DS}, {"Indirect_Diffuse_Filter", ID},
                                      {"Indirect_Specular_Filter", IS}, {"Filter", Full}, {"Filter_Bloom", Full | Bloom},
                                      {"Filter_FXAA", Full | Fxaa}, {"Filter_Bloom_FXAA", Full | Bloom | Fxaa}};
    static const Export withLens[] = {{"BaseColor_DepthFieldBlur", Photo::BaseColor | Dof}, {"Filter_DepthFieldBlur", Full | Dof},
                                      {"Filter_DepthFieldBlur_Bloom", Full | Dof | Bloom}, {"Filter_DepthFieldBlur_FXAA", Full | Dof | Fxaa},
                                      {"Filter_DepthFieldBlur_Bloom_FXAA", Full | Dof | Bloom | Fxaa}};
    auto write = [&](const Export *list, size_t count) {
        for (size_t i = 0; i < count; i++) {
            photo.postProcessing(list[i].options);
            photo.save((args.savePath + "(" + list[i].tag + ").png").c_str());
        }
    };
    write(gbufferViews, std::size(gbufferViews));
    photo.spatialClamp();
    write(clamped, std::size(clamped));
    photo.filter();
    write(filtered, std::size(filtered));
    if (args.CoC > 1e-4f) {                                 // eps_zero, include/geometry.h:13
        photo.focus = args.focus;
        photo.CoC = args.CoC;
        photo.cameraPosition = args.position;
        write(withLens, std::size(withLens));
    }

    std::cout << "Post processing finished. Total: " << msSince(startTime) << " ms." << std::endl;
}
