// render.cpp — render_multiThread (include/render.h:43, src/render.cpp:593-676): render one frame of `model` with `args`
// and export the same set of PNGs, in the same order, under the same names.  The pixel loop the reference deals out to
// `threads` CPU workers (605-626) is one rm_render call on a B200 - or, in a launch of several processes with one GPU
// each, this rank's sample shard followed by the library's reduction to rank 0; the image-space passes between the
// exports run on the device as well.  Errors are printed and the call returns, as everywhere in the reference.
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <cstdio>
#include <iterator>
#include <string>
#include <thread>

#include "host.hpp"

namespace {

int envInt(const char *first, const char *second, int fallback) {
    for (const char *name : {first, second}) {
        const char *v = name ? std::getenv(name) : nullptr;
        if (v && *v) return std::atoi(v);
    }
    return fallback;
}

// The process's place in a multi-GPU launch: one process per GPU, started by any launcher that sets RM_RANK / RM_WORLD
// (or torchrun's RANK / WORLD_SIZE / LOCAL_RANK, e.g. `python -m torch.distributed.run --no-python ... ./raym0nade`).
struct Launch { int rank, world, device; };
Launch launch() {
    Launch L{envInt("RM_RANK", "RANK", 0), envInt("RM_WORLD", "WORLD_SIZE", 1), envInt("RM_DEVICE", "LOCAL_RANK", 0)};
    if (L.world < 1 || L.rank < 0 || L.rank >= L.world) { std::cerr << "Ignoring rank " << L.rank << " of " << L.world << "." << std::endl; L.rank = 0; L.world = 1; }
    return L;
}

// One context per process and device (include/raym0nade_b200.h); RM_SEED picks the random stream.
RmContext *context(int device) {
    static RmContext *ctx = nullptr;
    if (ctx) return ctx;
    if (rm_context_create(device, nullptr, &ctx) != RM_OK) {
        std::cerr << "No CUDA context: " << rm_last_error() << std::endl;
        ctx = nullptr;
    }
    return ctx;
}

// The NCCL unique id travels from rank 0 to the others through a file (RM_COMM_FILE; a launcher should name a fresh
// path per launch): written under a temporary name and renamed, polled by the other ranks, removed by rank 0 once
// rm_comm_init - a collective - has returned, i.e. once everybody has read it.
bool joinRanks(RmContext *ctx, const Launch &L) {
    static bool joined = false;
    if (joined) return true;
    const char *named = std::getenv("RM_COMM_FILE");
    const char *port = std::getenv("MASTER_PORT");
    const std::string path = named ? named : std::string("/tmp/raym0nade_comm_") + (port ? port : "0") + ".id";
    uint8_t id[128];
    if (L.rank == 0) {
        if (rm_comm_unique_id(id) != RM_OK) { std::cerr << "No communicator: " << rm_last_error() << std::endl; return false; }
        const std::string tmp = path + ".tmp";
        FILE *f = std::fopen(tmp.c_str(), "wb");
        if (!f || std::fwrite(id, 1, sizeof id, f) != sizeof id || std::fclose(f) != 0 || std::rename(tmp.c_str(), path.c_str()) != 0) {
            std::cerr << "Could not write " << path << std::endl;
            return false;
        }
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        size_t got = 0;
        while (got != sizeof id) {
            if (FILE *f = std::fopen(path.c_str(), "rb")) { got = std::fread(id, 1, sizeof id, f); std::fclose(f); }
            if (got == sizeof id) break;
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) { std::cerr << "Rank 0 never wrote " << path << std::endl; return false; }
            std::this_thread::sleep_for(std::chrono::milliseconds(20));
        }
    }
    const int rc = rm_comm_init(ctx, id, L.rank, L.world);
    if (L.rank == 0) std::remove(path.c_str());
    if (rc != RM_OK) { std::cerr << "No communicator: " << rm_last_error() << std::endl; return false; }
    joined = true;
    return true;
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
    const Launch L = launch();
    RmContext *ctx = context(L.device);
    if (!ctx) return;                                      // no B200, no render: there is no CPU path behind this call
    if (L.world > 1 && !joinRanks(ctx, L)) return;

    std::cout << "Rendering started on CUDA device " << L.device << " (" << rm_version() << "), rank " << L.rank << " of " << L.world
              << "; the threads argument (" << args.threads << ") does not apply." << std::endl;

    Photo photo(width, height);
    photo.exposure = args.exposure;

    const char *seedEnv = std::getenv("RM_SEED");
    const uint64_t seed = seedEnv ? std::strtoull(seedEnv, nullptr, 10) : 0;
    rm_stats_reset(ctx);
    if (L.world == 1 ? !photo.render(ctx, model, args, seed) : !photo.renderSharded(ctx, model, args, seed, L.rank, L.world)) return;
    if (L.rank != 0) {                                     // the frame now lives on rank 0, which writes the exports
        std::cout << "Rank " << L.rank << ": sample shard rendered and reduced in " << msSince(startTime) << " ms." << std::endl;
        return;
    }

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
