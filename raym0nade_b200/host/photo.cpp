// photo.cpp — Photo (include/image.h:15-65): the frame buffers of one render and the image-space passes over them.
// The buffers keep the reference's AoS layouts (HitInfo 88 B, RadianceData 16 B, vec3 pixelarray); the passes run
// on the device through the C ABI and the host copies are refreshed after each one.
#include <cstdio>
#include <cstring>
#include <iostream>

#include <zlib.h>

#include "host.hpp"

static_assert(sizeof(vec3) == 12, "pixelarray is handed to the library as packed fp32 rgb");

RmRenderArgs toC(const RenderArgs &args) {
    RmRenderArgs a{};
    for (int k = 0; k < 3; k++) {
        a.position[k] = args.position[k]; a.direction[k] = args.direction[k];
        a.up[k] = args.up[k]; a.right[k] = args.right[k];
    }
    a.accuracy = args.accuracy; a.focus = args.focus; a.CoC = args.CoC; a.exposure = args.exposure;
    a.P_Direct = args.P_Direct; a.width = args.width; a.height = args.height; a.spp = args.spp;
    return a;
}

Photo::Photo(int width_, int height_) : width(width_), height(height_) {      // src/image.cpp:9-16
    const size_t n = size_t(width) * size_t(height);
    Gbuffer = new RmHitInfo[n]();
    radiance_Dd = new RmRadiance[n]();
    radiance_Ds = new RmRadiance[n]();
    radiance_Id = new RmRadiance[n]();
    radiance_Is = new RmRadiance[n]();
    pixelarray = new vec3[n];
}

Photo::~Photo() {
    delete[] Gbuffer;
    delete[] radiance_Dd; delete[] radiance_Ds; delete[] radiance_Id; delete[] radiance_Is;
    delete[] pixelarray;
}

static bool ok(int rc, const char *what) {
    if (rc == RM_OK) return true;
    std::cerr << what << ": " << rm_last_error() << std::endl;
    return false;
}

bool Photo::render(RmContext *ctx, const Model &model, const RenderArgs &args, uint64_t seed) {
    ctx_ = nullptr;
    if (!model.desc()) { std::cerr << "Model is empty, nothing to render." << std::endl; return false; }
    if (args.width != width || args.height != height) { std::cerr << "RenderArgs do not match the Photo size." << std::endl; return false; }
    if (!ok(rm_scene_upload(ctx, model.desc()), "Scene upload failed")) return false;
    const RmRenderArgs a = toC(args);
    if (!ok(rm_render(ctx, &a, seed, Gbuffer, radiance_Dd, radiance_Ds, radiance_Id, radiance_Is), "Rendering failed")) return false;
    ctx_ = ctx;
    args_ = args;
    return true;
}

// One rank's share of a frame rendered by `world` processes, one GPU each: this rank's interleaved sample shard
// (samples s with s mod world == rank, weighted with the global 1/spp), then the library's reduction to rank 0 over
// NCCL (rm_reduce), which alone resolves and owns the frame afterwards.  The sequence is the one
// scripts/reduce_check.py and bench.py run from Python.
bool Photo::renderSharded(RmContext *ctx, const Model &model, const RenderArgs &args, uint64_t seed, int rank, int world) {
    ctx_ = nullptr;
    if (!model.desc()) { std::cerr << "Model is empty, nothing to render." << std::endl; return false; }
    if (args.width != width || args.height != height) { std::cerr << "RenderArgs do not match the Photo size." << std::endl; return false; }
    if (!ok(rm_scene_upload(ctx, model.desc()), "Scene upload failed")) return false;
    const RmRenderArgs a = toC(args);
    if (!ok(rm_trace_primary(ctx, &a, nullptr, nullptr), "Primary rays failed") || !ok(rm_gbuffer(ctx, &a, nullptr), "G-buffer failed") ||
        !ok(rm_render_samples(ctx, &a, rank, world, seed, 1), "Rendering failed") || !ok(rm_reduce(ctx, 0), "Frame reduction failed")) return false;
    if (rank != 0) return ok(rm_context_synchronize(ctx), "Synchronize failed");          // the shard is delivered; rank 0 has the frame
    if (!ok(rm_resolve(ctx, &a, radiance_Dd, radiance_Ds, radiance_Id, radiance_Is), "Resolve failed") ||
        !ok(rm_download_resolved(ctx, Gbuffer, nullptr, nullptr, nullptr, nullptr), "Download failed")) return false;
    ctx_ = ctx;
    args_ = args;
    return true;
}

void Photo::syncPlanes() {
    ok(rm_download_resolved(ctx_, Gbuffer, radiance_Dd, radiance_Ds, radiance_Id, radiance_Is), "Download failed");
}

void Photo::spatialClamp() {
    if (!ctx_) { std::cerr << "Photo::spatialClamp: no rendered frame." << std::endl; return; }
    const RmRenderArgs a = toC(args_);
    if (ok(rm_spatial_clamp(ctx_, &a), "spatialClamp failed")) syncPlanes();
}

void Photo::filter() {
    if (!ctx_) { std::cerr << "Photo::filter: no rendered frame." << std::endl; return; }
    const RmRenderArgs a = toC(args_);
    if (ok(rm_filter(ctx_, &a), "filter failed")) syncPlanes();
}

void Photo::postProcessing(int shadeOptions) {
    if (!ctx_) { std::cerr << "Photo::postProcessing: no rendered frame." << std::endl; return; }
    RenderArgs withLens = args_;                 // focus / CoC / cameraPosition are Photo members in the reference
    withLens.focus = focus; withLens.CoC = CoC; withLens.position = cameraPosition;     // (src/render.cpp:665-668)
    withLens.exposure = exposure;
    const RmRenderArgs a = toC(withLens);
    ok(rm_postprocess(ctx_, &a, shadeOptions, reinterpret_cast<float *>(pixelarray)), "postProcessing failed");
}

// ------------------------------------------------------------------------------------------------ PNG
namespace {
void put32(std::vector<uint8_t> &v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back(uint8_t(x >> s)); }
bool chunk(FILE *fp, const char type[4], const std::vector<uint8_t> &data) {
    std::vector<uint8_t> head;
    put32(head, uint32_t(data.size()));
    uLong crc = crc32(0L, reinterpret_cast<const Bytef *>(type), 4);
    if (!data.empty()) crc = crc32(crc, data.data(), uInt(data.size()));
    std::vector<uint8_t> tail;
    put32(tail, uint32_t(crc));
    return fwrite(head.data(), 1, 4, fp) == 4 && fwrite(type, 1, 4, fp) == 4 &&
           (data.empty() || fwrite(data.data(), 1, data.size(), fp) == data.size()) && fwrite(tail.data(), 1, 4, fp) == 4;
}
}  // namespace

bool writePng(const char *file_name, const uint8_t *rgb, int width, int height) {
    if (width <= 0 || height <= 0) return false;
    const size_t stride = size_t(width) * 3;
    std::vector<uint8_t> raw((stride + 1) * size_t(height));
    for (int y = 0; y < height; y++) {
        raw[(stride + 1) * y] = 0;                                           // filter type None
        std::memcpy(&raw[(stride + 1) * y + 1], rgb + stride * y, stride);
    }
    uLongf bound = compressBound(uLong(raw.size()));
    std::vector<uint8_t> z(bound);
    if (compress2(z.data(), &bound, raw.data(), uLong(raw.size()), Z_DEFAULT_COMPRESSION) != Z_OK) return false;
    z.resize(bound);
    FILE *fp = fopen(file_name, "wb");
    if (!fp) return false;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    std::vector<uint8_t> ihdr;
    put32(ihdr, uint32_t(width));
    put32(ihdr, uint32_t(height));
    ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0});                                // 8 bit, RGB, deflate, adaptive, no interlace
    bool good = fwrite(sig, 1, 8, fp) == 8 && chunk(fp, "IHDR", ihdr) && chunk(fp, "IDAT", z) && chunk(fp, "IEND", {});
    good = (fclose(fp) == 0) && good;
    return good;
}

void Photo::save(const char *file_name) {                                    // src/image.cpp:481-529
    std::vector<uint8_t> image_data(size_t(width) * size_t(height) * 3);
    for (size_t id = 0; id < size_t(width) * size_t(height); id++)
        for (int k = 0; k < 3; k++)
            image_data[id * 3 + k] = static_cast<uint8_t>(pixelarray[id][k] * 255);      // the reference's conversion, truncating
    if (!writePng(file_name, image_data.data(), width, height))
        std::cerr << "Could not open file for writing" << std::endl;
}
