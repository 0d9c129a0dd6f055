// fxaa_tool.cpp — the stand-alone anti-aliasing program (the reference's second executable, fxaa.cpp): read a rendered
// PNG, take it back to linear light, run FXAA, tone-map / gamma-encode again, write a PNG.
//     raym0nade_fxaa [in.png [out.png]]        defaults: ./output/raw_2.png -> ./output/raw_FXAA.png, as in the reference
// Both image passes run on the B200 through the C ABI: rm_fxaa is Photo::FXAA (src/image.cpp:358-452); the final
// Photo::gammaCorrection (src/image.cpp:456-468) is the tail of rm_postprocess, reached by staging the filtered frame as the
// only radiance plane of an otherwise empty Photo (shade() then passes it through untouched: white base colour, no
// emission, src/image.cpp:215-246).  Decoding the 8-bit file to linear light (Photo::load + reverseGammaCorrection,
// src/image.cpp:531-612) is a 256-entry table.
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "host.hpp"

int main(int argc, char **argv) {
    const char *in_name = argc > 1 ? argv[1] : "./output/raw_2.png";
    const char *out_name = argc > 2 ? argv[2] : "./output/raw_FXAA.png";
    int w = 0, h = 0;
    std::vector<uint8_t> rgba;
    std::string why;
    if (!loadImageRGBA(in_name, w, h, rgba, why)) { std::cerr << "Could not open file for reading: " << in_name << " (" << why << ")" << std::endl; return 1; }
    std::cout << "width: " << w << " height: " << h << std::endl;

    float linear[256];                                       // value / 255 (Photo::load), then pow(., GammaFactor = 2.2)
    for (int k = 0; k < 256; k++) linear[k] = std::pow(float(k) / 255.0f, 2.2f);
    const size_t n = size_t(w) * size_t(h);
    std::vector<float> frame(n * 3), smooth(n * 3);
    for (size_t i = 0; i < n; i++)
        for (int c = 0; c < 3; c++) frame[i * 3 + c] = linear[rgba[i * 4 + c]];

    RmContext *ctx = nullptr;
    const char *dev = std::getenv("RM_DEVICE");
    if (rm_context_create(dev ? std::atoi(dev) : 0, nullptr, &ctx) != RM_OK) { std::cerr << "No CUDA context: " << rm_last_error() << std::endl; return 1; }
    auto fail = [&](const char *what) { std::cerr << what << ": " << rm_last_error() << std::endl; rm_context_destroy(ctx); return 1; };
    if (rm_fxaa(ctx, frame.data(), smooth.data(), w, h) != RM_OK) return fail("FXAA failed");

    Photo photo(w, h);                                       // zero G-buffer and planes; the frame becomes radiance_Dd
    for (size_t i = 0; i < n; i++)
        for (int c = 0; c < 3; c++) photo.radiance_Dd[i].radiance[c] = smooth[i * 3 + c];
    RmRenderArgs a{};
    a.width = w; a.height = h; a.exposure = 1.0f;
    if (rm_upload_resolved(ctx, &a, photo.Gbuffer, photo.radiance_Dd, photo.radiance_Ds, photo.radiance_Id, photo.radiance_Is) != RM_OK) return fail("Upload failed");
    if (rm_postprocess(ctx, &a, Photo::Direct_Diffuse, reinterpret_cast<float *>(photo.pixelarray)) != RM_OK) return fail("gammaCorrection failed");
    photo.save(out_name);
    rm_context_destroy(ctx);
    return 0;
}
