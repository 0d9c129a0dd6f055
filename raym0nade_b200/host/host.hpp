// host.hpp — the C++ host side above the C ABI (include/raym0nade_b200.h).
//
// Mirrors the reference's own host interface for the hot path, same names, same argument meaning, same error
// behaviour (print and return, never throw across a call):
//   RenderArgs            include/render.h:8-15
//   Model                 include/model.h:25-43   (constructor: src/model.cpp:172-215)
//   Photo                 include/image.h:15-65
//   render_multiThread    include/render.h:43     (src/render.cpp:593-676)
//   MyConsole             include/myconsole.h
// The reference loads scenes through assimp and decodes images through an embedded Python interpreter; neither is part
// of the hot path (and neither is installed here), so Model reads the raw post-import arrays - what processMesh /
// processMaterial hand on - from a `.rmscene` container (raym0nade_b200/scenes.py writes one) or from a Wavefront
// .obj/.mtl pair with PNG / DDS texture maps (image_io.cpp), and the sky from `.hdr` (Radiance RGBE) or `.pfm`.  Everything after that point - BVH build, mip
// chains, light objects, sky CDF (rm_prepare_scene), the render and every image-space pass - is the library's.
// There is no CPU renderer behind this interface: without a B200 the render prints the library's error and returns.
#ifndef RAYM0NADE_B200_HOST_HPP
#define RAYM0NADE_B200_HOST_HPP

#include <cstdint>
#include <iosfwd>
#include <map>
#include <string>
#include <vector>

#include "raym0nade_b200.h"

struct vec3 {
    float x = 0, y = 0, z = 0;
    vec3() = default;
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    float &operator[](int k) { return k == 0 ? x : (k == 1 ? y : z); }
    float operator[](int k) const { return k == 0 ? x : (k == 1 ? y : z); }
};
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }

// include/render.h:8-15, field for field.  `threads` is kept for the console surface; the device schedules itself.
struct RenderArgs {
    vec3 position, direction, up, right;
    float accuracy = 0, focus = 0, CoC = 0, exposure = 1, P_Direct = 0.7f;
    int width = 0, height = 0, spp = 0, threads = 1;
    std::string savePath;
};
RmRenderArgs toC(const RenderArgs &args);

// The loaded scene.  Like the reference's Model it must not be copied or moved after load (the prepared scene holds
// pointers into it), and a failed load leaves an empty model behind after printing the reason (src/model.cpp:185-188).
class Model {
public:
    std::string model_path, skyMap_path;
    // raw post-import arrays (RmRawScene, include/rm_types.h)
    std::vector<float> positions, uvs, normals;
    std::vector<RmRawMesh> meshes;
    std::vector<RmRawMaterial> materials;
    std::vector<std::string> materialNames;
    std::vector<std::vector<uint8_t>> texturePixels;
    std::vector<RmRawTexture> textures;
    std::vector<float> sky;
    int skyWidth = 0, skyHeight = 0;

    Model();
    Model(const std::string &model_folder, const std::string &model_name, const std::string &skyMap_name);
    ~Model();
    Model(const Model &) = delete;
    Model &operator=(const Model &) = delete;

    size_t faceCount() const { return positions.size() / 9; }
    bool empty() const { return prepared_ == nullptr; }
    // BVH build, mip chains, light objects, sky CDF: what Model::Model derives after the import (rm_prepare_scene)
    bool prepare();
    const RmSceneDesc *desc() const;

private:
    RmPrepared *prepared_ = nullptr;
    bool loadRmScene(const std::string &path);
    bool loadObj(const std::string &folder, const std::string &path);
    bool loadSky(const std::string &path);
};

class Photo {
public:
    enum ShadeOption {            // include/image.h:17-35
        BaseColor = 1, Emission = 2, DirectLight = 4, IndirectLight = 8, Diffuse = 16, Specular = 32,
        shapeNormal = 64, surfaceNormal = 128,
        Direct_Diffuse = DirectLight | Diffuse, Direct_Specular = DirectLight | Specular,
        Indirect_Diffuse = IndirectLight | Diffuse, Indirect_Specular = IndirectLight | Specular,
        Full = DirectLight | IndirectLight | Diffuse | Specular | BaseColor | Emission,
        DoBloom = 256, DoFXAA = 512, DoDepthFieldBlur = 1024
    };
    int width, height;
    float exposure = 1, focus = 0, CoC = 0;
    vec3 cameraPosition;
    RmHitInfo *Gbuffer;
    RmRadiance *radiance_Dd, *radiance_Ds, *radiance_Id, *radiance_Is;
    vec3 *pixelarray;

    Photo(int width, int height);
    ~Photo();
    Photo(const Photo &) = delete;
    Photo &operator=(const Photo &) = delete;

    // The frame lives on the device after render(); these forward to the library and keep the host copies current.
    bool render(RmContext *ctx, const Model &model, const RenderArgs &args, uint64_t seed);
    // the same frame rendered by `world` processes (one GPU each): this rank's sample shard + rm_reduce to rank 0
    bool renderSharded(RmContext *ctx, const Model &model, const RenderArgs &args, uint64_t seed, int rank, int world);
    void spatialClamp();                      // src/image.cpp:30-82   -> rm_spatial_clamp
    void filter();                            // src/image.cpp:84-213  -> rm_filter
    void postProcessing(int shadeOptions);    // src/image.cpp:470-479 -> rm_postprocess
    void save(const char *file_name);         // src/image.cpp:481-529: 8-bit RGB PNG, value = byte(pixel * 255)

private:
    RmContext *ctx_ = nullptr;
    RenderArgs args_;
    void syncPlanes();
};

// Decode a PNG or DDS (DXT1/3/5, uncompressed) file to 8-bit RGBA, top row first (image_io.cpp).  On failure `why` says why.
bool loadImageRGBA(const std::string &path, int &width, int &height, std::vector<uint8_t> &rgba, std::string &why);
// Encode an 8-bit RGB image as PNG (zlib deflate, filter 0 on every row).  Returns false if the file cannot be written.
bool writePng(const char *file_name, const uint8_t *rgb, int width, int height);

void render_multiThread(Model &model, const RenderArgs &args);

class MyConsole {                 // include/myconsole.h: same commands; reads its dialogue from the stream it is given
public:
    std::map<std::string, Model> models;
    std::map<std::string, RenderArgs> renderArgs;
    MyConsole();                                       // std::cin / std::cout
    MyConsole(std::istream &in, std::ostream &out);
    void createModel(const std::string &model_id);
    void createRenderArgs(const std::string &str);
    void deleteModel(const std::string &str);
    void deleteRenderArgs(const std::string &str);
    void viewModel(const std::string &str);
    void viewRenderArgs(const std::string &str);
    void render(const std::string &model_str, const std::string &args_str);
    std::ostream &out() { return out_; }

private:
    std::istream &in_;
    std::ostream &out_;
};
void parseCommand(MyConsole &console, const std::string &opt);
// the read-dispatch loop of the console program: one command per line until `exit` or end of input
int runConsole(MyConsole &console, std::istream &in, std::ostream &out);

#endif
