// gpu_bridge.h — see gpu_bridge.cpp.  Lives next to it in the Raym0nade tree (include/gpu_bridge.h).
#ifndef GPU_BRIDGE_H
#define GPU_BRIDGE_H

#include <cstdint>
#include <vector>

#include "render.h"
#include "image.h"
#include "raym0nade_b200.h"

struct B200Bridge {
    // storage the descriptor points into (the Model itself is pointed into wherever its layout already fits)
    std::vector<float> positions, uvs, normals;
    std::vector<int32_t> faceMaterial;
    std::vector<RmMaterialDesc> materials;
    std::vector<RmTextureDesc> textures;
    std::vector<RmLightDesc> lights;
    std::vector<std::vector<float>> lightArrays;
    std::vector<std::vector<uint8_t>> texelStorage;     // levels whose ImageData buffer is shorter than the slot's fetch stride needs (see fill)
    RmSceneDesc desc{};

    // view a loaded Model as the library's post-load scene; valid while both the Model and this object live
    const RmSceneDesc *fill(const Model &m);
};

// drop-in body for render_multiThread (src/render.cpp:593-634): fills `photo` on the GPU
void render_multiThread_b200(Model &model, const RenderArgs &args, Photo &photo);

#endif
