// host_check.cpp — test-only driver for the C++ host side (raym0nade_b200/host): exposes the pieces the console does not
// print so tests/test_cpu_console.py can compare them with the Python side.
//   host_check png  <rgb.raw> <w> <h> <out.png>          writePng on raw 8-bit RGB
//   host_check dump <folder> <model> <sky> <out.bin>     load a Model, write its raw arrays and prepared BVH nodes
//   host_check image <file.png|.dds> <out.bin>           loadImageRGBA: int32 w, h, then w*h*4 bytes
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "host.hpp"

template <class T>
static void put(std::ofstream &o, const T *p, size_t n) { o.write(reinterpret_cast<const char *>(p), std::streamsize(n * sizeof(T))); }

int main(int argc, char **argv) {
    if (argc == 6 && !std::strcmp(argv[1], "png")) {
        const int w = std::atoi(argv[3]), h = std::atoi(argv[4]);
        std::ifstream in(argv[2], std::ios::binary);
        std::vector<uint8_t> rgb(size_t(w) * h * 3);
        in.read(reinterpret_cast<char *>(rgb.data()), std::streamsize(rgb.size()));
        return writePng(argv[5], rgb.data(), w, h) ? 0 : 1;
    }
    if (argc == 6 && !std::strcmp(argv[1], "dump")) {
        Model m(argv[2], argv[3], argv[4]);
        if (m.empty()) return 1;
        const RmSceneDesc *d = m.desc();
        std::ofstream o(argv[5], std::ios::binary);
        const int32_t hdr[8] = {int32_t(m.faceCount()), int32_t(m.meshes.size()), int32_t(m.materials.size()), int32_t(m.textures.size()),
                                m.skyWidth, m.skyHeight, d->n_nodes, d->n_lights};
        put(o, hdr, 8);
        put(o, m.positions.data(), m.positions.size());
        put(o, m.uvs.data(), m.uvs.size());
        put(o, m.normals.data(), m.normals.size());
        put(o, m.meshes.data(), m.meshes.size());
        put(o, m.materials.data(), m.materials.size());
        put(o, m.sky.data(), m.sky.size());
        put(o, d->nodes, size_t(d->n_nodes));
        put(o, d->positions, size_t(d->n_faces) * 9);          // post-build order
        for (size_t i = 0; i < m.textures.size(); i++) {
            const int32_t whc[3] = {m.textures[i].width, m.textures[i].height, m.textures[i].channels};
            put(o, whc, 3);
            put(o, m.texturePixels[i].data(), m.texturePixels[i].size());
        }
        return o ? 0 : 1;
    }
    if (argc == 4 && !std::strcmp(argv[1], "image")) {
        int w = 0, h = 0;
        std::vector<uint8_t> rgba;
        std::string why;
        if (!loadImageRGBA(argv[2], w, h, rgba, why)) { std::cerr << why << std::endl; return 1; }
        std::ofstream o(argv[3], std::ios::binary);
        const int32_t wh[2] = {w, h};
        put(o, wh, 2);
        put(o, rgba.data(), rgba.size());
        return o ? 0 : 1;
    }
    std::cerr << "usage: host_check png|dump|image ..." << std::endl;
    return 2;
}
