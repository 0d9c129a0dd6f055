// model.cpp — Model: the loaded scene (include/model.h:25-43; constructor src/model.cpp:172-215).
//
// Import (assimp + the Python image decoders in the reference) is replaced by three small readers of already-decoded
// data; the derivations that follow the import are the library's (rm_prepare_scene = BVH::build + generateMipmaps +
// checkLightObject + SkyBox::Init, raym0nade_b200/csrc/host_prep.cpp).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "host.hpp"

namespace {

bool endsWith(const std::string &s, const char *suffix) {
    const size_t n = std::strlen(suffix);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; i++)
        if (std::tolower(static_cast<unsigned char>(s[s.size() - n + i])) != suffix[i]) return false;
    return true;
}

template <class T>
bool readArray(std::istream &in, std::vector<T> &v, size_t n) {
    v.resize(n);
    if (n == 0) return true;
    in.read(reinterpret_cast<char *>(v.data()), std::streamsize(n * sizeof(T)));
    return bool(in);
}

RmRawMaterial defaultMaterial() {          // Material::Material, src/material.cpp:109-111
    RmRawMaterial m{};
    m.tex_diffuse = m.tex_specular = m.tex_emissive = m.tex_normals = -1;
    m.opacity = 1.0f; m.ior = 1.0f; m.roughness = 0.8f;
    return m;
}

// Material::loadMaterialProperties (src/material.cpp:300-328): an opacity below 0.99 makes the material a dielectric;
// a handful of material names choose the transmitting colour and the roughness.
void applyOpacity(RmRawMaterial &m, const std::string &name, float opacity) {
    if (!(opacity < 0.99f)) return;
    m.opacity = 0.0f;
    m.ior = 1.25f;
    m.roughness = 5e-3f;
    if (name == "Ice") m.roughness = 0.5f;
    auto set = [&](float r, float g, float b) { m.transmitting_color[0] = r; m.transmitting_color[1] = g; m.transmitting_color[2] = b; };
    if (name == "TransparentGlassWine") set(0.2f, 0.08f, 0.07f);
    if (name == "TransparentGlass" || name == "Water" || name == "Ice") set(1.0f, 1.0f, 1.0f);
    if (name == "Beer") set(0.8f, 0.7f, 0.55f);
    if (name == "Red_Wine") set(0.24f, 0.09f, 0.07f);
    if (name == "White_Wine") set(0.85f, 0.78f, 0.6f);
}

}  // namespace

Model::Model() = default;

Model::~Model() {
    if (prepared_) rm_prepared_free(prepared_);
}

Model::Model(const std::string &model_folder, const std::string &model_name, const std::string &skyMap_name) {
    model_path = model_folder + model_name;
    skyMap_path = (skyMap_name == "null") ? "null" : model_folder + skyMap_name;

    bool ok;
    if (endsWith(model_path, ".rmscene")) ok = loadRmScene(model_path);
    else if (endsWith(model_path, ".obj")) ok = loadObj(model_folder, model_path);
    else {
        std::cerr << "Error loading model: " << model_path << ": this host reads .rmscene and .obj (assimp is not part of it)" << std::endl;
        ok = false;
    }
    if (!ok) {                       // like the reference: say why, leave an empty model (src/model.cpp:185-188)
        positions.clear(); uvs.clear(); normals.clear(); meshes.clear();
        return;
    }
    std::cout << "Materials: " << materials.size() << std::endl;
    std::cout << "Vertices: " << faceCount() * 3 << std::endl << "faces: " << faceCount() << std::endl;

    // A .rmscene may carry its sky inside: the sky map name "embedded" keeps it, "null" renders without one
    // (src/model.cpp:178), any other name is a file next to the model.
    if (skyMap_name == "embedded") {
        skyMap_path = sky.empty() ? "null" : "embedded";
    } else {
        sky.clear(); skyWidth = skyHeight = 0;
        if (skyMap_path != "null") {
            std::cout << "Loading sky map: " << skyMap_path << std::endl;
            if (!loadSky(skyMap_path)) { sky.clear(); skyWidth = skyHeight = 0; }
        }
    }
    // (emissive textures stay loaded with a sky present: the reference tests skyMap.empty() before the sky is read,
    // src/model.cpp:142-143 vs 207-211, so light objects are always formed)

    prepare();
}

bool Model::prepare() {
    if (prepared_) { rm_prepared_free(prepared_); prepared_ = nullptr; }
    textures.resize(texturePixels.size());
    for (size_t i = 0; i < textures.size(); i++) textures[i].pixels = texturePixels[i].data();
    RmRawScene raw{};
    raw.n_faces = int32_t(faceCount());
    raw.n_meshes = int32_t(meshes.size());
    raw.n_materials = int32_t(materials.size());
    raw.n_textures = int32_t(textures.size());
    raw.positions = positions.data(); raw.uvs = uvs.data(); raw.normals = normals.data();
    raw.meshes = meshes.data(); raw.materials = materials.data(); raw.textures = textures.data();
    raw.sky_width = skyWidth; raw.sky_height = skyHeight;
    raw.sky_rgb = sky.empty() ? nullptr : sky.data();
    if (rm_prepare_scene(&raw, &prepared_) != RM_OK) {
        std::cerr << "Error loading model: " << rm_last_error() << std::endl;
        prepared_ = nullptr;
        return false;
    }
    const RmSceneDesc *d = rm_prepared_desc(prepared_);
    for (int i = 0; i < d->n_lights; i++)          // the line checkLightObject prints (src/model.cpp:78-80)
        std::cout << "Light object with power: " << d->lights[i].power << ", color:" << d->lights[i].color[0] << ", "
                  << d->lights[i].color[1] << ", " << d->lights[i].color[2] << ", position:" << d->lights[i].center[0]
                  << ", " << d->lights[i].center[1] << ", " << d->lights[i].center[2] << std::endl;
    std::cout << "BVH has built with size " << d->n_nodes << std::endl;
    return true;
}

const RmSceneDesc *Model::desc() const { return prepared_ ? rm_prepared_desc(prepared_) : nullptr; }

// ------------------------------------------------------------------------------------------------ .rmscene
// Little-endian container of an RmRawScene (include/rm_types.h), written by scenes.RawScene.save():
//   "RMSCENE1" | int32 n_faces, n_meshes, n_materials, n_textures, sky_width, sky_height
//   | f32 positions[n_faces*9] | f32 uvs[n_faces*6] | f32 normals[n_faces*9]
//   | RmRawMesh[n_meshes] | RmRawMaterial[n_materials] | per texture: int32 w, h, c, then w*h*c bytes
//   | f32 sky[sky_height*sky_width*3]
bool Model::loadRmScene(const std::string &path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { std::cerr << "Error loading model: cannot open " << path << std::endl; return false; }
    char magic[8];
    int32_t hdr[6];
    in.read(magic, 8);
    in.read(reinterpret_cast<char *>(hdr), sizeof hdr);
    if (!in || std::memcmp(magic, "RMSCENE1", 8) != 0) { std::cerr << "Error loading model: " << path << " is not an RMSCENE1 file" << std::endl; return false; }
    const int64_t nf = hdr[0], nm = hdr[1], nmat = hdr[2], ntex = hdr[3], sw = hdr[4], sh = hdr[5];
    if (nf < 0 || nm < 0 || nmat < 0 || ntex < 0 || sw < 0 || sh < 0 || (sw == 0) != (sh == 0)) {
        std::cerr << "Error loading model: " << path << ": negative or inconsistent counts in the header" << std::endl;
        return false;
    }
    // the header must not promise more than the file holds (a damaged file would otherwise ask for absurd allocations)
    const std::streampos here = in.tellg();
    in.seekg(0, std::ios::end);
    const int64_t left = int64_t(in.tellg()) - int64_t(here);
    in.seekg(here);
    const int64_t fixed = nf * 96 + nm * int64_t(sizeof(RmRawMesh)) + nmat * int64_t(sizeof(RmRawMaterial)) + ntex * 12 + sw * sh * 12;
    if (fixed > left) { std::cerr << "Error loading model: " << path << " is truncated or malformed" << std::endl; return false; }
    bool ok = readArray(in, positions, size_t(nf) * 9) && readArray(in, uvs, size_t(nf) * 6) && readArray(in, normals, size_t(nf) * 9) &&
              readArray(in, meshes, size_t(nm)) && readArray(in, materials, size_t(nmat));
    texturePixels.clear();
    textures.clear();
    for (int64_t i = 0; ok && i < ntex; i++) {
        int32_t whc[3];
        in.read(reinterpret_cast<char *>(whc), sizeof whc);
        if (!in || whc[0] <= 0 || whc[1] <= 0 || (whc[2] != 3 && whc[2] != 4)) { ok = false; break; }
        RmRawTexture t{};
        t.width = whc[0]; t.height = whc[1]; t.channels = whc[2];
        textures.push_back(t);
        texturePixels.emplace_back();
        const int64_t bytes = int64_t(whc[0]) * whc[1] * whc[2];
        if (bytes > left) { ok = false; break; }
        ok = readArray(in, texturePixels.back(), size_t(bytes));
    }
    if (ok && sw > 0) {
        ok = readArray(in, sky, size_t(sw) * size_t(sh) * 3);
        skyWidth = int(sw); skyHeight = int(sh);
    }
    if (!ok) { std::cerr << "Error loading model: " << path << " is truncated or malformed" << std::endl; return false; }
    materialNames.assign(materials.size(), std::string());
    return true;
}

// ------------------------------------------------------------------------------------------------ .obj / .mtl
// Geometry: v / vt / vn / f (polygons are fanned like aiProcess_Triangulate does for convex faces), one mesh per material
// in order of first use (assimp splits an object per material; light objects are formed per mesh, src/model.cpp:121-123).
// Materials: Kd and Ke become 1x1 RGBA8 textures (Kd stored through the inverse of the 2.2 decode that
// Material::getDiffuseColor applies, src/material.cpp:337-352), d / Tr become the opacity that loadMaterialProperties
// reads.  map_Kd / map_Ks / map_Ke / map_Bump name PNG or DDS files (image_io.cpp), stored RGBA8 - RGB8 for normal maps -
// so that the slot's fetch type matches the data (the reference keeps a PNG as RGB8 in every slot and then strides it
// by four, src/material.cpp:58,273-281; that misread is not reproduced).
bool Model::loadObj(const std::string &folder, const std::string &path) {
    std::ifstream in(path);
    if (!in) { std::cerr << "Error loading model: cannot open " << path << std::endl; return false; }
    std::vector<float> P, T, N;
    struct Corner { int v, t, n; };
    std::vector<std::vector<Corner>> perMat;           // triangles (3 corners each) per material
    std::map<std::string, int> matIndex;
    std::vector<float> opacities;
    auto newMaterial = [&](const std::string &name) {
        matIndex[name] = int(materials.size());
        materials.push_back(defaultMaterial());
        materialNames.push_back(name);
        opacities.push_back(1.0f);
        perMat.emplace_back();
        return int(materials.size()) - 1;
    };
    auto texel = [&](float r, float g, float b, bool encode) {
        auto q = [&](float c) {
            c = std::min(std::max(c, 0.0f), 1.0f);
            if (encode) c = std::pow(c, 1.0f / 2.2f);
            return uint8_t(std::lround(c * 255.0f));
        };
        RmRawTexture t{};
        t.width = t.height = 1; t.channels = 4;
        textures.push_back(t);
        texturePixels.push_back({q(r), q(g), q(b), 255});
        return int(textures.size()) - 1;
    };
    auto loadMtl = [&](const std::string &library) {
        std::ifstream m(library);
        if (!m) { std::cerr << "Could not open material library " << library << std::endl; return; }
        std::string line, key;
        int cur = -1;
        while (std::getline(m, line)) {
            std::istringstream ls(line);
            if (!(ls >> key)) continue;
            if (key == "newmtl") { std::string name; ls >> name; std::cout << "Material Name: " << name << std::endl; cur = newMaterial(name); }
            else if (cur < 0) continue;
            else if (key == "Kd") { float r = 0, g = 0, b = 0; ls >> r >> g >> b; materials[cur].tex_diffuse = texel(r, g, b, true); }
            else if (key == "Ke") { float r = 0, g = 0, b = 0; ls >> r >> g >> b; if (r > 0 || g > 0 || b > 0) materials[cur].tex_emissive = texel(r, g, b, false); }
            else if (key == "d") { float d = 1; ls >> d; opacities[cur] = d; }
            else if (key == "Tr") { float tr = 0; ls >> tr; opacities[cur] = 1.0f - tr; }
            else if (key.rfind("map_", 0) == 0 || key == "bump" || key == "norm") {
                // slot: 0 diffuse, 1 specular (G = roughness, B = metallic, src/material.cpp:374-383), 2 emissive, 3 normals
                const int slot = key == "map_Kd" ? 0 : key == "map_Ks" ? 1 : key == "map_Ke" ? 2
                               : (key == "map_Bump" || key == "map_bump" || key == "bump" || key == "norm" || key == "map_Kn") ? 3 : -1;
                std::string file, tok;
                while (ls >> tok) file = tok;                     // options (-bm 1.0 ...) precede the file name
                if (slot < 0 || file.empty()) continue;
                std::replace(file.begin(), file.end(), '\\', '/');  // src/model.cpp:155
                std::cout << "- Texture path (" << slot << "): " << file << std::endl;
                int w = 0, h = 0;
                std::vector<uint8_t> rgba;
                std::string why;
                if (!loadImageRGBA(folder + file, w, h, rgba, why)) { std::cerr << "Could not load " << folder + file << ": " << why << std::endl; continue; }
                RmRawTexture t{};
                t.width = w; t.height = h; t.channels = slot == 3 ? 3 : 4;      // the fetch type of the slot (src/material.cpp:58)
                textures.push_back(t);
                if (slot == 3) {
                    std::vector<uint8_t> rgb(size_t(w) * h * 3);
                    for (size_t i = 0; i < size_t(w) * h; i++) std::memcpy(&rgb[i * 3], &rgba[i * 4], 3);
                    texturePixels.push_back(std::move(rgb));
                } else texturePixels.push_back(std::move(rgba));
                int32_t *slots[4] = {&materials[cur].tex_diffuse, &materials[cur].tex_specular, &materials[cur].tex_emissive, &materials[cur].tex_normals};
                *slots[slot] = int(textures.size()) - 1;          // a map replaces the constant of the same slot (Kd / Ke)
            }
        }
    };

    int cur = -1;
    std::string line, key;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        if (!(ls >> key)) continue;
        if (key == "v") { float x = 0, y = 0, z = 0; ls >> x >> y >> z; P.insert(P.end(), {x, y, z}); }
        else if (key == "vt") { float u = 0, v = 0; ls >> u >> v; T.insert(T.end(), {u, v}); }
        else if (key == "vn") { float x = 0, y = 0, z = 0; ls >> x >> y >> z; N.insert(N.end(), {x, y, z}); }
        else if (key == "mtllib") { std::string f; ls >> f; loadMtl(folder + f); }
        else if (key == "usemtl") {
            std::string name; ls >> name;
            auto it = matIndex.find(name);
            cur = (it == matIndex.end()) ? newMaterial(name) : it->second;
        } else if (key == "f") {
            if (cur < 0) cur = newMaterial("default");
            std::vector<Corner> poly;
            std::string tok;
            while (ls >> tok) {
                Corner c{0, 0, 0};
                int *slot[3] = {&c.v, &c.t, &c.n};
                size_t pos = 0;
                for (int k = 0; k < 3 && pos <= tok.size(); k++) {
                    size_t e = tok.find('/', pos);
                    if (e == std::string::npos) e = tok.size();
                    if (e > pos) *slot[k] = std::atoi(tok.substr(pos, e - pos).c_str());
                    pos = e + 1;
                }
                const int counts[3] = {int(P.size() / 3), int(T.size() / 2), int(N.size() / 3)};
                for (int k = 0; k < 3; k++) {             // 1-based, negative = relative to the end; 0 = absent
                    int &i = *slot[k];
                    i = i < 0 ? counts[k] + i : i - 1;
                    if (i >= counts[k]) i = -1;
                }
                if (c.v < 0) { std::cerr << "Error loading model: " << path << ": face refers to a missing vertex" << std::endl; return false; }
                poly.push_back(c);
            }
            for (size_t k = 2; k < poly.size(); k++) {
                perMat[cur].push_back(poly[0]); perMat[cur].push_back(poly[k - 1]); perMat[cur].push_back(poly[k]);
            }
        }
    }
    for (size_t m = 0; m < materials.size(); m++) applyOpacity(materials[m], materialNames[m], opacities[m]);

    for (size_t m = 0; m < perMat.size(); m++) {
        if (perMat[m].empty()) continue;
        RmRawMesh mesh{int32_t(faceCount()), 0, int32_t(m)};
        for (size_t f = 0; f + 2 < perMat[m].size(); f += 3) {
            const Corner *c = &perMat[m][f];
            const float *p0 = &P[size_t(c[0].v) * 3], *p1 = &P[size_t(c[1].v) * 3], *p2 = &P[size_t(c[2].v) * 3];
            float e1[3], e2[3], fn[3];
            for (int k = 0; k < 3; k++) { e1[k] = p1[k] - p0[k]; e2[k] = p2[k] - p0[k]; }
            fn[0] = e1[1] * e2[2] - e1[2] * e2[1]; fn[1] = e1[2] * e2[0] - e1[0] * e2[2]; fn[2] = e1[0] * e2[1] - e1[1] * e2[0];
            const float len = std::sqrt(fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2]);
            if (len > 0) for (float &x : fn) x /= len;
            for (int k = 0; k < 3; k++) {
                const float *p = &P[size_t(c[k].v) * 3];
                positions.insert(positions.end(), p, p + 3);
                if (c[k].t >= 0) uvs.insert(uvs.end(), &T[size_t(c[k].t) * 2], &T[size_t(c[k].t) * 2] + 2);
                else uvs.insert(uvs.end(), {0.0f, 0.0f});
                if (c[k].n >= 0) normals.insert(normals.end(), &N[size_t(c[k].n) * 3], &N[size_t(c[k].n) * 3] + 3);
                else normals.insert(normals.end(), fn, fn + 3);          // no vn: the geometric normal
            }
        }
        mesh.face_end = int32_t(faceCount());
        meshes.push_back(mesh);
    }
    if (faceCount() == 0) { std::cerr << "Error loading model: " << path << " has no faces" << std::endl; return false; }
    return true;
}

// ------------------------------------------------------------------------------------------------ sky
// Rows top to bottom, RGB fp32, as hdr_to_array hands them to SkyBox::load (scripts/hdr_to_array.py, src/component.cpp:69-117).
namespace {

bool loadPfm(const std::string &path, std::vector<float> &rgb, int &w, int &h) {
    std::ifstream in(path, std::ios::binary);
    std::string tag;
    float scale = 0;
    if (!(in >> tag >> w >> h >> scale) || tag != "PF" || w <= 0 || h <= 0) return false;
    in.get();                                             // the single whitespace byte after the header
    if (scale > 0) return false;                          // big-endian files are not produced by anything we use
    std::vector<float> rows;
    if (!readArray(in, rows, size_t(w) * size_t(h) * 3)) return false;
    rgb.resize(rows.size());
    for (int y = 0; y < h; y++)                           // PFM stores the bottom row first
        std::memcpy(&rgb[size_t(y) * w * 3], &rows[size_t(h - 1 - y) * w * 3], size_t(w) * 3 * sizeof(float));
    return true;
}

// Radiance RGBE (-Y h +X w), flat or new-style run-length encoded scanlines.
bool loadHdr(const std::string &path, std::vector<float> &rgb, int &w, int &h) {
    std::ifstream in(path, std::ios::binary);
    std::string line;
    if (!std::getline(in, line) || line.compare(0, 2, "#?") != 0) return false;
    while (std::getline(in, line) && !line.empty()) {}    // header ends with an empty line
    if (!std::getline(in, line) || std::sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) return false;
    rgb.resize(size_t(w) * size_t(h) * 3);
    std::vector<uint8_t> scan(size_t(w) * 4);
    for (int y = 0; y < h; y++) {
        uint8_t head[4];
        in.read(reinterpret_cast<char *>(head), 4);
        if (!in) return false;
        if (w >= 8 && w < 32768 && head[0] == 2 && head[1] == 2 && ((head[2] << 8) | head[3]) == w) {
            for (int ch = 0; ch < 4; ch++)
                for (int x = 0; x < w;) {
                    int code = in.get();
                    if (code < 0) return false;
                    if (code > 128) {
                        const int run = code - 128, val = in.get();
                        if (val < 0 || x + run > w) return false;
                        for (int k = 0; k < run; k++) scan[size_t(x++) * 4 + ch] = uint8_t(val);
                    } else {
                        if (code == 0 || x + code > w) return false;
                        for (int k = 0; k < code; k++) { const int val = in.get(); if (val < 0) return false; scan[size_t(x++) * 4 + ch] = uint8_t(val); }
                    }
                }
        } else {
            std::memcpy(scan.data(), head, 4);
            in.read(reinterpret_cast<char *>(scan.data() + 4), std::streamsize(size_t(w - 1) * 4));
            if (!in) return false;
        }
        for (int x = 0; x < w; x++) {
            const uint8_t *p = &scan[size_t(x) * 4];
            const float f = p[3] ? std::ldexp(1.0f, int(p[3]) - (128 + 8)) : 0.0f;
            float *o = &rgb[(size_t(y) * w + x) * 3];
            o[0] = p[0] * f; o[1] = p[1] * f; o[2] = p[2] * f;
        }
    }
    return true;
}

}  // namespace

bool Model::loadSky(const std::string &path) {
    std::cout << "Loading HDR image from file: " << path << std::endl;
    bool ok = endsWith(path, ".pfm") ? loadPfm(path, sky, skyWidth, skyHeight) : loadHdr(path, sky, skyWidth, skyHeight);
    if (!ok) { std::cerr << "Failed to read the sky map " << path << " (Radiance .hdr or little-endian .pfm)" << std::endl; return false; }
    std::cout << "Image dimensions: " << skyWidth << "x" << skyHeight << std::endl;
    return true;
}
