// Declaration-only stand-in for <assimp/material.h>.  TEST INFRASTRUCTURE.
// Lets the reference's translation units compile unmodified in an image that has
// no assimp; nothing declared here is ever called by the oracle harness (scenes are
// injected through Model's public members, see oracle/ref_harness.cpp).
#pragma once
#include <cstring>
#include <string>
typedef float ai_real;
enum aiReturn { aiReturn_SUCCESS = 0, aiReturn_FAILURE = -1 };
#define AI_SUCCESS aiReturn_SUCCESS
enum aiTextureType {
    aiTextureType_NONE = 0, aiTextureType_DIFFUSE = 1, aiTextureType_SPECULAR = 2,
    aiTextureType_AMBIENT = 3, aiTextureType_EMISSIVE = 4, aiTextureType_HEIGHT = 5,
    aiTextureType_NORMALS = 6, aiTextureType_SHININESS = 7, aiTextureType_OPACITY = 8,
    aiTextureType_DISPLACEMENT = 9, aiTextureType_LIGHTMAP = 10, aiTextureType_REFLECTION = 11,
    aiTextureType_BASE_COLOR = 12, aiTextureType_NORMAL_CAMERA = 13,
    aiTextureType_EMISSION_COLOR = 14, aiTextureType_METALNESS = 15,
    aiTextureType_DIFFUSE_ROUGHNESS = 16, aiTextureType_AMBIENT_OCCLUSION = 17,
    aiTextureType_UNKNOWN = 18
};
#define AI_TEXTURE_TYPE_MAX aiTextureType_UNKNOWN
struct aiString {
    unsigned int length = 0;
    char data[1024] = {0};
    const char *C_Str() const { return data; }
};
#define AI_MATKEY_NAME "?mat.name", 0, 0
#define AI_MATKEY_OPACITY "$mat.opacity", 0, 0
struct aiMaterial {
    template <typename T>
    aiReturn Get(const char *, unsigned int, unsigned int, T &) const { return aiReturn_FAILURE; }
    unsigned int GetTextureCount(aiTextureType) const { return 0; }
    aiReturn GetTexture(aiTextureType, unsigned int, aiString *) const { return aiReturn_FAILURE; }
};
