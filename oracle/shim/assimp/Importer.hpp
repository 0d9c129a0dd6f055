// Declaration-only stand-in for <assimp/Importer.hpp>.  TEST INFRASTRUCTURE.
// The importer always "fails": the reference's file-loading constructor is never used.
#pragma once
#include "scene.h"
namespace Assimp {
class Importer {
public:
    bool SetPropertyInteger(const char *, int) { return true; }
    const aiScene *ReadFile(const char *, unsigned int) { return nullptr; }
    const char *GetErrorString() const { return "assimp is not available (shim)"; }
    void FreeScene() {}
};
}
