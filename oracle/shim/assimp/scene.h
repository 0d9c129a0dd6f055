// Declaration-only stand-in for <assimp/scene.h>.  TEST INFRASTRUCTURE (see material.h).
#pragma once
#include "material.h"
struct aiVector3D { ai_real x, y, z; };
struct aiFace { unsigned int mNumIndices; unsigned int *mIndices; };
struct aiMesh {
    unsigned int mNumVertices, mNumFaces;
    aiVector3D *mVertices, *mNormals;
    aiVector3D *mTextureCoords[8];
    aiFace *mFaces;
    unsigned int mMaterialIndex;
};
struct aiNode;
struct aiScene {
    unsigned int mFlags;
    aiNode *mRootNode;
    unsigned int mNumMeshes;
    aiMesh **mMeshes;
    unsigned int mNumMaterials;
    aiMaterial **mMaterials;
};
#define AI_SCENE_FLAGS_INCOMPLETE 0x1
