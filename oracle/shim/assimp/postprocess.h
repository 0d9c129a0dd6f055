// Declaration-only stand-in for <assimp/postprocess.h>.  TEST INFRASTRUCTURE.
#pragma once
enum aiPostProcessSteps {
    aiProcess_Triangulate = 0x8, aiProcess_PreTransformVertices = 0x100,
    aiProcess_SortByPType = 0x8000, aiProcess_FixInfacingNormals = 0x2000
};
#define AI_CONFIG_PP_PTV_KEEP_HIERARCHY "PP_PTV_KEEP_HIERARCHY"
