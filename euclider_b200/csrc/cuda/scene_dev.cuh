// Device-resident flat scene.
//
// The host packs EuclFlatScene into ONE 16-byte-aligned blob (header + tables); every kernel
// stages the blob from global/L2 into shared memory with 128-bit loads and then walks the tables
// there.  Primitives are SoA (component k of v0 for primitive i lives at v0[k * n_prims + i]) so
// that lanes which diverge onto different primitives of a CSG program still hit distinct banks;
// in the brute-force entity loop all lanes read the same primitive and the load is a broadcast.
// Texels stay in HBM behind CUDA texture objects (point sampling; the bilinear filter is done in
// f64 to match the reference, surface.rs:453-489).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "euclider_b200.h"

namespace eucl {

struct SceneHeader {
    int32_t dim;
    int32_t n_prims, n_nodes, n_entities, n_materials, n_transforms, n_expr_ops;
    int32_t n_surfaces, n_color_ops, n_mapped_textures, n_textures;
    int32_t background;
    int32_t max_entity_nodes; // largest CSG program (nodes) of any entity
    int32_t blob_bytes;       // total size, multiple of 16
    // byte offsets from the start of the blob, all multiples of 16
    int32_t off_prim_kind, off_prim_v0, off_prim_v1, off_prim_s0, off_prim_s1;
    int32_t off_nodes, off_entities, off_materials, off_transforms, off_expr_ops;
    int32_t off_surfaces, off_color_ops, off_mapped, off_textures, off_tex_objects, off_perlin;
    int32_t _pad[2];
};
static_assert(sizeof(SceneHeader) % 16 == 0, "header must keep 16-byte alignment");

// Pointers into the staged copy; lives in shared memory next to the blob.
struct SceneView {
    int dim, n_prims, n_nodes, n_entities, n_surfaces, background;
    const int32_t* prim_kind;
    const double* prim_v0;
    const double* prim_v1;
    const double* prim_s0;
    const double* prim_s1;
    const EuclNode* nodes;
    const EuclEntity* entities;
    const EuclMaterial* materials;
    const EuclTransform* transforms;
    const EuclExprOp* expr_ops;
    const EuclSurface* surfaces;
    const EuclColorOp* color_ops;
    const EuclMappedTexture* mapped;
    const EuclTexture* textures;
    const cudaTextureObject_t* tex_objects;
    const uint8_t* perlin;
};

// Cooperative copy of the blob into shared memory + view construction.  `smem` must be 16-byte
// aligned and hold sizeof(SceneView) rounded to 16 + blob_bytes.
__device__ __forceinline__ const SceneView& stage_scene(const uint8_t* __restrict__ blob, unsigned char* smem) {
    SceneView* view = reinterpret_cast<SceneView*>(smem);
    unsigned char* dst = smem + ((sizeof(SceneView) + 15) & ~size_t(15));
    const SceneHeader* gh = reinterpret_cast<const SceneHeader*>(blob);
    const int n16 = gh->blob_bytes >> 4;
    const uint4* src4 = reinterpret_cast<const uint4*>(blob);
    uint4* dst4 = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst4[i] = __ldg(src4 + i);
    __syncthreads();
    if (threadIdx.x == 0) {
        const SceneHeader* h = reinterpret_cast<const SceneHeader*>(dst);
        view->dim = h->dim;
        view->n_prims = h->n_prims;
        view->n_nodes = h->n_nodes;
        view->n_entities = h->n_entities;
        view->n_surfaces = h->n_surfaces;
        view->background = h->background;
        view->prim_kind = reinterpret_cast<const int32_t*>(dst + h->off_prim_kind);
        view->prim_v0 = reinterpret_cast<const double*>(dst + h->off_prim_v0);
        view->prim_v1 = reinterpret_cast<const double*>(dst + h->off_prim_v1);
        view->prim_s0 = reinterpret_cast<const double*>(dst + h->off_prim_s0);
        view->prim_s1 = reinterpret_cast<const double*>(dst + h->off_prim_s1);
        view->nodes = reinterpret_cast<const EuclNode*>(dst + h->off_nodes);
        view->entities = reinterpret_cast<const EuclEntity*>(dst + h->off_entities);
        view->materials = reinterpret_cast<const EuclMaterial*>(dst + h->off_materials);
        view->transforms = reinterpret_cast<const EuclTransform*>(dst + h->off_transforms);
        view->expr_ops = reinterpret_cast<const EuclExprOp*>(dst + h->off_expr_ops);
        view->surfaces = reinterpret_cast<const EuclSurface*>(dst + h->off_surfaces);
        view->color_ops = reinterpret_cast<const EuclColorOp*>(dst + h->off_color_ops);
        view->mapped = reinterpret_cast<const EuclMappedTexture*>(dst + h->off_mapped);
        view->textures = reinterpret_cast<const EuclTexture*>(dst + h->off_textures);
        view->tex_objects = reinterpret_cast<const cudaTextureObject_t*>(dst + h->off_tex_objects);
        view->perlin = reinterpret_cast<const uint8_t*>(dst + h->off_perlin);
    }
    __syncthreads();
    return *view;
}

__host__ __device__ inline size_t scene_smem_bytes(int blob_bytes) {
    return ((sizeof(SceneView) + 15) & ~size_t(15)) + (size_t)blob_bytes;
}

} // namespace eucl
