// Host-side declarations shared by the C-ABI translation units (no CUDA types here).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace mplb {

struct TriangleSoup {
    std::vector<double> tris;  // [n][3][3]
    double center[3] = {0, 0, 0};
    size_t size() const { return tris.size() / 9; }
};

// load_mesh.hpp:232-293 replacement; throws std::invalid_argument like the reference.
TriangleSoup loadCollada(const std::string &path, bool shiftToCenter, bool identityRootTransform);

// One BVH node, 64 bytes = 4 x float4, read on the device with four 128-bit loads:
//   q[k] = (axis_k.x, axis_k.y, axis_k.z, center[k])  k = 0..2   (axis_k = k-th box axis, unit)
//   q[3] = (extent.x, extent.y, extent.z, link)        link (int bits): >= 0 first child index
//          (children are link, link+1), < 0 leaf holding triangle -(link+1)
struct alignas(16) BvhNode {
    float q[4][4];
};
static_assert(sizeof(BvhNode) == 64, "node must be 64 bytes");

// One triangle for the FP32 filter: 64 bytes = 4 x float4 (xyz + pad; the 4th is padding so that a triangle
// is two aligned 32-byte sectors: one 256-bit and one 128-bit load):
//   v[0] = vertex 0, v[1] = vertex 1 - vertex 0, v[2] = vertex 2 - vertex 0 (differences formed in FP64).
struct alignas(32) TriF32 {
    float v[4][4];
};
static_assert(sizeof(TriF32) == 64, "triangle must be 64 bytes");

struct FlatBvh {
    std::vector<BvhNode> nodes;   // 2n nodes: root = 0, node 1 is padding, sibling pairs start at even indices
    std::vector<TriF32> tris32;   // FP32 rounded copy, index = original triangle id
    std::vector<double> tris64;   // [n][9], the exact input
    int depth = 0;                // max root->leaf edge count
    double absMax = 0;            // max |coordinate|
    double radius = 0;            // max vertex distance from the origin
};

// Top-down OBB hierarchy: PCA-oriented boxes, median split of triangle centroids along the
// longest box axis (balanced => depth <= ceil(log2 n), which bounds the device pair stack).
FlatBvh buildBvh(const double *tris, size_t n);

// The same two through the on-disk cache (bvh_cache.cpp): identical results, looked up by content when a cache
// directory is set (mplb_set_cache_dir / MPLB_CACHE_DIR), plain calls otherwise.
FlatBvh buildBvhCached(const double *tris, size_t n);
TriangleSoup loadColladaCached(const std::string &path, bool shiftToCenter, bool identityRootTransform);
int setCacheDir(const char *dir);   // 0, or -1 when the directory cannot be created
void cacheStats(uint64_t *hits, uint64_t *misses);

}  // namespace mplb
