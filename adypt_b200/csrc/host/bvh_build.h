// From-scratch CPU builder for the reference's acceleration structure:
// triangles -> binary SBVH with spatial splits -> 8-wide compressed BVH (80-byte nodes + leaf-order index
// array). Output is BYTE-IDENTICAL to the reference's src/BVH pipeline on the same triangles (tests memcmp
// against oracle/_ref), which is what lets the traversal kernels keep the reference's arrays "uploaded
// unchanged" without carrying the reference's code.
#pragma once
#include <stdint.h>
#include <vector>
#include "../layouts.h"

namespace adypt {
namespace host {

// InstanceConfig::BVH (src/InstanceConfig.hpp:15-20): 12 bytes, also the header of the .bvh cache file
struct BvhConfig {
	int32_t max_spatial_depth = 48;
	float triangle_sah = 0.3f;
	float node_sah = 1.0f;
};
static_assert(sizeof(BvhConfig) == 12, "InstanceConfig::BVH is 12 bytes");

// axis-aligned box with glm's min/max argument conventions (they decide the sign of a zero)
struct Box {
	float lo[3], hi[3];
};

// SBVHNode (src/BVH/SBVH.hpp:11-16): leaf iff left == -1; right child is always index + 1
struct BinaryNode {
	Box box;
	int32_t tri;
	int32_t left;
};
static_assert(sizeof(BinaryNode) == 32, "SBVHNode is 32 bytes");

struct BinaryBvh {
	std::vector<BinaryNode> nodes;
	int32_t leaf_count = 0;
};

struct WideBvh {
	std::vector<Node> nodes;          // 80-byte CWBVH nodes
	std::vector<int32_t> tri_indices; // leaf order -> scene triangle id
};

Box triangle_box(const Triangle &t);
Box scene_box(const Triangle *tris, size_t n);

// SBVHBuilder::Run (src/BVH/SBVHBuilder.cpp:46-71). Needs at least one triangle.
void build_binary(const Triangle *tris, size_t n_tris, const Box &scene, const BvhConfig &cfg, BinaryBvh *out);
// WideBVHBuilder::Run (src/BVH/WideBVHBuilder.cpp:8-22). The reference dereferences child -1 when the
// binary root is a leaf (a 1-triangle scene); this builder reports that case instead: returns false.
bool build_wide(const BinaryBvh &sbvh, const BvhConfig &cfg, WideBvh *out);

// .bvh cache file (src/BVH/WideBVH.cpp:9-66): "CWBVH_1.0\0" | BvhConfig | u32 n_idx | i32[n_idx] | nodes to EOF
bool save_bvh_file(const char *path, const WideBvh &bvh, const BvhConfig &cfg);
// false when the file is missing, has another magic, or was built with other parameters (WideBVH.cpp:42-45)
bool load_bvh_file(const char *path, const BvhConfig &expected, WideBvh *out);

} // namespace host
} // namespace adypt
