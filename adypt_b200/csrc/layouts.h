// Data layouts shared by host and device code. These are the reference's GPU ABI and
// are uploaded unchanged.
#pragma once
#include <stdint.h>

namespace adypt {

// WideBVHNode (src/BVH/WideBVH.hpp:13-26) viewed as the five 128-bit words of struct Node
// (shaders/traversal.glsl:1-5).
struct alignas(16) Node {
	float px, py, pz;
	uint32_t head_w;      // ex | ey<<8 | ez<<16 | imask<<24
	uint32_t child_base;  // m_base_meta.x
	uint32_t tri_base;    // m_base_meta.y
	uint32_t meta_lo, meta_hi;
	uint32_t lox_lo, lox_hi, loy_lo, loy_hi;  // m_lox_loy
	uint32_t loz_lo, loz_hi, hix_lo, hix_hi;  // m_loz_hix
	uint32_t hiy_lo, hiy_hi, hiz_lo, hiz_hi;  // m_hiy_hiz
};
static_assert(sizeof(Node) == 80, "CWBVH node must be 80 bytes");

// struct Woop (traversal.glsl:6): m0 = (row2.xyz, -row2.w), m1 = row0, m2 = row1 of the inverse of
// [v0-v2, v1-v2, (v0-v2)x(v1-v2), v2] (OglScene.cpp:93-116)
struct alignas(16) Woop {
	float m0[4], m1[4], m2[4];
};
static_assert(sizeof(Woop) == 48, "Woop must be 48 bytes");

// Triangle (src/Util/Shape.hpp:70-88 == pathtracer.glsl:2-8); stride 100, only 4-byte aligned
struct Triangle {
	float p[3][3], n[3][3], tc[3][2];
	int32_t matid;
};
static_assert(sizeof(Triangle) == 100, "Triangle must be 100 bytes");

// GPUMaterial (src/Tracer/OglScene.hpp:19-28 == pathtracer.glsl:9-18)
struct alignas(16) Material {
	int32_t dtex; float dr, dg, db;
	int32_t etex; float er, eg, eb;
	int32_t stex; float sr, sg, sb;
	int32_t illum; float shininess, dissolve, ior;
};
static_assert(sizeof(Material) == 64, "Material must be 64 bytes");

// batch ray: vec4 origin_tmin + vec3 dir (+pad) of BVHIntersection (traversal.glsl:14)
struct alignas(16) Ray {
	float ox, oy, oz, tmin, dx, dy, dz, pad;
};
static_assert(sizeof(Ray) == 32, "Ray must be 32 bytes");

} // namespace adypt
