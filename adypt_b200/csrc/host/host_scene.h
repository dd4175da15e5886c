// Host scene: Triangle[] + materials (+ BVH once built). The CPU-side stand-in for Scene (src/Util/Scene.hpp).
#pragma once
#include <string>
#include <vector>
#include "bvh_build.h"

struct adypt_host_scene {
	std::vector<adypt::Triangle> tris;
	std::vector<adypt::Material> mats;
	std::vector<std::string> diffuse_textures; // deduplicated names, index = Material::dtex
	adypt::host::Box box;
	adypt::host::BinaryBvh binary;
	adypt::host::WideBvh wide;
};

namespace adypt {
namespace host {
// flat normal of Scene.cpp:117-123: glm::normalize(glm::cross(p1 - p0, p2 - p0))
void flat_normal(const float p0[3], const float p1[3], const float p2[3], float out[3]);
// Scene::LoadFromFile; returns an error string (empty on success)
std::string load_obj(const char *path, adypt_host_scene *out);
} // namespace host
} // namespace adypt
