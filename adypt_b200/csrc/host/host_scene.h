// Host scene: Triangle[] + materials (+ BVH once built). The CPU-side stand-in for Scene (src/Util/Scene.hpp).
#pragma once
#include <string>
#include <vector>
#include "bvh_build.h"

namespace adypt {
namespace host {
struct DecodedImage { // RGB8, top row first (what stbi_load(..., 3) returns)
	int width = 0, height = 0;
	std::vector<uint8_t> rgb;
};
// PNG / JPEG / TGA file -> RGB8; false when the file is missing or in a format that is not decoded
bool decode_image_file(const char *path, DecodedImage *out);
bool decode_jpeg(const std::vector<uint8_t> &file, DecodedImage *out); // jpeg_decode.cpp
// GIF / PSD / PIC / PNM / HDR (image_decode_more.cpp); *recognised = the file carries one of their signatures
bool decode_rare_formats(const std::vector<uint8_t> &file, DecodedImage *out, bool *recognised);
} // namespace host
} // namespace adypt

struct adypt_host_scene {
	std::vector<adypt::Triangle> tris;
	std::vector<adypt::Material> mats;
	std::vector<std::string> diffuse_textures; // deduplicated names, index = Material::dtex (before textures are loaded)
	std::vector<adypt::host::DecodedImage> textures; // after adypt_host_scene_load_textures: index = Material::dtex
	bool textures_loaded = false;
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
