// Test harness (not part of the product): builds the SBVH + wide BVH of a 60 000-triangle scene through the C-ABI, so that
// tests/test_builder_tsan.py can run the builder's forked subtree threads under ThreadSanitizer.
#include <cstdio>
#include <vector>
#include <cstdint>
#include <cmath>
#include <string>
#include "adypt_b200.h"
namespace adypt { int fail(int code, const std::string &) { return code; } }
extern "C" {
int adypt_scene_create(const adypt_scene_desc *, adypt_scene **) { return ADYPT_ENODEV; }
int adypt_scene_set_textures(adypt_scene *, const adypt_texture *, uint32_t) { return ADYPT_ENODEV; }
int adypt_scene_destroy(adypt_scene *) { return ADYPT_OK; }
}
int main()
{
	// a jittered grid of small triangles: big enough for the builder to fork subtrees onto threads
	const int n = 60000;
	std::vector<float> pos((size_t)n * 9);
	uint32_t s = 12345;
	auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); };
	for (int i = 0; i < n; ++i) {
		const float cx = (i % 250) * 1.0f + rnd(), cy = (i / 250) * 1.0f + rnd(), cz = rnd() * 5.0f;
		for (int k = 0; k < 3; ++k) { pos[9 * i + 3 * k] = cx + rnd() * 0.8f; pos[9 * i + 3 * k + 1] = cy + rnd() * 0.8f; pos[9 * i + 3 * k + 2] = cz + rnd() * 0.8f; }
	}
	std::vector<int32_t> mat(n, 0);
	std::vector<uint8_t> mats(64, 0);
	adypt_host_scene *sc = nullptr;
	if (adypt_host_scene_from_triangles(pos.data(), mat.data(), n, mats.data(), 1, &sc) != ADYPT_OK) return 2;
	adypt_bvh_config cfg; cfg.max_spatial_depth = 48; cfg.triangle_sah = 0.3f; cfg.node_sah = 1.0f;
	const int rc = adypt_host_scene_build_bvh(sc, &cfg);
	printf("build rc %d\n", rc);
	adypt_host_scene_destroy(sc);
	return rc;
}
