// Test harness (not part of the product): runs the host-side file parsers over every file of a directory --
// *.obj through the OBJ / MTL loader, *.config through the .config reader, *.bvh through the .bvh cache reader.
// tests/test_host_parsers_sanitized.py builds it with -fsanitize=address,undefined and feeds it mutated files.
#include <cstdio>
#include <cstring>
#include <dirent.h>
#include <new>
#include <string>
#include "adypt_b200.h"
#include "host/bvh_build.h"
#include "host/host_scene.h"

namespace adypt {
int fail(int code, const std::string &) { return code; } // the library defines this next to its CUDA code
} // namespace adypt

// host_api.cpp's upload entry point calls into the CUDA side of the library; the harness never uploads
extern "C" {
int adypt_scene_create(const adypt_scene_desc *, adypt_scene **) { return ADYPT_ENODEV; }
int adypt_scene_set_textures(adypt_scene *, const adypt_texture *, uint32_t) { return ADYPT_ENODEV; }
int adypt_scene_destroy(adypt_scene *) { return ADYPT_OK; }
}

static bool ends_with(const std::string &s, const char *suffix)
{
	const size_t n = strlen(suffix);
	return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

int main(int, char **argv)
{
	DIR *d = opendir(argv[1]);
	if (!d) return 2;
	int ok = 0, bad = 0, built = 0;
	while (dirent *e = readdir(d)) {
		if (e->d_name[0] == '.') continue;
		const std::string p = std::string(argv[1]) + "/" + e->d_name;
		bool good = false;
		try {
			if (ends_with(p, ".obj")) {
				// through the C-ABI: load, then build the SBVH + wide BVH over whatever coordinates the file held (NaN, inf, ...)
				adypt_host_scene *scene = nullptr;
				good = adypt_host_scene_load_obj(p.c_str(), &scene) == ADYPT_OK;
				if (good) {
					adypt_bvh_config cfg;
					cfg.max_spatial_depth = 48;
					cfg.triangle_sah = 0.3f;
					cfg.node_sah = 1.0f;
					if (adypt_host_scene_build_bvh(scene, &cfg) == ADYPT_OK) ++built;
				}
				adypt_host_scene_destroy(scene);
			} else if (ends_with(p, ".config")) {
				adypt_instance_config c;
				good = adypt_config_load(p.c_str(), &c) == ADYPT_OK;
			} else if (ends_with(p, ".bvh")) {
				adypt::host::WideBvh w;
				adypt::host::BvhConfig cfg;
				cfg.max_spatial_depth = 48;
				cfg.triangle_sah = 0.3f;
				cfg.node_sah = 1.0f;
				good = adypt::host::load_bvh_file(p.c_str(), cfg, &w);
			} else
				continue;
		} catch (const std::bad_alloc &) {
			good = false;
		}
		good ? ++ok : ++bad;
	}
	printf("accepted %d refused %d built %d\n", ok, bad, built);
	return 0;
}
