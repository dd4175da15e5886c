// C-ABI of the host-side stages (include/adypt_b200.h, "Host side").
#include <cfloat>
#include <cmath>
#include <cstring>
#include <memory>
#include "../common.h"
#include "host_scene.h"

using namespace adypt;
using namespace adypt::host;

namespace adypt {
namespace host {
void flat_normal(const float p0[3], const float p1[3], const float p2[3], float out[3])
{
	const float a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
	// glm::cross: (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
	const float c[3] = {a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]};
	// glm::normalize: v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x)
	const float inv = 1.0f / sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
	out[0] = c[0] * inv;
	out[1] = c[1] * inv;
	out[2] = c[2] * inv;
}
} // namespace host
} // namespace adypt

extern "C" {

int adypt_host_scene_from_triangles(const float *positions, const int32_t *material_ids, uint32_t n_tris, const void *materials64,
                                    uint32_t n_mats, adypt_host_scene **out)
{
	return guarded([&]() -> int {
	if (!positions || !material_ids || !out || n_tris == 0) return fail(ADYPT_EINVAL, "positions/material_ids/out is NULL or n_tris is 0");
	if (n_mats && !materials64) return fail(ADYPT_EINVAL, "materials is NULL");
	std::unique_ptr<adypt_host_scene> s(new adypt_host_scene);
	s->tris.resize(n_tris);
	for (uint32_t i = 0; i < n_tris; ++i) {
		Triangle &t = s->tris[i];
		memset(&t, 0, sizeof(t));
		memcpy(t.p, positions + (size_t)i * 9, 36);
		float n[3];
		flat_normal(t.p[0], t.p[1], t.p[2], n);
		for (int k = 0; k < 3; ++k) memcpy(t.n[k], n, 12);
		t.matid = material_ids[i];
		if (n_mats && (t.matid < 0 || (uint32_t)t.matid >= n_mats)) return fail(ADYPT_EINVAL, "material id out of range");
	}
	s->mats.resize(n_mats);
	if (n_mats) memcpy(s->mats.data(), materials64, (size_t)n_mats * 64);
	s->box = scene_box(s->tris.data(), s->tris.size());
	*out = s.release();
	return ADYPT_OK;
	});
}

int adypt_host_scene_load_obj(const char *obj_path, adypt_host_scene **out)
{
	return guarded([&]() -> int {
	if (!obj_path || !out) return fail(ADYPT_EINVAL, "NULL argument");
	std::unique_ptr<adypt_host_scene> s(new adypt_host_scene);
	const std::string err = load_obj(obj_path, s.get());
	if (!err.empty()) return fail(ADYPT_EIO, err);
	s->box = scene_box(s->tris.data(), s->tris.size());
	*out = s.release();
	return ADYPT_OK;
	});
}

int adypt_host_scene_destroy(adypt_host_scene *s)
{
	return guarded([&]() -> int {
	delete s;
	return ADYPT_OK;
	});
}

int adypt_host_scene_build_bvh(adypt_host_scene *s, const adypt_bvh_config *c)
{
	return guarded([&]() -> int {
	if (!s || !c) return fail(ADYPT_EINVAL, "NULL argument");
	if (s->tris.empty()) return fail(ADYPT_EINVAL, "scene has no triangles");
	BvhConfig cfg;
	cfg.max_spatial_depth = c->max_spatial_depth;
	cfg.triangle_sah = c->triangle_sah;
	cfg.node_sah = c->node_sah;
	build_binary(s->tris.data(), s->tris.size(), s->box, cfg, &s->binary);
	if (!build_wide(s->binary, cfg, &s->wide))
		return fail(ADYPT_EINVAL, "a one-triangle scene has no wide BVH (the reference builder reads out of bounds there)");
	return ADYPT_OK;
	});
}

int adypt_host_scene_load_bvh(adypt_host_scene *s, const char *path, const adypt_bvh_config *c)
{
	return guarded([&]() -> int {
	if (!s || !path || !c) return fail(ADYPT_EINVAL, "NULL argument");
	BvhConfig cfg;
	cfg.max_spatial_depth = c->max_spatial_depth;
	cfg.triangle_sah = c->triangle_sah;
	cfg.node_sah = c->node_sah;
	s->binary.nodes.clear();
	if (!load_bvh_file(path, cfg, &s->wide)) return fail(ADYPT_EIO, std::string("no usable .bvh cache at ") + path);
	return ADYPT_OK;
	});
}

int adypt_host_scene_save_bvh(adypt_host_scene *s, const char *path, const adypt_bvh_config *c)
{
	return guarded([&]() -> int {
	if (!s || !path || !c) return fail(ADYPT_EINVAL, "NULL argument");
	BvhConfig cfg;
	cfg.max_spatial_depth = c->max_spatial_depth;
	cfg.triangle_sah = c->triangle_sah;
	cfg.node_sah = c->node_sah;
	if (!save_bvh_file(path, s->wide, cfg)) return fail(ADYPT_EIO, std::string("cannot write ") + path);
	return ADYPT_OK;
	});
}

int adypt_host_scene_get(adypt_host_scene *s, adypt_host_scene_info *o)
{
	return guarded([&]() -> int {
	if (!s || !o) return fail(ADYPT_EINVAL, "NULL argument");
	o->n_tris = (uint32_t)s->tris.size();
	o->n_mats = (uint32_t)s->mats.size();
	o->n_nodes = (uint32_t)s->wide.nodes.size();
	o->n_refs = (uint32_t)s->wide.tri_indices.size();
	o->n_binary_nodes = (uint32_t)s->binary.nodes.size();
	o->triangles = s->tris.data();
	o->materials = s->mats.empty() ? nullptr : s->mats.data();
	o->nodes = s->wide.nodes.empty() ? nullptr : s->wide.nodes.data();
	o->tri_indices = s->wide.tri_indices.empty() ? nullptr : s->wide.tri_indices.data();
	o->binary_nodes = s->binary.nodes.empty() ? nullptr : s->binary.nodes.data();
	memcpy(o->aabb, s->box.lo, 12);
	memcpy(o->aabb + 3, s->box.hi, 12);
	return ADYPT_OK;
	});
}

int adypt_host_scene_load_textures(adypt_host_scene *s, uint32_t *n_loaded, uint32_t *n_failed)
{
	return guarded([&]() -> int {
	if (!s) return fail(ADYPT_EINVAL, "scene is NULL");
	uint32_t ok = 0, bad = 0;
	if (!s->textures_loaded) {
		// OglScene::init_materials walks the materials in order; load_texture caches successes by file name and
		// returns -1 for a file that fails (every time it is asked for again)
		std::vector<int> index_of(s->diffuse_textures.size(), -2); // -2 = not tried yet
		for (Material &m : s->mats) {
			if (m.dtex < 0 || (size_t)m.dtex >= s->diffuse_textures.size()) continue;
			int &slot = index_of[(size_t)m.dtex];
			if (slot == -2) {
				DecodedImage img;
				if (decode_image_file(s->diffuse_textures[(size_t)m.dtex].c_str(), &img)) {
					printf("[SCENE]Info: Loaded texture %s\n", s->diffuse_textures[(size_t)m.dtex].c_str());
					s->textures.push_back(std::move(img));
					slot = (int)s->textures.size() - 1;
				} else {
					printf("[SCENE]Err: Unable to load texture %s\n", s->diffuse_textures[(size_t)m.dtex].c_str());
					slot = -1;
				}
			}
			m.dtex = slot;
		}
		for (int v : index_of) {
			if (v >= 0) ++ok;
			else if (v == -1) ++bad;
		}
		s->textures_loaded = true;
	} else
		ok = (uint32_t)s->textures.size();
	if (n_loaded) *n_loaded = ok;
	if (n_failed) *n_failed = bad;
	return ADYPT_OK;
	});
}

int adypt_host_scene_texture(adypt_host_scene *s, uint32_t i, const uint8_t **rgb8, int32_t *width, int32_t *height)
{
	return guarded([&]() -> int {
	if (!s || !rgb8 || !width || !height) return fail(ADYPT_EINVAL, "NULL argument");
	if (i >= s->textures.size()) return fail(ADYPT_EINVAL, "texture index out of range");
	*rgb8 = s->textures[i].rgb.data();
	*width = s->textures[i].width;
	*height = s->textures[i].height;
	return ADYPT_OK;
	});
}

int adypt_host_scene_upload(adypt_host_scene *s, int32_t device, adypt_scene **out)
{
	return guarded([&]() -> int {
	if (!s || !out) return fail(ADYPT_EINVAL, "NULL argument");
	if (s->wide.nodes.empty()) return fail(ADYPT_EINVAL, "build or load a BVH first");
	adypt_scene_desc d;
	memset(&d, 0, sizeof(d));
	d.device = device;
	d.nodes = s->wide.nodes.data();
	d.n_nodes = (uint32_t)s->wide.nodes.size();
	d.tri_indices = s->wide.tri_indices.data();
	d.n_refs = (uint32_t)s->wide.tri_indices.size();
	d.woop = nullptr; // built on the GPU
	d.triangles = s->tris.data();
	d.n_tris = (uint32_t)s->tris.size();
	d.materials = s->mats.empty() ? nullptr : s->mats.data();
	d.n_mats = (uint32_t)s->mats.size();
	// textures named by the materials but never loaded: the scene then behaves like a reference build with
	// TEXTURE_COUNT == 0 (Material::dtex is ignored by the kernels when no texture was uploaded)
	ADYPT_TRY(adypt_scene_create(&d, out));
	if (!s->textures.empty()) {
		std::vector<adypt_texture> tx(s->textures.size());
		for (size_t i = 0; i < tx.size(); ++i) tx[i] = adypt_texture{s->textures[i].rgb.data(), s->textures[i].width, s->textures[i].height};
		const int rc = adypt_scene_set_textures(*out, tx.data(), (uint32_t)tx.size());
		if (rc != ADYPT_OK) {
			adypt_scene_destroy(*out);
			*out = nullptr;
			return rc;
		}
	}
	return ADYPT_OK;
	});
}

} // extern "C"
