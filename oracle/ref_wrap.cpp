// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// C wrapper around the UNMODIFIED reference CPU pipeline, compiled IN PLACE from
// /root/reference (see oracle/Makefile; nothing from the reference is copied into this repo):
//   src/Util/Scene.cpp      OBJ/MTL -> Triangle[] + AABB         (Scene.cpp:9-136)
//   src/BVH/SBVHBuilder.cpp SBVH build with spatial splits        (SBVHBuilder.cpp:8-71)
//   src/BVH/WideBVHBuilder.cpp collapse to 80-byte CWBVH nodes    (WideBVHBuilder.cpp:8-273)
//   src/BVH/WideBVH.cpp     .bvh cache file I/O                    (WideBVH.cpp:9-66)
//   src/Util/Sobol.cpp      Gray-code Sobol generator              (Sobol.cpp:16-21)
//   src/InstanceConfig.cpp  .config JSON reader/writer             (InstanceConfig.cpp:10-192)
// The GL-bound pieces of the reference (OglScene.cpp, Camera.cpp) cannot compile here, so the
// three small host functions they contain are restated below against the reference's own
// vendored glm, each citing the lines it follows.
//
// Output: oracle/_ref/libadypt_ref.so (git-ignored; travels to the GPU box with the snapshot).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#include <fcntl.h>

#include "Util/Scene.hpp"
#include "Util/Sobol.hpp"
#include "BVH/SBVH.hpp"
#include "BVH/SBVHBuilder.hpp"
#include "BVH/WideBVH.hpp"
#include "BVH/WideBVHBuilder.hpp"
#include "InstanceConfig.hpp"
#include <glm/gtc/matrix_transform.hpp>
#include <rapidjson/internal/dtoa.h>
#define STB_IMAGE_IMPLEMENTATION
#include <stb_image.h> // the reference's texture decoder (OglScene.cpp:9-10, 24), vendored under dep/
#define TINYEXR_IMPLEMENTATION
#include <tinyexr.h>   // the reference's image writer (OglPathTracer.cpp:8-9, 206), vendored under dep/

static_assert(sizeof(WideBVHNode) == 80, "CWBVH node ABI");
static_assert(sizeof(Triangle) == 100, "Triangle ABI");

namespace {

// GPUMaterial as declared in OglScene.hpp:19-28 (64 bytes).
struct RefMaterial {
	int32_t m_dtex; float m_dr, m_dg, m_db;
	int32_t m_etex; float m_er, m_eg, m_eb;
	int32_t m_stex; float m_sr, m_sg, m_sb;
	int32_t m_illum;
	float m_shininess, m_dissolve, m_refraction_index;
};
static_assert(sizeof(RefMaterial) == 64, "GPUMaterial ABI");

struct RefScene {
	Scene scene;
	SBVH sbvh;
	WideBVH wbvh;
	std::vector<glm::vec4> woop;
	std::vector<RefMaterial> materials;
	std::vector<std::string> dtex_names;
};

// The reference prints one line per SBVH leaf (SBVHBuilder.hpp:91); silence stdout while it runs.
struct StdoutMute {
	int saved;
	StdoutMute() {
		fflush(stdout);
		saved = dup(1);
		int nul = open("/dev/null", O_WRONLY);
		dup2(nul, 1);
		close(nul);
	}
	~StdoutMute() {
		fflush(stdout);
		dup2(saved, 1);
		close(saved);
	}
};

// Restates OglScene::init_triangles (OglScene.cpp:93-116) with the reference's glm::inverse.
void init_triangles(const Scene &scene, const WideBVH &bvh, std::vector<glm::vec4> *tri_matrices)
{
	tri_matrices->clear();
	tri_matrices->reserve(bvh.GetTriIndices().size() * 3u);
	for (int32_t t : bvh.GetTriIndices()) {
		const Triangle &tri = scene.GetTriangles()[t];
		const glm::vec3 &v0 = tri.m_positions[0], &v1 = tri.m_positions[1], &v2 = tri.m_positions[2];
		glm::vec4 c0{v0 - v2, 0.0f};
		glm::vec4 c1{v1 - v2, 0.0f};
		glm::vec4 c2{glm::cross(v0 - v2, v1 - v2), 0.0f};
		glm::vec4 c3{v2, 1.0f};
		glm::mat4 mtx{c0.x, c1.x, c2.x, c3.x, c0.y, c1.y, c2.y, c3.y,
		              c0.z, c1.z, c2.z, c3.z, c0.w, c1.w, c2.w, c3.w};
		mtx = glm::inverse(mtx);
		tri_matrices->emplace_back(mtx[2][0], mtx[2][1], mtx[2][2], -mtx[2][3]);
		tri_matrices->emplace_back(mtx[0][0], mtx[0][1], mtx[0][2], mtx[0][3]);
		tri_matrices->emplace_back(mtx[1][0], mtx[1][1], mtx[1][2], mtx[1][3]);
	}
}

// Restates OglScene::init_materials (OglScene.cpp:51-91) minus the GL texture upload: a material
// with a diffuse texture gets m_dtex = ordinal of its (deduplicated) texture name and leaves
// m_dr/g/b unset (here: zero); etex/stex are never written by the reference (here: zero).
void init_materials(const Scene &scene, std::vector<RefMaterial> *materials, std::vector<std::string> *names)
{
	materials->clear();
	names->clear();
	for (const auto &ml : scene.GetTinyobjMaterials()) {
		RefMaterial gml;
		memset(&gml, 0, sizeof(gml));
		if (!ml.diffuse_texname.empty()) {
			std::string full = scene.GetBasePath() + ml.diffuse_texname;
			int idx = -1;
			for (size_t i = 0; i < names->size(); ++i)
				if ((*names)[i] == full) idx = (int)i;
			if (idx < 0) { names->push_back(full); idx = (int)names->size() - 1; }
			gml.m_dtex = idx;
		} else {
			gml.m_dtex = -1;
			gml.m_dr = ml.diffuse[0]; gml.m_dg = ml.diffuse[1]; gml.m_db = ml.diffuse[2];
		}
		gml.m_er = ml.emission[0]; gml.m_eg = ml.emission[1]; gml.m_eb = ml.emission[2];
		gml.m_sr = ml.specular[0]; gml.m_sg = ml.specular[1]; gml.m_sb = ml.specular[2];
		gml.m_illum = ml.illum;
		gml.m_shininess = ml.shininess;
		gml.m_dissolve = ml.dissolve;
		gml.m_refraction_index = ml.ior;
		materials->push_back(gml);
	}
}

} // namespace

extern "C" {

// OBJ -> Triangle[] -> SBVH -> CWBVH, exactly as Instance::Initialize does (Instance.cpp:12-24).
// Returns nullptr on load failure.
void *ref_scene_build(const char *obj_path, int max_spatial_depth, float triangle_sah, float node_sah)
{
	StdoutMute mute;
	RefScene *s = new RefScene;
	if (!s->scene.LoadFromFile(obj_path)) { delete s; return nullptr; }
	InstanceConfig::BVH cfg;
	cfg.m_max_spatial_depth = max_spatial_depth;
	cfg.m_triangle_sah = triangle_sah;
	cfg.m_node_sah = node_sah;
	SBVHBuilder{cfg, &s->sbvh, s->scene}.Run();
	WideBVHBuilder{cfg, &s->wbvh, s->sbvh}.Run();
	init_triangles(s->scene, s->wbvh, &s->woop);
	init_materials(s->scene, &s->materials, &s->dtex_names);
	return s;
}

// OBJ -> Triangle[] only, then load a .bvh cache with WideBVH::LoadFromFile (WideBVH.cpp:25-66).
void *ref_scene_load_bvh(const char *obj_path, const char *bvh_path, int max_spatial_depth,
                         float triangle_sah, float node_sah)
{
	StdoutMute mute;
	RefScene *s = new RefScene;
	if (!s->scene.LoadFromFile(obj_path)) { delete s; return nullptr; }
	InstanceConfig::BVH cfg;
	cfg.m_max_spatial_depth = max_spatial_depth;
	cfg.m_triangle_sah = triangle_sah;
	cfg.m_node_sah = node_sah;
	if (!s->wbvh.LoadFromFile(bvh_path, cfg)) { delete s; return nullptr; }
	init_triangles(s->scene, s->wbvh, &s->woop);
	init_materials(s->scene, &s->materials, &s->dtex_names);
	return s;
}

int ref_scene_save_bvh(void *h, const char *bvh_path, int max_spatial_depth, float triangle_sah, float node_sah)
{
	RefScene *s = (RefScene *)h;
	InstanceConfig::BVH cfg;
	cfg.m_max_spatial_depth = max_spatial_depth;
	cfg.m_triangle_sah = triangle_sah;
	cfg.m_node_sah = node_sah;
	return s->wbvh.SaveToFile(bvh_path, cfg) ? 0 : -1;
}

void ref_scene_destroy(void *h) { delete (RefScene *)h; }

uint32_t ref_scene_n_tris(void *h) { return (uint32_t)((RefScene *)h)->scene.GetTriangles().size(); }
uint32_t ref_scene_n_nodes(void *h) { return (uint32_t)((RefScene *)h)->wbvh.GetNodes().size(); }
uint32_t ref_scene_n_refs(void *h) { return (uint32_t)((RefScene *)h)->wbvh.GetTriIndices().size(); }
uint32_t ref_scene_n_mats(void *h) { return (uint32_t)((RefScene *)h)->materials.size(); }
uint32_t ref_scene_n_sbvh_nodes(void *h) { return (uint32_t)((RefScene *)h)->sbvh.GetNodes().size(); }
const void *ref_scene_tris(void *h) { return ((RefScene *)h)->scene.GetTriangles().data(); }
const void *ref_scene_nodes(void *h) { return ((RefScene *)h)->wbvh.GetNodes().data(); }
const int32_t *ref_scene_tri_indices(void *h) { return ((RefScene *)h)->wbvh.GetTriIndices().data(); }
const float *ref_scene_woop(void *h) { return (const float *)((RefScene *)h)->woop.data(); }
const void *ref_scene_mats(void *h) { return ((RefScene *)h)->materials.data(); }
// SBVHNode[] (32 bytes each: AABB min/max, tri idx, left idx; SBVH.hpp:11-16), for builder parity tests.
const void *ref_scene_sbvh_nodes(void *h) { return ((RefScene *)h)->sbvh.GetNodes().data(); }
void ref_scene_aabb(void *h, float out[6])
{
	const AABB &b = ((RefScene *)h)->scene.GetAABB();
	out[0] = b.m_min.x; out[1] = b.m_min.y; out[2] = b.m_min.z;
	out[3] = b.m_max.x; out[4] = b.m_max.y; out[5] = b.m_max.z;
}
const char *ref_scene_dtex_name(void *h, uint32_t i)
{
	RefScene *s = (RefScene *)h;
	return i < s->dtex_names.size() ? s->dtex_names[i].c_str() : nullptr;
}

// stbi_load(filename, &w, &h, &channels, 3) exactly as OglScene::load_texture calls it (OglScene.cpp:24).
// Copies width*height*3 bytes into out (capacity cap); returns 0, -1 on decode failure, -2 if cap is too small.
int ref_load_image_rgb8(const char *path, int *width, int *height, unsigned char *out, unsigned long long cap)
{
	int w = 0, h = 0, ch = 0;
	unsigned char *data = stbi_load(path, &w, &h, &ch, 3);
	if (!data) return -1;
	*width = w;
	*height = h;
	const unsigned long long need = (unsigned long long)w * h * 3;
	int rc = 0;
	if (need <= cap) memcpy(out, data, need);
	else rc = -2;
	stbi_image_free(data);
	return rc;
}

// SaveEXR(pixels, w, h, 3, save_as_fp16, filename, &err) exactly as OglPathTracer::SaveResult calls it
// (OglPathTracer.cpp:199-212) on a tightly packed RGB float image. Returns tinyexr's status (0 = success).
int ref_save_exr(const float *rgb, int width, int height, int save_as_fp16, const char *filename)
{
	const char *err = nullptr;
	const int rc = SaveEXR(rgb, width, height, 3, save_as_fp16, filename, &err);
	if (err) free((void *)err);
	return rc;
}

// rapidjson's double -> text (Writer::WriteDouble -> internal::dtoa, dep/rapidjson/internal/dtoa.h) for n doubles;
// each result is NUL-terminated at out + 32 * i. Pins the product's own Grisu2 + Prettify.
void ref_dtoa(const double *values, unsigned long long n, char *out)
{
	for (unsigned long long i = 0; i < n; ++i) {
		char *end = rapidjson::internal::dtoa(values[i], out + 32 * i, 324);
		*end = 0;
	}
}

// glm::inverse(mat4) of the reference's vendored glm (dep/glm/detail/func_matrix.inl:294-351),
// column-major float[16] in and out. Pins the oracle's / product's own cofactor restatement.
void ref_mat4_inverse(const float in[16], float out[16])
{
	glm::mat4 m;
	memcpy(&m, in, 64);
	m = glm::inverse(m);
	memcpy(out, &m, 64);
}

// Sobol::Reset(dim) followed by n_calls x Sobol::Next (Sobol.cpp:16-21); writes n_calls*dim floats.
void ref_sobol_sequence(uint32_t dim, uint32_t n_calls, float *out)
{
	static Sobol gen; // 40 KB of state, keep off the stack
	gen.Reset(dim);
	for (uint32_t i = 0; i < n_calls; ++i) gen.Next(out + (size_t)i * dim);
}

// Camera::GetView / GetProjection (Camera.cpp:13-23) and the inverses OglPathTracer::SetCamera
// takes (OglPathTracer.cpp:27-32). Column-major float[16] each.
void ref_camera_matrices(float fov_deg, float yaw_deg, float pitch_deg, int width, int height,
                         float proj[16], float view[16], float inv_proj[16], float inv_view[16])
{
	float aspect = width / (float)height; // Camera.hpp:30
	glm::mat4 v = glm::rotate(glm::identity<glm::mat4>(), glm::radians(-pitch_deg), glm::vec3(1.0f, 0.0f, 0.0f));
	v = glm::rotate(v, glm::radians(-yaw_deg), glm::vec3(0.0f, 1.0f, 0.0f));
	glm::mat4 p = glm::tweakedInfinitePerspective(glm::radians(fov_deg), aspect, 0.01f);
	glm::mat4 ip = glm::inverse(p), iv = glm::inverse(v);
	memcpy(proj, &p, 64); memcpy(view, &v, 64); memcpy(inv_proj, &ip, 64); memcpy(inv_view, &iv, 64);
}

// InstanceConfig::LoadFromFile (InstanceConfig.cpp:10-101) flattened into a POD.
struct RefConfig {
	int32_t width, height;
	int32_t invocation_size, stack_size, max_bounce, subpixel, tmp_lifetime;
	float ray_tmin, clamp, sun[3];
	int32_t max_spatial_depth; float triangle_sah, node_sah;
	float speed, mouse_sensitive, fov, yaw, pitch, position[3];
	char obj_filename[512], bvh_filename[512];
};

static void flatten(const InstanceConfig &c, RefConfig *o)
{
	memset(o, 0, sizeof(*o));
	o->width = c.m_width; o->height = c.m_height;
	o->invocation_size = c.m_pt_cfg.m_invocation_size; o->stack_size = c.m_pt_cfg.m_stack_size;
	o->max_bounce = c.m_pt_cfg.m_max_bounce; o->subpixel = c.m_pt_cfg.m_subpixel;
	o->tmp_lifetime = c.m_pt_cfg.m_tmp_lifetime; o->ray_tmin = c.m_pt_cfg.m_ray_tmin;
	o->clamp = c.m_pt_cfg.m_clamp;
	for (int i = 0; i < 3; ++i) o->sun[i] = c.m_pt_cfg.m_sun[i];
	o->max_spatial_depth = c.m_bvh_cfg.m_max_spatial_depth;
	o->triangle_sah = c.m_bvh_cfg.m_triangle_sah; o->node_sah = c.m_bvh_cfg.m_node_sah;
	o->speed = c.m_cam_cfg.m_speed; o->mouse_sensitive = c.m_cam_cfg.m_mouse_sensitive;
	o->fov = c.m_cam_cfg.m_fov; o->yaw = c.m_cam_cfg.m_yaw; o->pitch = c.m_cam_cfg.m_pitch;
	for (int i = 0; i < 3; ++i) o->position[i] = c.m_cam_cfg.m_position[i];
	snprintf(o->obj_filename, sizeof(o->obj_filename), "%s", c.m_obj_filename.c_str());
	snprintf(o->bvh_filename, sizeof(o->bvh_filename), "%s", c.m_bvh_filename.c_str());
}

int ref_config_load(const char *path, RefConfig *out)
{
	StdoutMute mute;
	InstanceConfig c;
	if (!c.LoadFromFile(path)) return -1;
	flatten(c, out);
	return 0;
}

// InstanceConfig::GetJson (InstanceConfig.cpp:103-192) of a config loaded from `path`.
int ref_config_roundtrip_json(const char *path, char *buf, uint32_t cap)
{
	StdoutMute mute;
	InstanceConfig c;
	if (!c.LoadFromFile(path)) return -1;
	std::string js = c.GetJson();
	if (js.size() + 1 > cap) return -2;
	memcpy(buf, js.c_str(), js.size() + 1);
	return (int)js.size();
}

} // extern "C"
