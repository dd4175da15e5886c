// C++ shim over the C-ABI (adypt_b200.h) with the REFERENCE'S class surface, so that code written against
// Adypt's Instance / InstanceConfig / Scene / WideBVH / Camera / OglScene / OglPathTracer
// (src/Instance.hpp:16-35 and the headers it includes) keeps its shape when the OpenGL path is swapped for
// the CUDA one. Header-only, C++11, no glm: matrices are column-major float[16] (glm::value_ptr(m) on the
// caller's side). Conventions follow the reference: bool / void returns and "[TAG]..." messages on stdout,
// no exceptions, single-threaded use, non-copyable GPU handles.
//
// Define ADYPT_B200_REFERENCE_NAMES before including to also get the aliases OglScene / OglPathTracer.
#ifndef ADYPT_B200_HPP
#define ADYPT_B200_HPP

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include "adypt_b200.h"

namespace adypt_b200 {

// ---------------------------------------------------------------------------------------------------
// src/InstanceConfig.hpp:12-48
struct InstanceConfig {
	struct BVH {
		int32_t m_max_spatial_depth = 48;
		float m_triangle_sah = 0.3f, m_node_sah = 1.0f;
	};
	struct PT { // same memory layout as adypt_pt_config
		int32_t m_invocation_size = 8, m_stack_size = 12, m_max_bounce = 5, m_subpixel = 8, m_tmp_lifetime = 16;
		float m_ray_tmin = 0.0001f, m_clamp = 4.0f;
		float m_sun[3] = {0.f, 0.f, 0.f};
	};
	struct Cam {
		float m_speed = 1.0f, m_mouse_sensitive = 0.3f, m_fov = 45.0f, m_yaw = 0.f, m_pitch = 0.f;
		float m_position[3] = {0.f, 0.f, 0.f};
	};
	int m_width = 1280, m_height = 720;
	BVH m_bvh_cfg;
	PT m_pt_cfg;
	Cam m_cam_cfg;
	std::string m_obj_filename, m_bvh_filename;

	bool LoadFromFile(const char *filename)
	{
		adypt_instance_config c;
		if (adypt_config_load(filename, &c) != ADYPT_OK) {
			printf("%s\n", adypt_last_error());
			return false;
		}
		from_c(c);
		return true;
	}
	bool SaveToFile(const char *filename) const
	{
		adypt_instance_config c;
		to_c(&c);
		return adypt_config_save(&c, filename) == ADYPT_OK;
	}
	std::string GetJson() const
	{
		adypt_instance_config c;
		to_c(&c);
		uint64_t need = 0;
		adypt_config_to_json(&c, nullptr, 0, &need);
		std::string s((size_t)need, '\0');
		adypt_config_to_json(&c, &s[0], need, nullptr);
		s.resize(strlen(s.c_str()));
		return s;
	}
	void SetDefault()
	{
		m_pt_cfg = PT{};
		m_bvh_cfg = BVH{};
		m_cam_cfg = Cam{};
		m_width = 1280;
		m_height = 720;
	}

	void to_c(adypt_instance_config *c) const
	{
		memset(c, 0, sizeof(*c));
		c->width = m_width;
		c->height = m_height;
		memcpy(&c->bvh, &m_bvh_cfg, sizeof(c->bvh));
		memcpy(&c->pt, &m_pt_cfg, sizeof(c->pt));
		memcpy(&c->cam, &m_cam_cfg, sizeof(c->cam));
		snprintf(c->obj_filename, sizeof(c->obj_filename), "%s", m_obj_filename.c_str());
		snprintf(c->bvh_filename, sizeof(c->bvh_filename), "%s", m_bvh_filename.c_str());
	}
	void from_c(const adypt_instance_config &c)
	{
		m_width = c.width;
		m_height = c.height;
		memcpy(&m_bvh_cfg, &c.bvh, sizeof(c.bvh));
		memcpy(&m_pt_cfg, &c.pt, sizeof(c.pt));
		memcpy(&m_cam_cfg, &c.cam, sizeof(c.cam));
		m_obj_filename = c.obj_filename;
		m_bvh_filename = c.bvh_filename;
	}
};
static_assert(sizeof(InstanceConfig::PT) == sizeof(adypt_pt_config), "PT layout");
static_assert(sizeof(InstanceConfig::BVH) == sizeof(adypt_bvh_config), "BVH layout");

// ---------------------------------------------------------------------------------------------------
// src/Util/Scene.hpp:8-27 (+ the WideBVH it is paired with, src/BVH/WideBVH.hpp:28-43)
struct Scene {
	adypt_host_scene *m_handle = nullptr;
	Scene() = default;
	Scene(const Scene &) = delete;
	Scene &operator=(const Scene &) = delete;
	~Scene() { adypt_host_scene_destroy(m_handle); }
	bool LoadFromFile(const char *filename)
	{
		adypt_host_scene_destroy(m_handle);
		m_handle = nullptr;
		if (adypt_host_scene_load_obj(filename, &m_handle) != ADYPT_OK) {
			printf("%s\n", adypt_last_error());
			return false;
		}
		adypt_host_scene_info i;
		adypt_host_scene_get(m_handle, &i);
		printf("[SCENE]Info: %u triangles loaded from %s\n", i.n_tris, filename);
		return true;
	}
};

// WideBVH + the two builders: the arrays live inside the host scene handle
struct WideBVH {
	Scene *m_scene = nullptr;
	explicit WideBVH(Scene *scene) : m_scene(scene) {}
	bool LoadFromFile(const char *filename, const InstanceConfig::BVH &expected)
	{
		return adypt_host_scene_load_bvh(m_scene->m_handle, filename, (const adypt_bvh_config *)&expected) == ADYPT_OK;
	}
	bool SaveToFile(const char *filename, const InstanceConfig::BVH &config)
	{
		return adypt_host_scene_save_bvh(m_scene->m_handle, filename, (const adypt_bvh_config *)&config) == ADYPT_OK;
	}
	// SBVHBuilder{cfg, &sbvh, scene}.Run(); WideBVHBuilder{cfg, &wbvh, sbvh}.Run();  (Instance.cpp:22-24)
	bool Build(const InstanceConfig::BVH &config)
	{
		const auto t0 = std::chrono::steady_clock::now();
		if (adypt_host_scene_build_bvh(m_scene->m_handle, (const adypt_bvh_config *)&config) != ADYPT_OK) {
			printf("[WideBVH]Err: %s\n", adypt_last_error());
			return false;
		}
		adypt_host_scene_info i;
		adypt_host_scene_get(m_scene->m_handle, &i);
		const long ms = (long)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
		printf("[SBVH]building lasted %ld ms\n[SBVH]built with %u nodes\n[WideBVH]built with %u nodes\n", ms, i.n_binary_nodes, i.n_nodes);
		return true;
	}
};

// ---------------------------------------------------------------------------------------------------
// src/Tracer/Camera.hpp:14-39 (the interactive Control() has no headless meaning)
class Camera {
	InstanceConfig::Cam *m_config = nullptr;
	int m_width = 1, m_height = 1;

public:
	void Initialize(InstanceConfig::Cam *cam_cfg, int width, int height)
	{
		m_config = cam_cfg;
		m_width = width;
		m_height = height;
	}
	void GetView(float view[16]) const
	{
		float p[16];
		adypt_camera_matrices(m_config->m_fov, m_config->m_yaw, m_config->m_pitch, m_width, m_height, p, view);
	}
	void GetProjection(float projection[16]) const
	{
		float v[16];
		adypt_camera_matrices(m_config->m_fov, m_config->m_yaw, m_config->m_pitch, m_width, m_height, projection, v);
	}
};

// ---------------------------------------------------------------------------------------------------
// src/Tracer/OglScene.hpp:15-52
struct CudaScene {
	adypt_scene *m_handle = nullptr;
	int m_device = 0;
	CudaScene() = default;
	CudaScene(const CudaScene &) = delete;
	CudaScene &operator=(const CudaScene &) = delete;
	~CudaScene() { adypt_scene_destroy(m_handle); }
	void Initialize(const Scene &scene, const WideBVH &)
	{
		adypt_scene_destroy(m_handle);
		m_handle = nullptr;
		adypt_host_scene_load_textures(scene.m_handle, nullptr, nullptr); // init_materials -> load_texture (OglScene.cpp:51-91)
		if (adypt_host_scene_upload(scene.m_handle, m_device, &m_handle) != ADYPT_OK) printf("[SCENE]Err: %s\n", adypt_last_error());
	}
};

// src/Tracer/OglPathTracer.hpp:17-83
class CudaPathTracer {
public:
	enum ViewerTypes { kDiffuse = 0, kSpecular, kEmissive, kPTRadiance, kNormal, kPosition };

private:
	adypt_tracer *m_handle = nullptr;
	const InstanceConfig::PT *m_config = nullptr; // borrowed, re-read whenever spp == 0 (OglPathTracer.cpp:39-41)
	std::chrono::time_point<std::chrono::high_resolution_clock> m_tracing_start_time;

public:
	ViewerTypes m_viewer_type = kDiffuse;
	uint64_t m_bias_seed = 0; // the reference seeds its bias image from std::random_device

	CudaPathTracer() = default;
	CudaPathTracer(const CudaPathTracer &) = delete;
	CudaPathTracer &operator=(const CudaPathTracer &) = delete;
	~CudaPathTracer() { adypt_tracer_destroy(m_handle); }

	void Initialize(const InstanceConfig::PT *config, const CudaScene &scene, int width, int height)
	{
		m_config = config;
		adypt_tracer_destroy(m_handle);
		m_handle = nullptr;
		if (adypt_tracer_create(scene.m_handle, (const adypt_pt_config *)config, width, height, m_bias_seed, &m_handle) != ADYPT_OK)
			printf("[PT]ERR: %s\n", adypt_last_error());
	}
	void SetCamera(const float projection[16], const float view[16], const float position[3])
	{
		adypt_tracer_set_camera(m_handle, projection, view, position);
	}
	void Trace(bool enable_pt, int samples = 1)
	{
		if (enable_pt) {
			if (GetSPP() == 0) {
				adypt_tracer_set_config(m_handle, (const adypt_pt_config *)m_config); // update_config_args()
				m_viewer_type = kPTRadiance;
				m_tracing_start_time = std::chrono::high_resolution_clock::now();
			}
			if (adypt_tracer_sample(m_handle, samples) != ADYPT_OK) printf("[PT]ERR: %s\n", adypt_last_error());
		} else {
			if (GetSPP()) m_viewer_type = kDiffuse;
			if (adypt_tracer_primary(m_handle, (int32_t)m_viewer_type) != ADYPT_OK) printf("[PT]ERR: %s\n", adypt_last_error());
		}
	}
	void DrawScreen() {} // display only (shaders/screen.glsl); nothing to do headless
	void SaveResult(const char *filename, bool save_as_fp16)
	{
		if (adypt_tracer_save_exr(m_handle, filename, save_as_fp16 ? 1 : 0) != ADYPT_OK) printf("[PT]ERR: %s\n", adypt_last_error());
		else printf("[PT]INFO: Saved image to %s\n", filename);
	}
	int GetSPP() const
	{
		int32_t spp = 0;
		adypt_tracer_spp(m_handle, &spp);
		return spp;
	}
	long GetPTSec() const
	{
		return (long)std::chrono::duration_cast<std::chrono::seconds>(std::chrono::high_resolution_clock::now() - m_tracing_start_time).count();
	}
	void Sync() { adypt_tracer_sync(m_handle); }
	adypt_tracer *Handle() { return m_handle; }
};

// ---------------------------------------------------------------------------------------------------
// src/Instance.hpp:16-35, src/Instance.cpp:10-86 (the GLFWwindow* / Framerate arguments are gone)
class Instance {
	CudaScene m_scene;
	Camera m_camera;
	bool m_valid = false;

public:
	std::string m_filename;
	CudaPathTracer m_path_tracer;
	InstanceConfig m_config;
	bool m_lock_flag = false;
	bool m_enable_pt_flag = false;
	bool m_use_bvh_cache = true;
	int m_device = 0;

	bool Initialize()
	{
		Scene scene;
		if (!scene.LoadFromFile(m_config.m_obj_filename.c_str())) {
			printf("[INSTANCE]Err: Unable to load scene %s\n", m_config.m_obj_filename.c_str());
			return (m_valid = false);
		}
		WideBVH wbvh(&scene);
		if (!m_use_bvh_cache || !wbvh.LoadFromFile(m_config.m_bvh_filename.c_str(), m_config.m_bvh_cfg)) {
			if (!wbvh.Build(m_config.m_bvh_cfg)) return (m_valid = false);
			if (!wbvh.SaveToFile(m_config.m_bvh_filename.c_str(), m_config.m_bvh_cfg)) {
				printf("[INSTANCE]Err: Unable to load bvh %s\n", m_config.m_bvh_filename.c_str());
				return (m_valid = false);
			}
		}
		m_scene.m_device = m_device;
		m_scene.Initialize(scene, wbvh);
		if (!m_scene.m_handle) return (m_valid = false);
		m_path_tracer.Initialize(&m_config.m_pt_cfg, m_scene, m_config.m_width, m_config.m_height);
		if (!m_path_tracer.Handle()) return (m_valid = false);
		m_camera.Initialize(&m_config.m_cam_cfg, m_config.m_width, m_config.m_height);
		printf("[INSTANCE]Info: Initialized from %s\n", m_filename.c_str());
		return (m_valid = true);
	}
	void Update(int samples = 1)
	{
		if (!m_lock_flag) {
			if (!m_enable_pt_flag) {
				float proj[16], view[16];
				m_camera.GetProjection(proj);
				m_camera.GetView(view);
				m_path_tracer.SetCamera(proj, view, m_config.m_cam_cfg.m_position);
			}
			m_path_tracer.Trace(m_enable_pt_flag, samples);
		}
		m_path_tracer.DrawScreen();
	}
	bool InitializeFromFile(const char *filename)
	{
		m_filename = filename;
		if (!m_config.LoadFromFile(filename)) {
			printf("[INSTANCE]Err: Invalid instance %s\n", filename);
			return (m_valid = false);
		}
		printf("[INSTANCE]Info: Instance loaded from %s\n", filename);
		return (m_valid = Initialize());
	}
	bool SaveToFile()
	{
		if (!m_valid) return false;
		if (m_config.SaveToFile(m_filename.c_str())) {
			printf("[INSTANCE]Info: %s saved\n", m_filename.c_str());
			return true;
		}
		return false;
	}
	bool m_autosave = true; // the reference's destructor rewrites the .config (Instance.cpp:83-86)
	~Instance()
	{
		if (m_autosave) SaveToFile();
	}
};

#ifdef ADYPT_B200_REFERENCE_NAMES
using OglScene = CudaScene;
using OglPathTracer = CudaPathTracer;
#endif

} // namespace adypt_b200

#endif // ADYPT_B200_HPP
