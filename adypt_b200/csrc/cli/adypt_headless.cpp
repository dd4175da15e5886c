// Headless replacement for the reference's GLFW/ImGui viewer (src/main.cpp, src/Application.*): loads an Adypt
// .config instance, renders N samples per pixel (or one AOV frame) on the GPU and writes an OpenEXR file.
//
//   adypt_headless scene.config [--spp N] [--out result.exr] [--fp16] [--seed S] [--device D]
//                  [--viewer diffuse|specular|emissive|normal|position] [--per-frame] [--no-bvh-cache] [--keep-config]
//                  [--preview-every N]   (progressive preview: also writes <out>.<spp>spp.exr every N samples)
//                  [--sun-visibility]    (connect stage: the any-hit sun test of pathtracer.glsl:132)
//                  [--gpus N]            (sample-sharded over devices 0..N-1 of this box, one NCCL reduce)
//
// Same file formats as the reference (.config JSON, OBJ/MTL, .bvh cache); --spp/--seed/--out are new (the
// reference renders until the user stops it and seeds from std::random_device).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include "adypt_b200.hpp"

using namespace adypt_b200;

static int usage()
{
	fprintf(stderr, "usage: adypt_headless <instance.config> [--spp N] [--out file.exr] [--fp16] [--seed S] [--device D]\n"
	                "                      [--viewer diffuse|specular|emissive|normal|position] [--per-frame] [--no-bvh-cache] [--keep-config]\n"
	                "                      [--gpus N] [--preview-every K] [--sun-visibility] [--russian-roulette FIRST_BOUNCE]\n");
	return 2;
}

// --gpus N: the same load path as Instance::Initialize (Instance.cpp:10-42), then a render group instead of one tracer
static int render_on_group(const char *config, int spp, int gpus, const char *out, bool fp16, unsigned long long seed, bool cache, bool sun_visibility,
                           int rr_start)
{
	InstanceConfig cfg;
	if (!cfg.LoadFromFile(config)) {
		printf("[INSTANCE]Err: Invalid instance %s\n", config);
		return 1;
	}
	Scene scene;
	if (!scene.LoadFromFile(cfg.m_obj_filename.c_str())) {
		printf("[INSTANCE]Err: Unable to load scene %s\n", cfg.m_obj_filename.c_str());
		return 1;
	}
	WideBVH wbvh(&scene);
	if (!cache || !wbvh.LoadFromFile(cfg.m_bvh_filename.c_str(), cfg.m_bvh_cfg)) {
		if (!wbvh.Build(cfg.m_bvh_cfg)) return 1;
		wbvh.SaveToFile(cfg.m_bvh_filename.c_str(), cfg.m_bvh_cfg);
	}
	adypt_host_scene_load_textures(scene.m_handle, nullptr, nullptr);
	int32_t devices[64];
	for (int i = 0; i < gpus && i < 64; ++i) devices[i] = i;
	adypt_group *group = nullptr;
	if (adypt_group_create(scene.m_handle, (const adypt_pt_config *)&cfg.m_pt_cfg, cfg.m_width, cfg.m_height, seed, devices, (uint32_t)gpus, &group) != ADYPT_OK) {
		printf("[PT]ERR: %s\n", adypt_last_error());
		return 1;
	}
	Camera camera;
	camera.Initialize(&cfg.m_cam_cfg, cfg.m_width, cfg.m_height);
	float proj[16], view[16];
	camera.GetProjection(proj);
	camera.GetView(view);
	adypt_group_set_camera(group, proj, view, cfg.m_cam_cfg.m_position);
	if (sun_visibility) {
		const float dir[3] = {0.6f, 1.0f, 0.2f};
		adypt_group_set_sun_visibility(group, 1, dir);
	}
	if (rr_start >= 0 && adypt_group_set_russian_roulette(group, rr_start) != ADYPT_OK) printf("[PT]Err: %s\n", adypt_last_error());
	adypt_group_render(group, 16); // warm-up: allocations, NCCL channels
	const auto t0 = std::chrono::steady_clock::now();
	int rc = adypt_group_render(group, spp);
	const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (rc != ADYPT_OK) {
		printf("[PT]ERR: %s\n", adypt_last_error());
		return 1;
	}
	if (adypt_group_save_exr(group, out, fp16 ? 1 : 0) != ADYPT_OK) printf("[PT]ERR: %s\n", adypt_last_error());
	else printf("[PT]INFO: Saved image to %s\n", out);
	printf("{\"spp\": %d, \"width\": %d, \"height\": %d, \"gpus\": %d, \"seconds\": %.6f, \"samples_per_s\": %.1f, \"out\": \"%s\"}\n", spp, cfg.m_width,
	       cfg.m_height, gpus, seconds, (double)cfg.m_width * cfg.m_height * spp / seconds, out);
	adypt_group_destroy(group);
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 2) return usage();
	const char *config = nullptr, *out = "result.exr", *viewer = nullptr;
	int spp = 64, device = 0, preview_every = 0, gpus = 1, rr_start = -1;
	bool fp16 = false, per_frame = false, cache = true, keep = false, sun_visibility = false;
	unsigned long long seed = 0;
	for (int i = 1; i < argc; ++i) {
		const char *a = argv[i];
		auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
		if (!strcmp(a, "--spp")) spp = atoi(next());
		else if (!strcmp(a, "--out")) out = next();
		else if (!strcmp(a, "--fp16")) fp16 = true;
		else if (!strcmp(a, "--seed")) seed = strtoull(next(), nullptr, 10);
		else if (!strcmp(a, "--device")) device = atoi(next());
		else if (!strcmp(a, "--gpus")) gpus = atoi(next()); // sample-sharded over devices 0..N-1, one NCCL reduce
		else if (!strcmp(a, "--viewer")) viewer = next();
		else if (!strcmp(a, "--preview-every")) preview_every = atoi(next());
		else if (!strcmp(a, "--sun-visibility")) sun_visibility = true; // the shader's commented-out any-hit sun test (pathtracer.glsl:132)
		else if (!strcmp(a, "--russian-roulette")) rr_start = atoi(next()); // opt-in extension: roulette from bounce N on
		else if (!strcmp(a, "--per-frame")) per_frame = true; // one Trace(true) per sample, like the viewer's main loop
		else if (!strcmp(a, "--no-bvh-cache")) cache = false;
		else if (!strcmp(a, "--keep-config")) keep = true;    // do not rewrite the .config on exit
		else if (a[0] == '-') return usage();
		else config = a;
	}
	if (!config || spp < 0) return usage();
	{ // --gpus / --device are checked against the box before the (expensive) OBJ load and BVH build
		int n_devices = 0;
		if (adypt_device_count(&n_devices) != ADYPT_OK || n_devices < 1) {
			fprintf(stderr, "adypt_headless: no CUDA device (%s)\n", adypt_last_error());
			return 1;
		}
		const int most = n_devices < 64 ? n_devices : 64;
		if (gpus < 1 || gpus > most) {
			fprintf(stderr, "adypt_headless: --gpus %d is outside 1..%d (devices on this box, at most 64)\n", gpus, most);
			return 2;
		}
		if (device < 0 || device >= n_devices) {
			fprintf(stderr, "adypt_headless: --device %d is outside 0..%d\n", device, n_devices - 1);
			return 2;
		}
	}

	if (gpus > 1 && !viewer) return render_on_group(config, spp, gpus, out, fp16, seed, cache, sun_visibility, rr_start);

	Instance instance;
	instance.m_device = device;
	instance.m_use_bvh_cache = cache;
	instance.m_autosave = !keep;
	instance.m_path_tracer.m_bias_seed = seed;
	if (!instance.InitializeFromFile(config)) return 1;

	if (sun_visibility) {
		const float dir[3] = {0.6f, 1.0f, 0.2f};
		adypt_tracer_set_sun_visibility(instance.m_path_tracer.Handle(), 1, dir);
	}
	if (rr_start >= 0 && adypt_tracer_set_russian_roulette(instance.m_path_tracer.Handle(), rr_start) != ADYPT_OK)
		printf("[PT]Err: %s\n", adypt_last_error());
	instance.m_enable_pt_flag = false;
	if (viewer) {
		static const char *names[] = {"diffuse", "specular", "emissive", "radiance", "normal", "position"};
		int t = -1;
		for (int k = 0; k < 6; ++k)
			if (!strcmp(viewer, names[k])) t = k;
		if (t < 0 || t == 3) return usage();
		instance.m_path_tracer.m_viewer_type = (CudaPathTracer::ViewerTypes)t;
	}
	instance.Update(); // sets the camera and renders the AOV frame, exactly like the viewer before "Start"
	double seconds = 0.0;
	if (!viewer) {
		instance.m_enable_pt_flag = true;
		const auto t0 = std::chrono::steady_clock::now();
		if (preview_every > 0) {
			for (int done = 0; done < spp;) {
				const int n = spp - done < preview_every ? spp - done : preview_every;
				instance.Update(n);
				done += n;
				if (done < spp) {
					char name[1200];
					snprintf(name, sizeof(name), "%s.%dspp.exr", out, done);
					instance.m_path_tracer.SaveResult(name, fp16); // like the showcase renders' "<scene>-<spp>spp.exr"
				}
			}
		} else if (per_frame)
			for (int s = 0; s < spp; ++s) instance.Update(1);
		else
			instance.Update(spp);
		instance.m_path_tracer.Sync();
		seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	}
	instance.m_path_tracer.SaveResult(out, fp16);
	uint64_t segments = 0, launches = 0;
	adypt_tracer_stats(instance.m_path_tracer.Handle(), &segments, &launches);
	const double samples = (double)instance.m_config.m_width * instance.m_config.m_height * instance.m_path_tracer.GetSPP();
	printf("{\"spp\": %d, \"width\": %d, \"height\": %d, \"seconds\": %.6f, \"samples_per_s\": %.1f, \"segments\": %llu, \"launches\": %llu, \"out\": \"%s\"}\n",
	       instance.m_path_tracer.GetSPP(), instance.m_config.m_width, instance.m_config.m_height, seconds,
	       seconds > 0 ? samples / seconds : 0.0, (unsigned long long)segments, (unsigned long long)launches, out);
	return 0;
}
