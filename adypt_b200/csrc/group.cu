// Multi-GPU render group for ONE process driving several B200s of a box (the C++ counterpart of
// tools/render_sharded.py, which does the same with one process per GPU under torch.distributed).
// Not in the reference (single GPU). Partition = whole tmpLifetime blocks of sample indices, round-robin over
// devices; scene replicated per device; every device adds its blocks into its tracer's SUM
// accumulator; ONE ncclReduce(sum, fp32, W*H*4) brings them to the first device, which divides by the sample
// count. NCCL is resolved with dlopen at first use so the library itself carries no NCCL dependency (and never
// clashes with the copy PyTorch bundles); a group of one device needs no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>
#include <string>
#include <vector>
#include "scene.h"

namespace {

struct NcclApi {
	void *lib = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	bool load(std::string *err)
	{
		if (lib) return true;
		for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
			lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
			if (lib) break;
		}
		if (!lib) { *err = "libnccl.so.2 not found (needed for a multi-GPU group)"; return false; }
#define SYM(field, sym) field = (decltype(field))dlsym(lib, sym); if (!field) { *err = std::string("NCCL symbol missing: ") + sym; return false; }
		SYM(CommInitAll, "ncclCommInitAll") SYM(CommDestroy, "ncclCommDestroy") SYM(Reduce, "ncclReduce")
		SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
		return true;
	}
};
NcclApi g_nccl;

} // namespace

struct adypt_group {
	std::vector<int> devices;
	std::vector<adypt_scene *> scenes;
	std::vector<adypt_tracer *> tracers;
	std::vector<ncclComm_t> comms;
	int tmp_lifetime = 16;
	int32_t spp = 0;
};

using namespace adypt;

#define ADYPT_NCCL(expr)                                                                                          \
	do {                                                                                                          \
		ncclResult_t _r = (expr);                                                                                 \
		if (_r != ncclSuccess) return fail(ADYPT_ECUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));  \
	} while (0)

extern "C" {

int adypt_group_destroy(adypt_group *g)
{
	return guarded([&]() -> int {
	if (!g) return ADYPT_OK;
	for (ncclComm_t c : g->comms)
		if (c) g_nccl.CommDestroy(c);
	for (adypt_tracer *t : g->tracers) adypt_tracer_destroy(t);
	for (adypt_scene *s : g->scenes) adypt_scene_destroy(s);
	delete g;
	return ADYPT_OK;
	});
}

int adypt_group_create(adypt_host_scene *scene, const adypt_pt_config *config, int32_t width, int32_t height, uint64_t bias_seed,
                       const int32_t *devices, uint32_t n_devices, adypt_group **out)
{
	return guarded([&]() -> int {
	if (!scene || !config || !devices || !out || n_devices == 0) return fail(ADYPT_EINVAL, "NULL argument or empty device list");
	*out = nullptr;
	adypt_group *g = new adypt_group;
	g->tmp_lifetime = config->tmp_lifetime;
	for (uint32_t i = 0; i < n_devices; ++i) {
		adypt_scene *s = nullptr;
		adypt_tracer *t = nullptr;
		int rc = adypt_host_scene_upload(scene, devices[i], &s);
		if (rc == ADYPT_OK) rc = adypt_tracer_create(s, config, width, height, bias_seed, &t); // same seed => same bias image everywhere
		if (rc != ADYPT_OK) {
			adypt_scene_destroy(s);
			adypt_group_destroy(g);
			return rc;
		}
		g->devices.push_back(devices[i]);
		g->scenes.push_back(s);
		g->tracers.push_back(t);
	}
	if (n_devices > 1) {
		std::string err;
		if (!g_nccl.load(&err)) {
			adypt_group_destroy(g);
			return fail(ADYPT_ENODEV, err);
		}
		g->comms.assign(n_devices, nullptr);
		ncclResult_t r = g_nccl.CommInitAll(g->comms.data(), (int)n_devices, g->devices.data());
		if (r != ncclSuccess) {
			g->comms.clear();
			adypt_group_destroy(g);
			return fail(ADYPT_ECUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r));
		}
	}
	*out = g;
	return ADYPT_OK;
	});
}

int adypt_group_set_camera(adypt_group *g, const float projection[16], const float view[16], const float position[3])
{
	return guarded([&]() -> int {
	if (!g) return fail(ADYPT_EINVAL, "group is NULL");
	for (adypt_tracer *t : g->tracers) ADYPT_TRY(adypt_tracer_set_camera(t, projection, view, position));
	return ADYPT_OK;
	});
}

int adypt_group_set_russian_roulette(adypt_group *g, int32_t start_bounce)
{
	return guarded([&]() -> int {
	if (!g) return fail(ADYPT_EINVAL, "group is NULL");
	for (adypt_tracer *t : g->tracers) ADYPT_TRY(adypt_tracer_set_russian_roulette(t, start_bounce));
	return ADYPT_OK;
	});
}

int adypt_group_set_sun_visibility(adypt_group *g, int32_t enabled, const float direction[3])
{
	return guarded([&]() -> int {
	if (!g) return fail(ADYPT_EINVAL, "group is NULL");
	for (adypt_tracer *t : g->tracers) ADYPT_TRY(adypt_tracer_set_sun_visibility(t, enabled, direction));
	return ADYPT_OK;
	});
}

// Renders samples [0, total_spp) across the group and leaves the resolved image on the first device.
int adypt_group_render(adypt_group *g, int32_t total_spp)
{
	return guarded([&]() -> int {
	if (!g || total_spp <= 0) return fail(ADYPT_EINVAL, "bad argument");
	const int n = (int)g->tracers.size(), L = g->tmp_lifetime;
	for (adypt_tracer *t : g->tracers) ADYPT_TRY(adypt_tracer_clear_sum(t));
	// block k -> device k mod n; enqueue round by round so every device starts working at once
	const int n_blocks = (total_spp + L - 1) / L;
	for (int k = 0; k < n_blocks; ++k) {
		const int first = k * L, cnt = (first + L <= total_spp) ? L : total_spp - first;
		const int rc = adypt_tracer_accumulate(g->tracers[(size_t)(k % n)], first, cnt);
		if (rc != ADYPT_OK) {
			const std::string why = adypt_last_error();
			for (adypt_tracer *t : g->tracers) adypt_tracer_sync(t); // the other devices' queued blocks finish before we report
			return fail(rc, why);
		}
	}
	if (n > 1) {
		// everything that can fail on our side is fetched BEFORE the NCCL group opens; inside it an early return would
		// leave the thread's group open and every later NCCL call of the process queued for ever
		std::vector<float *> bufs((size_t)n, nullptr);
		std::vector<cudaStream_t> streams((size_t)n, nullptr);
		uint64_t count = 0;
		for (int i = 0; i < n; ++i) {
			ADYPT_TRY(adypt_tracer_sum_buffer(g->tracers[(size_t)i], &bufs[(size_t)i], &count));
			ADYPT_TRY(adypt_tracer_stream(g->tracers[(size_t)i], (void **)&streams[(size_t)i]));
		}
		ADYPT_NCCL(g_nccl.GroupStart());
		ncclResult_t first_error = ncclSuccess;
		for (int i = 0; i < n && first_error == ncclSuccess; ++i)
			first_error = g_nccl.Reduce(bufs[(size_t)i], bufs[(size_t)i], (size_t)count, ncclFloat, ncclSum, 0, g->comms[(size_t)i], streams[(size_t)i]);
		const ncclResult_t end = g_nccl.GroupEnd(); // always closed
		if (first_error != ncclSuccess || end != ncclSuccess) {
			for (adypt_tracer *t : g->tracers) adypt_tracer_sync(t); // nobody keeps running with a half-reduced buffer
			return fail(ADYPT_ECUDA, std::string("ncclReduce: ") + g_nccl.GetErrorString(first_error != ncclSuccess ? first_error : end));
		}
	}
	ADYPT_TRY(adypt_tracer_resolve_sum(g->tracers[0]));
	for (adypt_tracer *t : g->tracers) ADYPT_TRY(adypt_tracer_sync(t));
	g->spp = total_spp;
	return ADYPT_OK;
	});
}

int adypt_group_read(adypt_group *g, float *out, int32_t channels)
{
	return guarded([&]() -> int {
	if (!g) return fail(ADYPT_EINVAL, "group is NULL");
	return adypt_tracer_read(g->tracers[0], out, channels);
	});
}

int adypt_group_save_exr(adypt_group *g, const char *filename, int32_t save_as_fp16)
{
	return guarded([&]() -> int {
	if (!g) return fail(ADYPT_EINVAL, "group is NULL");
	return adypt_tracer_save_exr(g->tracers[0], filename, save_as_fp16);
	});
}

} // extern "C"
