// Device-resident scene: the CUDA stand-in for OglScene (src/Tracer/OglScene.hpp).
#pragma once
#include "common.h"
#include "layouts.h"

namespace adypt {
// Work counters of the persistent kernels. A slot is zeroed by a memset queued on the launch stream right before the
// launch that uses it, so a slot may only be shared by launches that are ordered on ONE stream (then a single slot is
// enough). Every tracer owns a slot (used by all launches on its stream), each of the three pipeline streams of the
// host-array calls owns one, and launches on caller-provided streams take slots from a ring with an atomic cursor:
// callers that trace on several of their own streams at once get distinct slots as long as fewer than kCounterRing
// launches are in flight. Layout of adypt_scene::d_counters: [0, kCounterRing) ring, +0..4 statistics, +8..10 pipeline.
constexpr unsigned kStatSlots = 5; // nodes visited, triangles tested, rays that hit, deepest stack, rays traced
constexpr unsigned kCounterRing = 256, kCounterStats = kCounterRing, kCounterPipe = kCounterRing + 8, kCounterSlots = kCounterRing + 12;
} // namespace adypt

struct adypt_scene {
	int device = 0;
	int sm_count = 0;
	uint32_t n_nodes = 0, n_refs = 0, n_tris = 0, n_mats = 0;
	uint4 *d_nodes = nullptr;          // n_nodes * 5
	uint4 *d_nodes_wide = nullptr;     // n_nodes * 6: derived 96-byte layout (build_wide_nodes)
	uint4 *d_nodes_wide128 = nullptr;  // experiment (variants 20, 21): n_nodes * 8, built on demand
	float4 *d_woop64 = nullptr;        // experiment (variant 22): n_refs * 4, built on demand
	float4 *d_woop = nullptr;          // n_refs * 3
	int32_t *d_tri_indices = nullptr;  // n_refs
	uint8_t *d_tris = nullptr;         // n_tris * 100
	adypt::Material *d_mats = nullptr; // n_mats
	// derived at upload, like the Woop rows: the Triangle record on its own 128-byte line (8 x float4: floats 0..24 = the
	// 100-byte record, float 25 = shading class) and a word per triangle with the shading class and the material id
	float4 *d_shade = nullptr;         // n_tris * 8
	uint32_t *d_tri_class = nullptr;   // n_tris: shading class << 24 | material id
	uchar4 *d_texels = nullptr;        // all textures back to back, RGBX8
	int4 *d_tex_table = nullptr;       // per texture: (first texel, width, height, 0)
	uint32_t n_textures = 0;           // TEXTURE_COUNT
	unsigned long long *d_counters = nullptr; // work counters of the persistent kernels (layout above)
	std::atomic<unsigned> counter_cursor{0};
	int64_t bad_matid_tri = -1; // first triangle whose material id is outside [0, n_mats): such a scene can be traversed but not shaded
	int ctas_per_sm = 0;       // 0 = occupancy query
	int refill_threshold = 0;  // 0 = default
	int variant = 0;           // code-generation variant of the closest-hit kernel (tuning only)
	int occ_closest = 0, occ_any = 0;
	adypt::DeviceBuffer stage_in, stage_out; // staging for host-pointer batch calls
	cudaStream_t pipe[3] = {nullptr, nullptr, nullptr}; // copy/compute pipeline of host-pointer batch calls
	uint64_t device_bytes = 0;
};

namespace adypt {

// queue a closest-hit (occ == nullptr) or any-hit (occ != nullptr) traversal of n device rays on `stream`;
// d_n (nullable): the actual ray count lives in device memory and n is only its upper bound;
// d_counter (nullable): a work-counter slot owned by `stream` (see above), else one is taken from the scene's ring;
// d_stats (nullable): kStatSlots counters the INSTRUMENTED kernel adds its work to (measurement runs only);
// d_dirs (nullable): the rays come as two arrays, origins+tmin in d_rays and directions in d_dirs (the wavefront's queues), instead of
// the batch ABI's 32-byte records
int launch_trace(adypt_scene *s, const float4 *d_rays, uint64_t n, int32_t *d_tri, float *d_t, float2 *d_uv,
                 uint8_t *d_occ, cudaStream_t stream, const unsigned long long *d_n = nullptr, unsigned long long *d_counter = nullptr,
                 unsigned long long *d_stats = nullptr, const float4 *d_dirs = nullptr);

struct TraceParams;
using TraceKernel = void (*)(const TraceParams);
TraceKernel trace_kernel_for(bool any, bool stats, int variant);

inline void count_launch(unsigned n = 1) { g_launches.fetch_add(n); }

} // namespace adypt
