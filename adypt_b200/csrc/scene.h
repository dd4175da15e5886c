// Device-resident scene: the CUDA stand-in for OglScene (src/Tracer/OglScene.hpp).
#pragma once
#include "common.h"
#include "layouts.h"

struct adypt_scene {
	int device = 0;
	int sm_count = 0;
	uint32_t n_nodes = 0, n_refs = 0, n_tris = 0, n_mats = 0;
	uint4 *d_nodes = nullptr;          // n_nodes * 5
	float4 *d_woop = nullptr;          // n_refs * 3
	int32_t *d_tri_indices = nullptr;  // n_refs
	uint8_t *d_tris = nullptr;         // n_tris * 100
	adypt::Material *d_mats = nullptr; // n_mats
	uchar4 *d_texels = nullptr;        // all textures back to back, RGBX8
	int4 *d_tex_table = nullptr;       // per texture: (first texel, width, height, 0)
	uint32_t n_textures = 0;           // TEXTURE_COUNT
	unsigned long long *d_counters = nullptr; // ring of work counters for the persistent kernels
	unsigned counter_cursor = 0;
	int ctas_per_sm = 0;       // 0 = occupancy query
	int refill_threshold = 0;  // 0 = default
	int variant = 0;           // code-generation variant of the closest-hit kernel (tuning only)
	int occ_closest = 0, occ_any = 0;
	adypt::DeviceBuffer stage_in, stage_out; // staging for host-pointer batch calls
	cudaStream_t pipe[3] = {nullptr, nullptr, nullptr}; // copy/compute pipeline of host-pointer batch calls
	uint64_t device_bytes = 0;
};

namespace adypt {

constexpr unsigned kCounterRing = 256;

// queue a closest-hit (occ == nullptr) or any-hit (occ != nullptr) traversal of n device rays on `stream`;
// d_n (nullable): the actual ray count lives in device memory and n is only its upper bound
int launch_trace(adypt_scene *s, const float4 *d_rays, uint64_t n, int32_t *d_tri, float *d_t, float2 *d_uv,
                 uint8_t *d_occ, cudaStream_t stream, const unsigned long long *d_n = nullptr);

inline void count_launch(unsigned n = 1) { g_launches.fetch_add(n); }

} // namespace adypt
