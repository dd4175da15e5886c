// Scene upload, GPU Woop construction and the batch traversal entry points of the C-ABI.
#include "scene.h"
#include "traverse.cuh"
#include "traverse_pool.cuh"
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace adypt {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string &msg) { t_error = msg; }
int fail(int code, const std::string &msg)
{
	t_error = msg;
	return code;
}

// ------------------------------------------------------------------------------------------------
// OglScene::init_triangles (OglScene.cpp:93-116) on the GPU, one thread per leaf reference. The 4x4
// inverse follows glm::inverse's cofactor expansion (dep/glm/detail/func_matrix.inl:294-351) operation by
// operation; the library is built with -fmad=false so nothing is contracted and the rows are
// bit-identical to the host build of the reference.
__device__ void mat4_inverse_glm(const float m[4][4], float out[4][4])
{
	const float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
	const float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
	const float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
	const float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
	const float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
	const float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
	const float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
	const float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
	const float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
	const float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
	const float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
	const float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
	const float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
	const float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
	const float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
	const float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
	const float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
	const float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
	const float Fac0[4] = {Coef00, Coef00, Coef02, Coef03}, Fac1[4] = {Coef04, Coef04, Coef06, Coef07};
	const float Fac2[4] = {Coef08, Coef08, Coef10, Coef11}, Fac3[4] = {Coef12, Coef12, Coef14, Coef15};
	const float Fac4[4] = {Coef16, Coef16, Coef18, Coef19}, Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
	const float V0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, V1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
	const float V2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, V3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
	const float SignA[4] = {+1.f, -1.f, +1.f, -1.f}, SignB[4] = {-1.f, +1.f, -1.f, +1.f};
	float inv[4][4];
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const float Inv0 = V1[i] * Fac0[i] - V2[i] * Fac1[i] + V3[i] * Fac2[i];
		const float Inv1 = V0[i] * Fac0[i] - V2[i] * Fac3[i] + V3[i] * Fac4[i];
		const float Inv2 = V0[i] * Fac1[i] - V1[i] * Fac3[i] + V3[i] * Fac5[i];
		const float Inv3 = V0[i] * Fac2[i] - V1[i] * Fac4[i] + V2[i] * Fac5[i];
		inv[0][i] = Inv0 * SignA[i];
		inv[1][i] = Inv1 * SignB[i];
		inv[2][i] = Inv2 * SignA[i];
		inv[3][i] = Inv3 * SignB[i];
	}
	const float d0 = m[0][0] * inv[0][0], d1 = m[0][1] * inv[1][0], d2 = m[0][2] * inv[2][0], d3 = m[0][3] * inv[3][0];
	const float Dot1 = (d0 + d1) + (d2 + d3);
	const float OneOverDeterminant = __fdiv_rn(1.0f, Dot1);
#pragma unroll
	for (int c = 0; c < 4; ++c)
#pragma unroll
		for (int r = 0; r < 4; ++r) out[c][r] = inv[c][r] * OneOverDeterminant;
}

__global__ void build_woop_kernel(const uint8_t *__restrict__ tris, const int32_t *__restrict__ tri_indices, uint32_t n_refs,
                                  float4 *__restrict__ woop)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_refs) return;
	const float *p = (const float *)(tris + (size_t)tri_indices[i] * 100u); // 100-byte stride: 4-byte aligned
	const float v0[3] = {p[0], p[1], p[2]}, v1[3] = {p[3], p[4], p[5]}, v2[3] = {p[6], p[7], p[8]};
	const float e0[3] = {v0[0] - v2[0], v0[1] - v2[1], v0[2] - v2[2]};
	const float e1[3] = {v1[0] - v2[0], v1[1] - v2[1], v1[2] - v2[2]};
	// glm::cross: (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
	const float cr[3] = {e0[1] * e1[2] - e1[1] * e0[2], e0[2] * e1[0] - e1[2] * e0[0], e0[0] * e1[1] - e1[0] * e0[1]};
	// the mat4 constructor of OglScene.cpp:104-109 takes column-major scalars
	const float m[4][4] = {{e0[0], e1[0], cr[0], v2[0]}, {e0[1], e1[1], cr[1], v2[1]}, {e0[2], e1[2], cr[2], v2[2]}, {0.f, 0.f, 0.f, 1.f}};
	float inv[4][4];
	mat4_inverse_glm(m, inv);
	woop[3 * (size_t)i + 0] = make_float4(inv[2][0], inv[2][1], inv[2][2], -inv[2][3]);
	woop[3 * (size_t)i + 1] = make_float4(inv[0][0], inv[0][1], inv[0][2], inv[0][3]);
	woop[3 * (size_t)i + 2] = make_float4(inv[1][0], inv[1][1], inv[1][2], inv[1][3]);
}

// Wide nodes (traverse.cuh, MODE 2): the reference's 80-byte node followed by its three plane scales as floats, at a 96-byte stride
__global__ void build_wide_nodes(const uint4 *__restrict__ nodes, uint32_t n_nodes, uint4 *__restrict__ wide)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_nodes) return;
	uint4 *o = wide + (size_t)i * 6u;
#pragma unroll
	for (int k = 0; k < 5; ++k) o[k] = nodes[(size_t)i * 5u + k];
	const uint32_t w = nodes[(size_t)i * 5u].w;
	o[5] = make_uint4((w & 0xffu) << 23, ((w >> 8) & 0xffu) << 23, ((w >> 16) & 0xffu) << 23, 0u);
}

// EXPERIMENT (traverse.cuh, MODE 3): 128-byte nodes with the children's hit-mask words
__global__ void build_wide_nodes128(const uint4 *__restrict__ nodes, uint32_t n_nodes, uint4 *__restrict__ wide)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_nodes) return;
	const uint4 n0 = nodes[(size_t)i * 5u], n1 = nodes[(size_t)i * 5u + 1], n2 = nodes[(size_t)i * 5u + 2], n3 = nodes[(size_t)i * 5u + 3], n4 = nodes[(size_t)i * 5u + 4];
	uint32_t c[8];
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const uint32_t b = ((k < 4 ? n1.z : n1.w) >> (8 * (k & 3))) & 0xffu;
		c[k] = (b >> 5) << (b & 31u);
	}
	uint4 *o = wide + (size_t)i * 8u;
	o[0] = make_uint4(n0.x, n0.y, n0.z, n0.w >> 24);
	o[1] = make_uint4((n0.w & 0xffu) << 23, ((n0.w >> 8) & 0xffu) << 23, ((n0.w >> 16) & 0xffu) << 23, n1.x);
	o[2] = n2;
	o[3] = n3;
	o[4] = n4;
	o[5] = make_uint4(c[0], c[1], c[2], c[3]);
	o[6] = make_uint4(c[4], c[5], c[6], c[7]);
	o[7] = make_uint4(n1.y, 0u, 0u, 0u);
}

// EXPERIMENT (traverse.cuh, MODE 5): Woop rows at a 64-byte stride
__global__ void build_woop64(const float4 *__restrict__ woop, uint32_t n_refs, float4 *__restrict__ out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_refs) return;
	out[(size_t)i * 4u] = woop[(size_t)i * 3u];
	out[(size_t)i * 4u + 1] = woop[(size_t)i * 3u + 1];
	out[(size_t)i * 4u + 2] = woop[(size_t)i * 3u + 2];
	out[(size_t)i * 4u + 3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Shading records: the wavefront's shading stage gathers one Triangle per segment. The reference's record is
// 100 bytes at a 100-byte stride (Shape.hpp:70-88) -- 25 scalar loads over 4 or 5 sectors; here it is copied, unchanged, to
// the start of a 128-byte line (7 vector loads, exactly one L2 line). The shading branch its material selects
// (pathtracer.glsl:144-201) is also kept in a word per triangle (class << 24 | material id), so that the stage can regroup a block's
// segments by branch -- and request the material -- without touching the record. 1 diffuse (illum 1, and illum 2 with shininess*0.01 <= 0.3), 2 glossy,
// 3 mirror (illum 3-5), 4 dielectric (6, 7), 5 everything else (passes straight through).
__global__ void build_shade_records(const uint8_t *__restrict__ tris, const Material *__restrict__ mats, uint32_t n_tris, uint32_t n_mats,
                                    float4 *__restrict__ shade, uint32_t *__restrict__ tri_class)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_tris) return;
	const float *p = (const float *)(tris + (size_t)i * 100u);
	float f[32];
#pragma unroll
	for (int k = 0; k < 25; ++k) f[k] = p[k];
	const int32_t matid = __float_as_int(f[24]);
	uint32_t cls = 5u;
	if (matid >= 0 && (uint32_t)matid < n_mats) {
		const int32_t illum = mats[matid].illum;
		cls = illum == 1 ? 1u : illum == 2 ? (mats[matid].shininess * 0.01f > 0.3f ? 2u : 1u) : (illum >= 3 && illum <= 5) ? 3u : (illum == 6 || illum == 7) ? 4u : 5u;
	}
	f[25] = __uint_as_float(cls);
#pragma unroll
	for (int k = 26; k < 32; ++k) f[k] = 0.0f;
#pragma unroll
	for (int k = 0; k < 8; ++k) shade[(size_t)i * 8u + k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
	// class in the top byte, material id below it: the shading stage learns both one round ahead of the record itself, so the
	// material's 64 bytes are requested together with the record instead of after it
	tri_class[i] = (cls << 24) | ((uint32_t)matid < 0x00ffffffu ? (uint32_t)matid : 0x00ffffffu); // 0xffffff: look in the record
}

// ------------------------------------------------------------------------------------------------
// The kernel a launch uses. Variants are code-generation variants of the same algorithm (identical results; tuning and
// A/B measurements only); 0 = tuned default: packed slab / Woop evaluations on 96-byte nodes (MODE 2), 3 conversion planes on the I2F
// pipe, 8 CTAs/SM, triangle batch 2 with both fetches up front, staged ray set-up. 19 = the product kernel of round 1 and most of round 2
// (scalar evaluations, the reference's 80-byte nodes). `stats` selects the instrumented build (work counters; slower).
TraceKernel trace_kernel_for(bool any, bool stats, int variant)
{
	if (stats) return any ? trace_kernel<true, true> : trace_kernel<false, true>;
	if (any) switch (variant) {
	case 1: return trace_kernel<true, false, 2, 8, 0>;
	case 2: return trace_kernel<true, false, 2, 8, 2>;
	case 5: return trace_kernel<true, false, 4, 8, 2>;
	case 8: return trace_kernel<true, false, 4, 8, 12, false>;
	case 9: return trace_kernel<true, false, 4, 8, 12, true, true>;
	case 13: return trace_kernel<true, false, 4, 8, 12, true, false, 1>;
	case 14: return trace_kernel<true, false, 3, 8, 12, true, false, 1>;
	case 15: return trace_kernel<true, false, 2, 8, 12, true, false, 1>;
	case 16: return trace_kernel<true, false, 4, 8, 12, true, false, 2>;
	case 17: return trace_kernel<true, false, 3, 8, 12, true, false, 2>;
	case 18: return trace_kernel<true, false, 2, 8, 12, true, false, 2>;
	case 19: return trace_kernel<true>;
	case 20: return trace_kernel<true, false, 3, 8, 12, true, false, 3>;
	case 21: return trace_kernel<true, false, 4, 8, 12, true, false, 3>;
	case 22: return trace_kernel<true, false, 3, 8, 12, true, false, 5>;
	default: return trace_kernel<true, false, 3, 8, 12, true, false, 2>;
	}
	switch (variant) {
	case 1: return trace_kernel<false, false, 2, 8, 0>; // unbounded triangle loop
	case 2: return trace_kernel<false, false, 2, 8, 2>;
	case 3: return trace_kernel<false, false, 2, 8, 12>;
	case 4: return trace_kernel<false, false, 3, 8, 12>;
	case 5: return trace_kernel<false, false, 4, 8, 2>;
	case 6: return trace_kernel<false, false, 5, 8, 12>;
	case 7: return trace_kernel<false, false, 4, 8, 1>;
	case 8: return trace_kernel<false, false, 4, 8, 12, false>; // ray set-up at refill time
	case 9: return trace_kernel<false, false, 4, 8, 12, true, true>; // hit-mask contributions from a shared-memory table
	case 10: return trace_kernel<false, false, 3, 8, 12, true, true>;
	case 11: return trace_kernel<false, false, 5, 8, 12, true, true>;
	case 13: return trace_kernel<false, false, 4, 8, 12, true, false, 1>; // slab evaluations as packed FFMA2 / FADD2
	case 14: return trace_kernel<false, false, 3, 8, 12, true, false, 1>;
	case 15: return trace_kernel<false, false, 2, 8, 12, true, false, 1>;
	case 16: return trace_kernel<false, false, 4, 8, 12, true, false, 2>; // packed + 96-byte nodes fetched with three 256-bit loads
	case 17: return trace_kernel<false, false, 3, 8, 12, true, false, 2>;
	case 18: return trace_kernel<false, false, 2, 8, 12, true, false, 2>;
	case 19: return trace_kernel<false>; // scalar evaluations, 80-byte nodes
	case 20: return trace_kernel<false, false, 3, 8, 12, true, false, 3>; // experiment: 128-byte nodes with hit-mask words
	case 21: return trace_kernel<false, false, 4, 8, 12, true, false, 3>;
	case 22: return trace_kernel<false, false, 3, 8, 12, true, false, 5>; // experiment: the product kernel with 64-byte Woop rows, two 256-bit loads per test
	default: return trace_kernel<false, false, 3, 8, 12, true, false, 2>;
	}
}

// ------------------------------------------------------------------------------------------------
int launch_trace(adypt_scene *s, const float4 *d_rays, uint64_t n, int32_t *d_tri, float *d_t, float2 *d_uv, uint8_t *d_occ,
                 cudaStream_t stream, const unsigned long long *d_n, unsigned long long *d_counter, unsigned long long *d_stats, const float4 *d_dirs)
{
	if (n == 0) return ADYPT_OK;
	const bool any = d_occ != nullptr;
	TraceParams p;
	p.nodes = s->d_nodes;
	p.nodes_wide = s->d_nodes_wide;
	p.nodes_wide128 = s->d_nodes_wide128;
	p.woop64 = s->d_woop64;
	p.woop = s->d_woop;
	p.tri_indices = s->d_tri_indices;
	p.rays = d_rays;
	p.dirs = d_dirs ? d_dirs : d_rays + 1;
	p.ray_stride = d_dirs ? 1u : 2u;
	p.n = n;
	p.n_ptr = d_n;
	p.out_tri = d_tri;
	p.out_t = d_t;
	p.out_uv = d_uv;
	p.out_occ = d_occ;
	p.magic = 0x4B000000u;
	p.counter = d_counter ? d_counter : s->d_counters + (s->counter_cursor.fetch_add(1u) % kCounterRing);
	p.refill_threshold = s->refill_threshold > 0 ? s->refill_threshold : 28;
	int per_sm = s->ctas_per_sm > 0 ? s->ctas_per_sm : (any ? s->occ_any : s->occ_closest);
	if (per_sm < 1) per_sm = 1;
	unsigned long long warps_needed = (n + 31) / 32;
	unsigned long long ctas_needed = (warps_needed + (kTraceBlock / 32) - 1) / (kTraceBlock / 32);
	unsigned grid = (unsigned)s->sm_count * (unsigned)per_sm;
	if (ctas_needed < grid) grid = (unsigned)ctas_needed;
	// pool size: big enough to amortise the atomic, small enough that every resident warp gets several pools
	// (a 1M-ray batch over 4736 warps would otherwise hand 3906 warps one 256-ray pool each and idle the rest)
	{
		const unsigned long long warps = (unsigned long long)grid * (kTraceBlock / 32);
		unsigned long long chunk = n / (warps * 6ull);
		chunk = (chunk / 32ull) * 32ull;
		p.pool_chunk = (uint32_t)(chunk < 32ull ? 32ull : chunk > kPoolChunk ? kPoolChunk : chunk);
		// guided tail: a request is 1 / 2^shift of the rays not yet handed out, 2^shift >= 4 x the warps (factors 1 / 2 / 4 / 8
		// measured 1.810 / 1.770 / 1.733 / 1.739 ms on C2 against 1.823 ms with fixed-size pools)
		p.guided_shift = 0;
		while ((1ull << p.guided_shift) < warps * 4ull) ++p.guided_shift;
	}
	ADYPT_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), stream));
	p.stats = d_stats;
	if (s->variant == 12 && d_stats == nullptr) {
		// EXPERIMENT (traverse_pool.cuh): rays held in shared memory and bound to lanes one round at a time
		const unsigned pool_grid = (unsigned)s->sm_count * 3u;
		const unsigned long long warps = (unsigned long long)pool_grid * (kPoolBlock / 32);
		unsigned long long chunk = n / (warps * 6ull);
		chunk = (chunk / 32ull) * 32ull;
		p.pool_chunk = (uint32_t)(chunk < 32ull ? 32ull : chunk > kPoolChunk ? kPoolChunk : chunk);
		p.guided_shift = 0;
		while ((1ull << p.guided_shift) < warps * 4ull) ++p.guided_shift;
		if (any) trace_pool_kernel<true><<<pool_grid, kPoolBlock, sizeof(PoolShared), stream>>>(p);
		else trace_pool_kernel<false><<<pool_grid, kPoolBlock, sizeof(PoolShared), stream>>>(p);
	} else
	trace_kernel_for(any, d_stats != nullptr, s->variant)<<<grid, kTraceBlock, 0, stream>>>(p);
	count_launch();
	ADYPT_CUDA(cudaGetLastError());
	return ADYPT_OK;
}

static int upload(void **dst, const void *src, size_t bytes, uint64_t *total)
{
	*dst = nullptr;
	if (bytes == 0) return ADYPT_OK;
	ADYPT_CUDA(cudaMalloc(dst, bytes));
	ADYPT_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
	*total += bytes;
	return ADYPT_OK;
}

static void free_scene(adypt_scene *s)
{
	DeviceGuard g(s->device);
	cudaFree(s->d_nodes);
	cudaFree(s->d_nodes_wide);
	cudaFree(s->d_nodes_wide128);
	cudaFree(s->d_woop64);
	cudaFree(s->d_woop);
	cudaFree(s->d_tri_indices);
	cudaFree(s->d_tris);
	cudaFree(s->d_mats);
	cudaFree(s->d_shade);
	cudaFree(s->d_tri_class);
	cudaFree(s->d_texels);
	cudaFree(s->d_tex_table);
	cudaFree(s->d_counters);
	s->stage_in.release();
	s->stage_out.release();
	for (int i = 0; i < 3; ++i)
		if (s->pipe[i]) cudaStreamDestroy(s->pipe[i]);
	delete s;
}

} // namespace adypt

using namespace adypt;

extern "C" {

const char *adypt_last_error(void) { return t_error.c_str(); }
int adypt_version(void) { return ADYPT_B200_VERSION; }

int adypt_device_count(int *count)
{
	return guarded([&]() -> int {
	if (!count) return fail(ADYPT_EINVAL, "count is NULL");
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) {
		*count = 0;
		return fail(ADYPT_ENODEV, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
	}
	*count = n;
	return ADYPT_OK;
	});
}

int adypt_launch_count(uint64_t *launches)
{
	return guarded([&]() -> int {
	if (!launches) return fail(ADYPT_EINVAL, "launches is NULL");
	*launches = g_launches.load();
	return ADYPT_OK;
	});
}

int adypt_scene_create(const adypt_scene_desc *d, adypt_scene **out)
{
	return guarded([&]() -> int {
	if (!d || !out) return fail(ADYPT_EINVAL, "desc/out is NULL");
	*out = nullptr;
	if (!d->nodes || d->n_nodes == 0) return fail(ADYPT_EINVAL, "scene needs at least the root node");
	if (d->n_refs && !d->tri_indices) return fail(ADYPT_EINVAL, "tri_indices is NULL");
	if (d->n_refs && !d->woop && !d->triangles) return fail(ADYPT_EINVAL, "need woop or triangles to build it from");
	if (d->n_tris && !d->triangles) return fail(ADYPT_EINVAL, "triangles is NULL");
	if (d->n_mats && !d->materials) return fail(ADYPT_EINVAL, "materials is NULL");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
		return fail(ADYPT_ENODEV, "no CUDA device: adypt_b200 has no CPU fallback");
	if (d->device < 0 || d->device >= ndev) return fail(ADYPT_ENODEV, "device ordinal out of range");
	// validate what traversal dereferences so a bad array cannot read out of bounds on the device
	for (uint32_t i = 0; i < d->n_refs; ++i)
		if (d->triangles && (d->tri_indices[i] < 0 || (uint32_t)d->tri_indices[i] >= d->n_tris))
			return fail(ADYPT_EINVAL, "tri_indices entry out of range");
	{
		const Node *nodes = (const Node *)d->nodes;
		for (uint32_t i = 0; i < d->n_nodes; ++i) {
			const Node &n = nodes[i];
			uint32_t n_inner = 0, max_tri = 0, inner_ordinals = 0;
			const uint32_t metas[2] = {n.meta_lo, n.meta_hi};
			for (int k = 0; k < 8; ++k) {
				const uint32_t m = (metas[k >> 2] >> (8 * (k & 3))) & 0xffu;
				if (m == 0) continue;
				if ((m & 0x1fu) >= 24u) { ++n_inner; inner_ordinals |= 1u << ((m & 0x1fu) - 24u); }
				else {
					const uint32_t cnt = __builtin_popcount(m >> 5), end = (m & 0x1fu) + cnt;
					if (end > max_tri) max_tri = end;
				}
			}
			// the kernel addresses an inner child as child_base + popc(imask & lowmask(ordinal)) (traversal.glsl:61-67)
			const uint32_t imask = n.head_w >> 24;
			if (n_inner && (uint64_t)n.child_base + n_inner > d->n_nodes) return fail(ADYPT_EINVAL, "node child index out of range");
			if ((inner_ordinals & ~imask) != 0u) return fail(ADYPT_EINVAL, "node has an inner child that its imask does not cover");
			if (imask && (uint64_t)n.child_base + (uint32_t)__builtin_popcount(imask) > d->n_nodes) return fail(ADYPT_EINVAL, "node imask reaches past the node array");
			if (max_tri && (uint64_t)n.tri_base + max_tri > d->n_refs) return fail(ADYPT_EINVAL, "node triangle index out of range");
		}
	}

	// Shading indexes materials[triangle.matid] (pathtracer.glsl:73-76). The OBJ ingest gives faces without a (known) material
	// id -1 like the reference (Scene.cpp:52); a GL buffer read shrugs that off, a CUDA one faults. Such a scene can still be
	// traversed; tracers refuse it (adypt_tracer_create).
	int64_t bad_matid_tri = -1;
	if (d->triangles && d->n_mats)
		for (uint32_t i = 0; i < d->n_tris; ++i) {
			int32_t matid;
			memcpy(&matid, (const uint8_t *)d->triangles + (size_t)i * 100u + 96u, 4);
			if (matid < 0 || (uint32_t)matid >= d->n_mats) { bad_matid_tri = (int64_t)i; break; }
		}

	DeviceGuard g(d->device);
	if (!g.ok) return fail(ADYPT_ENODEV, "cudaSetDevice failed");
	adypt_scene *s = new adypt_scene;
	s->bad_matid_tri = bad_matid_tri;
	s->device = d->device;
	s->n_nodes = d->n_nodes;
	s->n_refs = d->n_refs;
	s->n_tris = d->n_tris;
	s->n_mats = d->n_mats;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, d->device) != cudaSuccess) { delete s; return fail(ADYPT_ECUDA, "cudaGetDeviceProperties failed"); }
	s->sm_count = prop.multiProcessorCount;
	int rc;
#define UP(dst, src, bytes)                                                          \
	if ((rc = upload((void **)&(dst), (src), (bytes), &s->device_bytes)) != ADYPT_OK) { \
		free_scene(s);                                                               \
		return rc;                                                                   \
	}
	UP(s->d_nodes, d->nodes, (size_t)d->n_nodes * 80u);
	if (d->n_nodes) {
		if (cudaMalloc((void **)&s->d_nodes_wide, (size_t)d->n_nodes * 96u) != cudaSuccess) { free_scene(s); return fail(ADYPT_ENOMEM, "cudaMalloc wide nodes"); }
		s->device_bytes += (size_t)d->n_nodes * 96u;
		build_wide_nodes<<<(d->n_nodes + 127) / 128, 128>>>(s->d_nodes, d->n_nodes, s->d_nodes_wide);
		count_launch();
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { free_scene(s); return fail(ADYPT_ECUDA, std::string("build_wide_nodes: ") + cudaGetErrorString(e)); }
	}
	UP(s->d_tri_indices, d->tri_indices, (size_t)d->n_refs * 4u);
	UP(s->d_tris, d->triangles, (size_t)d->n_tris * 100u);
	UP(s->d_mats, d->materials, (size_t)d->n_mats * 64u);
	if (d->woop) {
		UP(s->d_woop, d->woop, (size_t)d->n_refs * 48u);
	} else if (d->n_refs) {
		if (cudaMalloc((void **)&s->d_woop, (size_t)d->n_refs * 48u) != cudaSuccess) { free_scene(s); return fail(ADYPT_ENOMEM, "cudaMalloc woop"); }
		s->device_bytes += (size_t)d->n_refs * 48u;
		build_woop_kernel<<<(d->n_refs + 255) / 256, 256>>>(s->d_tris, s->d_tri_indices, d->n_refs, s->d_woop);
		count_launch();
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { free_scene(s); return fail(ADYPT_ECUDA, std::string("build_woop_kernel: ") + cudaGetErrorString(e)); }
	}
#undef UP
	if (d->n_tris && d->n_mats) {
		if (cudaMalloc((void **)&s->d_shade, (size_t)d->n_tris * 128u) != cudaSuccess || cudaMalloc((void **)&s->d_tri_class, (size_t)d->n_tris * 4u) != cudaSuccess) {
			free_scene(s);
			return fail(ADYPT_ENOMEM, "cudaMalloc shading records");
		}
		s->device_bytes += (size_t)d->n_tris * 132u;
		build_shade_records<<<(d->n_tris + 127) / 128, 128>>>(s->d_tris, s->d_mats, d->n_tris, d->n_mats, s->d_shade, s->d_tri_class);
		count_launch();
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { free_scene(s); return fail(ADYPT_ECUDA, std::string("build_shade_records: ") + cudaGetErrorString(e)); }
	}
	if (cudaMalloc((void **)&s->d_counters, kCounterSlots * sizeof(unsigned long long)) != cudaSuccess) { free_scene(s); return fail(ADYPT_ENOMEM, "cudaMalloc counters"); }
	cudaFuncSetAttribute(trace_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PoolShared));
	cudaFuncSetAttribute(trace_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PoolShared));
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->occ_closest, trace_kernel<false>, kTraceBlock, 0);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->occ_any, trace_kernel<true>, kTraceBlock, 0);
	*out = s;
	return ADYPT_OK;
	});
}

int adypt_scene_set_textures(adypt_scene *s, const adypt_texture *tex, uint32_t n)
{
	return guarded([&]() -> int {
	if (!s || (n && !tex)) return fail(ADYPT_EINVAL, "NULL argument");
	DeviceGuard g(s->device);
	ADYPT_CUDA(cudaDeviceSynchronize());
	cudaFree(s->d_texels);
	cudaFree(s->d_tex_table);
	s->d_texels = nullptr;
	s->d_tex_table = nullptr;
	s->n_textures = 0;
	if (n == 0) return ADYPT_OK;
	std::vector<int4> table(n);
	size_t total = 0;
	for (uint32_t i = 0; i < n; ++i) {
		if (!tex[i].rgb8 || tex[i].width <= 0 || tex[i].height <= 0) return fail(ADYPT_EINVAL, "bad texture");
		if (total + (size_t)tex[i].width * (size_t)tex[i].height > 0x7fffffffull) return fail(ADYPT_ERANGE, "more than 2^31 texels");
		table[i] = make_int4((int)total, tex[i].width, tex[i].height, 0);
		total += (size_t)tex[i].width * (size_t)tex[i].height;
	}
	std::vector<uchar4> texels(total);
	for (uint32_t i = 0; i < n; ++i) {
		const size_t np = (size_t)tex[i].width * (size_t)tex[i].height;
		uchar4 *dst = texels.data() + table[i].x;
		for (size_t k = 0; k < np; ++k) dst[k] = make_uchar4(tex[i].rgb8[3 * k], tex[i].rgb8[3 * k + 1], tex[i].rgb8[3 * k + 2], 255);
	}
	// a material may only name a texture that exists
	std::vector<Material> mats(s->n_mats);
	if (s->n_mats) ADYPT_CUDA(cudaMemcpy(mats.data(), s->d_mats, (size_t)s->n_mats * 64u, cudaMemcpyDeviceToHost));
	for (const Material &m : mats)
		if (m.dtex < -1 || m.dtex >= (int32_t)n) return fail(ADYPT_EINVAL, "a material's diffuse texture index is out of range");
	ADYPT_CUDA(cudaMalloc((void **)&s->d_texels, total * 4u));
	ADYPT_CUDA(cudaMalloc((void **)&s->d_tex_table, (size_t)n * sizeof(int4)));
	ADYPT_CUDA(cudaMemcpy(s->d_texels, texels.data(), total * 4u, cudaMemcpyHostToDevice));
	ADYPT_CUDA(cudaMemcpy(s->d_tex_table, table.data(), (size_t)n * sizeof(int4), cudaMemcpyHostToDevice));
	s->n_textures = n;
	s->device_bytes += total * 4u + (size_t)n * sizeof(int4);
	return ADYPT_OK;
	});
}

int adypt_scene_destroy(adypt_scene *scene)
{
	return guarded([&]() -> int {
	if (!scene) return ADYPT_OK;
	free_scene(scene);
	return ADYPT_OK;
	});
}

int adypt_scene_read_woop(adypt_scene *s, float *out)
{
	return guarded([&]() -> int {
	if (!s || !out) return fail(ADYPT_EINVAL, "scene/out is NULL");
	DeviceGuard g(s->device);
	ADYPT_CUDA(cudaMemcpy(out, s->d_woop, (size_t)s->n_refs * 48u, cudaMemcpyDeviceToHost));
	return ADYPT_OK;
	});
}

int adypt_scene_device_bytes(adypt_scene *s, uint64_t *bytes)
{
	return guarded([&]() -> int {
	if (!s || !bytes) return fail(ADYPT_EINVAL, "scene/bytes is NULL");
	*bytes = s->device_bytes;
	return ADYPT_OK;
	});
}

int adypt_trace_configure(adypt_scene *s, int ctas_per_sm, int refill_threshold, int variant)
{
	return guarded([&]() -> int {
	if (!s) return fail(ADYPT_EINVAL, "scene is NULL");
	if (ctas_per_sm < 0 || refill_threshold < 0 || refill_threshold > 32 || variant < 0 || variant > 22) return fail(ADYPT_EINVAL, "bad tuning value");
	if (variant == 12 && !getenv("ADYPT_EXPERIMENTAL"))
		return fail(ADYPT_EINVAL, "variant 12 (shared-memory ray pool) is an experiment without a deep-stack path: set ADYPT_EXPERIMENTAL=1 to select it");
	if ((variant == 20 || variant == 21) && !s->d_nodes_wide128 && s->n_nodes) { // the experiment's node copy is built when it is first asked for
		DeviceGuard g(s->device);
		ADYPT_CUDA(cudaMalloc((void **)&s->d_nodes_wide128, (size_t)s->n_nodes * 128u));
		build_wide_nodes128<<<(s->n_nodes + 127) / 128, 128>>>(s->d_nodes, s->n_nodes, s->d_nodes_wide128);
		ADYPT_CUDA(cudaDeviceSynchronize());
	}
	if (variant == 22 && !s->d_woop64 && s->n_refs) { // the experiment's Woop copy is built when it is first asked for
		DeviceGuard g(s->device);
		ADYPT_CUDA(cudaMalloc((void **)&s->d_woop64, (size_t)s->n_refs * 64u));
		build_woop64<<<(s->n_refs + 255) / 256, 256>>>(s->d_woop, s->n_refs, s->d_woop64);
		ADYPT_CUDA(cudaDeviceSynchronize());
	}
	s->ctas_per_sm = ctas_per_sm;
	s->refill_threshold = refill_threshold;
	s->variant = variant;
	return ADYPT_OK;
	});
}

static int trace_host(adypt_scene *s, const float *rays, uint64_t n, int32_t *tri, float *t, float *uv, uint8_t *occ, cudaStream_t user_stream)
{
	// Host arrays: the batch is cut into chunks that flow through three internal streams, so the H2D copy of
	// chunk i+1, the traversal of chunk i and the D2H copy of chunk i-1 overlap (PCIe is full duplex). With
	// pinned host memory the whole call is bound by the 32 B/ray upload; pageable memory still works, just
	// without the overlap. Returns when every result is in the caller's arrays.
	const size_t in_bytes = (size_t)n * 32u;
	const size_t o_tri = 0, o_t = o_tri + (size_t)n * 4u, o_uv = o_t + (size_t)n * 4u, o_occ = o_uv + (size_t)n * 8u;
	const size_t out_bytes = o_occ + (size_t)n;
	ADYPT_TRY(s->stage_in.reserve(in_bytes));
	ADYPT_TRY(s->stage_out.reserve(out_bytes));
	if (!s->pipe[0])
		for (int i = 0; i < 3; ++i) ADYPT_CUDA(cudaStreamCreateWithFlags(&s->pipe[i], cudaStreamNonBlocking));
	ADYPT_CUDA(cudaStreamSynchronize(user_stream)); // earlier work queued by the caller on its stream comes first
	uint8_t *o = s->stage_out.as<uint8_t>();
	static const uint64_t chunk = []() -> uint64_t { // rays per pipeline stage (tunable for experiments)
		const char *e = getenv("ADYPT_HOST_CHUNK");
		const long v = e ? atol(e) : 0;
		return v >= 1024 ? (uint64_t)v : (uint64_t)(1u << 19);
	}();
	// The call's time is the upload's plus what cannot overlap it: the first chunk's upload (nothing to trace or download yet) and
	// the last chunk's traversal and download (nothing left to upload). So the first and last chunks are short -- 1/8, 1/4, 1/2 of a
	// full chunk on the way in, the same on the way out -- when the batch is long enough to have a steady state in between.
	static const bool ramp = []() { const char *e = getenv("ADYPT_HOST_RAMP"); return !e || atoi(e) != 0; }();
	const uint64_t ramp_rays = ramp ? ((chunk >> 3) + (chunk >> 2) + (chunk >> 1)) : 0; // one side
	const bool ramped = ramp && (chunk >> 3) >= 1024 && n >= 2 * ramp_rays + 2 * chunk;
	int k = 0;
	for (uint64_t b = 0; b < n; ++k) {
		uint64_t m = chunk;
		if (ramped) {
			const uint64_t left = n - b;
			if (k < 3) m = chunk >> (3 - k);
			else if (left <= ramp_rays) m = left > (chunk >> 1) + (chunk >> 3) ? chunk >> 1 : left > (chunk >> 3) ? chunk >> 2 : left;
			else if (left - ramp_rays < chunk) m = left - ramp_rays; // the steady state's short last chunk
		}
		if (m > n - b) m = n - b;
		cudaStream_t st = s->pipe[k % 3];
		unsigned long long *ctr = s->d_counters + kCounterPipe + (k % 3); // owned by that stream
		float4 *d_in = s->stage_in.as<float4>() + 2 * b;
		ADYPT_CUDA(cudaMemcpyAsync(d_in, rays + 8 * b, (size_t)m * 32u, cudaMemcpyHostToDevice, st));
		if (occ) {
			ADYPT_TRY(launch_trace(s, d_in, m, nullptr, nullptr, nullptr, o + o_occ + b, st, nullptr, ctr));
			ADYPT_CUDA(cudaMemcpyAsync(occ + b, o + o_occ + b, (size_t)m, cudaMemcpyDeviceToHost, st));
		} else {
			int32_t *d_tri = (int32_t *)(o + o_tri) + b;
			float *d_t = t ? (float *)(o + o_t) + b : nullptr;
			float2 *d_uv = uv ? (float2 *)(o + o_uv) + b : nullptr;
			ADYPT_TRY(launch_trace(s, d_in, m, d_tri, d_t, d_uv, nullptr, st, nullptr, ctr));
			ADYPT_CUDA(cudaMemcpyAsync(tri + b, d_tri, (size_t)m * 4u, cudaMemcpyDeviceToHost, st));
			if (t) ADYPT_CUDA(cudaMemcpyAsync(t + b, d_t, (size_t)m * 4u, cudaMemcpyDeviceToHost, st));
			if (uv) ADYPT_CUDA(cudaMemcpyAsync(uv + 2 * b, d_uv, (size_t)m * 8u, cudaMemcpyDeviceToHost, st));
		}
		b += m;
	}
	for (int i = 0; i < 3; ++i) ADYPT_CUDA(cudaStreamSynchronize(s->pipe[i]));
	return ADYPT_OK;
}

int adypt_trace_closest(adypt_scene *s, const float *rays, uint64_t n, int32_t *tri, float *t, float *uv, int memspace, void *stream)
{
	return guarded([&]() -> int {
	if (!s) return fail(ADYPT_EINVAL, "scene is NULL");
	if (n == 0) return ADYPT_OK;
	if (!rays || !tri) return fail(ADYPT_EINVAL, "rays/tri is NULL");
	if (memspace != ADYPT_MEM_HOST && memspace != ADYPT_MEM_DEVICE) return fail(ADYPT_EINVAL, "bad memspace");
	if (memspace == ADYPT_MEM_DEVICE && (((uintptr_t)rays & 15u) || ((uintptr_t)tri & 3u) || ((uintptr_t)t & 3u) || ((uintptr_t)uv & 7u)))
		return fail(ADYPT_EINVAL, "device arrays must be aligned: rays 16 B, tri/t 4 B, uv 8 B");
	DeviceGuard g(s->device);
	if (memspace == ADYPT_MEM_DEVICE)
		return launch_trace(s, (const float4 *)rays, n, tri, t, (float2 *)uv, nullptr, (cudaStream_t)stream);
	return trace_host(s, rays, n, tri, t, uv, nullptr, (cudaStream_t)stream);
	});
}

int adypt_trace_stats(adypt_scene *s, const float *rays, uint64_t n, int memspace, uint64_t out[4])
{
	return guarded([&]() -> int {
	if (!s || !out) return fail(ADYPT_EINVAL, "scene/out is NULL");
	out[0] = out[1] = out[2] = out[3] = 0;
	if (n == 0) return ADYPT_OK;
	if (!rays) return fail(ADYPT_EINVAL, "rays is NULL");
	DeviceGuard g(s->device);
	const float4 *d_rays = (const float4 *)rays;
	if (memspace == ADYPT_MEM_HOST) {
		ADYPT_TRY(s->stage_in.reserve((size_t)n * 32u));
		ADYPT_CUDA(cudaMemcpy(s->stage_in.ptr, rays, (size_t)n * 32u, cudaMemcpyHostToDevice));
		d_rays = s->stage_in.as<float4>();
	} else if (memspace != ADYPT_MEM_DEVICE)
		return fail(ADYPT_EINVAL, "bad memspace");
	ADYPT_TRY(s->stage_out.reserve((size_t)n * 4u + 64u));
	unsigned long long *d_stats = s->d_counters + kCounterStats;
	ADYPT_CUDA(cudaMemset(d_stats, 0, kStatSlots * sizeof(unsigned long long)));
	ADYPT_TRY(launch_trace(s, d_rays, n, s->stage_out.as<int32_t>(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, d_stats));
	ADYPT_CUDA(cudaDeviceSynchronize());
	unsigned long long h[4];
	ADYPT_CUDA(cudaMemcpy(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost));
	for (int i = 0; i < 4; ++i) out[i] = h[i];
	return ADYPT_OK;
	});
}

int adypt_trace_kernel_name(adypt_scene *s, int32_t any_hit, char *buf, uint64_t cap)
{
	return guarded([&]() -> int {
	if (!s || !buf || cap == 0) return fail(ADYPT_EINVAL, "NULL argument");
	DeviceGuard g(s->device);
	const char *name = nullptr;
	ADYPT_CUDA(cudaFuncGetName(&name, (const void *)trace_kernel_for(any_hit != 0, false, s->variant)));
	snprintf(buf, (size_t)cap, "%s", name ? name : "");
	return ADYPT_OK;
	});
}

int adypt_trace_any(adypt_scene *s, const float *rays, uint64_t n, uint8_t *occluded, int memspace, void *stream)
{
	return guarded([&]() -> int {
	if (!s) return fail(ADYPT_EINVAL, "scene is NULL");
	if (n == 0) return ADYPT_OK;
	if (!rays || !occluded) return fail(ADYPT_EINVAL, "rays/occluded is NULL");
	if (memspace != ADYPT_MEM_HOST && memspace != ADYPT_MEM_DEVICE) return fail(ADYPT_EINVAL, "bad memspace");
	if (memspace == ADYPT_MEM_DEVICE && ((uintptr_t)rays & 15u)) return fail(ADYPT_EINVAL, "device rays must be 16-byte aligned");
	DeviceGuard g(s->device);
	if (memspace == ADYPT_MEM_DEVICE)
		return launch_trace(s, (const float4 *)rays, n, nullptr, nullptr, nullptr, occluded, (cudaStream_t)stream);
	return trace_host(s, rays, n, nullptr, nullptr, nullptr, occluded, (cudaStream_t)stream);
	});
}

} // extern "C"
