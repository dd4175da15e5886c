// TEST INFRASTRUCTURE ONLY -- the CPU ORACLE. Never linked into, imported by or executed from the
// product path (adypt_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.
//
// A plain C++ restatement of the reference's GPU hot path (which is GLSL and cannot run here):
//   shaders/traversal.glsl:14-255   closest-hit CWBVH traversal      -> trace_one<false>
//   shaders/traversal.glsl:257-494  any-hit CWBVH traversal          -> trace_one<true>
//   src/Tracer/OglScene.cpp:93-116  Woop matrices per leaf reference -> oracle_build_woop
//   dep/glm func_matrix.inl:294-351 glm::inverse(mat4)               -> mat4_inverse
//   src/Tracer/Camera.cpp:13-23     view / projection matrices       -> oracle_camera_matrices
//   shaders/primaryray.glsl:39-44, pathtracer.glsl:213-218 Camera()  -> oracle_primary_rays
//   src/Util/Sobol.cpp:5-21         Gray-code Sobol                  -> oracle_sobol_*
//   shaders/pathtracer.glsl:49-227 + OglPathTracer.cpp:34-61 path tracer -> oracle_pt.inc
//
// PARITY PINNING: the reference ships no golden vectors or tests for this path and its GL
// program cannot run here, so the oracle is pinned against the reference's SOURCES compiled in place:
//  * host side (Woop rows, mat4 inverse, camera matrices, Sobol vectors, CWBVH arrays) against the reference's
//    own C++ (oracle/_ref/libadypt_ref.so, tests/test_oracle_pins.py);
//  * traversal and shading against the reference's own SHADER TEXT run on the CPU (oracle/_ref/libadypt_glsl.so,
//    built by glsl_transpile.py from shaders/*.glsl; tests/test_glsl_reference.py): ids, uv, any-hit bits,
//    viewer images and path-traced images agree bit for bit;
//  * plus an O(N) brute-force Woop test over all leaf references and hand-checked tiny scenes (tests/golden/).
// The GPU's arithmetic is whatever a GL driver's GLSL compiler emits, so the one thing DEFINED here rather than
// taken from the reference is the FP policy in DESIGN.md §3 ("un-fused left-to-right IEEE fp32, except explicit
// fmaf in the slab test and the Woop dot chains"); the shader build applies the same policy.
//
// Build: make -C oracle oracle   (-O3 -mavx2 -mfma -ffp-contract=off)
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <mutex>
#include <vector>

#include "detmath.h"

namespace {

// ---------------------------------------------------------------------------------------------
// data layouts

struct Node { // traversal.glsl:1-5 == WideBVH.hpp:13-26, 80 bytes
	float px, py, pz;
	uint32_t head_w; // ex | ey<<8 | ez<<16 | imask<<24
	uint32_t child_base, tri_base;
	uint32_t meta_lo, meta_hi;
	uint32_t lox_lo, lox_hi, loy_lo, loy_hi;
	uint32_t loz_lo, loz_hi, hix_lo, hix_hi;
	uint32_t hiy_lo, hiy_hi, hiz_lo, hiz_hi;
};
static_assert(sizeof(Node) == 80, "node");

struct Woop { float m0[4], m1[4], m2[4]; }; // traversal.glsl:6
static_assert(sizeof(Woop) == 48, "woop");

struct Ray { float ox, oy, oz, tmin, dx, dy, dz, pad; }; // batch ABI, 32 bytes
static_assert(sizeof(Ray) == 32, "ray");

struct Vec3 { float x, y, z; };

inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// FP policy helpers ----------------------------------------------------------------------------
// un-fused dot, left to right (GLSL dot / glm compute_dot) -- used by normalize()
inline float dot_lr(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// normalize(v) = v * (1 / sqrt(dot(v,v)))  (glm func_geometric.inl:82-90)
inline Vec3 normalize(Vec3 v)
{
	float inv = 1.0f / sqrtf(dot_lr(v, v));
	return Vec3{v.x * inv, v.y * inv, v.z * inv};
}
// fused dot used in the Woop test: fma(a.z,b.z, fma(a.y,b.y, a.x*b.x))
inline float dot_fma(const float *a3, float bx, float by, float bz)
{
	return fmaf(a3[2], bz, fmaf(a3[1], by, a3[0] * bx));
}
inline float max2(float a, float b) { return fmaxf(a, b); }
inline float min2(float a, float b) { return fminf(a, b); }

struct Counters { uint64_t nodes = 0, tris = 0, max_stack = 0, hits = 0; };

// ---------------------------------------------------------------------------------------------
// traversal.glsl:14-255 (closest) / :257-494 (any). Line numbers refer to the closest-hit overload.
template <bool ANY>
inline bool trace_one(const Node *nodes, const Woop *woop, const Ray &ray, int32_t *o_tri, float *o_u,
                      float *o_v, float *o_t, Counters &cnt)
{
	// :16-23 ray setup
	const float ooeps = 5.42101086242752217e-20f; // exp2(-64)
	Vec3 dir{ray.dx, ray.dy, ray.dz};
	dir.x = fabsf(dir.x) > ooeps ? dir.x : (dir.x >= 0 ? ooeps : -ooeps);
	dir.y = fabsf(dir.y) > ooeps ? dir.y : (dir.y >= 0 ? ooeps : -ooeps);
	dir.z = fabsf(dir.z) > ooeps ? dir.z : (dir.z >= 0 ? ooeps : -ooeps);
	dir = normalize(dir);
	const float idx = 1.0f / dir.x, idy = 1.0f / dir.y, idz = 1.0f / dir.z;
	const uint32_t octinv = 7u - ((dir.x < 0 ? 1u : 0u) | (dir.y < 0 ? 2u : 0u) | (dir.z < 0 ? 4u : 0u));
	const uint32_t octinv4 = octinv * 0x01010101u;
	const float ox = ray.ox, oy = ray.oy, oz = ray.oz;
	const float hit_tmin = ray.tmin;
	float hit_t = 1e9f; // :28
	int32_t hit_idx = -1;
	float hit_u = 0.f, hit_v = 0.f;

	uint32_t stack_x[64], stack_y[64]; // reference: TRAVERSAL_STACK_SIZE, unchecked (:12)
	int stack_ptr = 0;
	uint32_t tri_x = 0, tri_y = 0, node_x = 0, node_y = 0x80000000u; // :35

	while (true) {
		if (node_y > 0x00ffffffu) { // :47
			const uint32_t imask = node_y;
			const uint32_t child_bit_index = 31u - (uint32_t)__builtin_clz(node_y); // findMSB :52
			const uint32_t child_node_base_index = node_x;
			node_y &= ~(1u << child_bit_index);
			if (node_y > 0x00ffffffu) { // :59-60
				stack_x[stack_ptr] = node_x;
				stack_y[stack_ptr] = node_y;
				++stack_ptr;
				if ((uint64_t)stack_ptr > cnt.max_stack) cnt.max_stack = stack_ptr;
			}
			const uint32_t slot_index = (child_bit_index - 24u) ^ octinv;
			const uint32_t relative_index = (uint32_t)__builtin_popcount(imask & ~(0xffffffffu << slot_index));
			const Node &n = nodes[child_node_base_index + relative_index]; // :69-74
			++cnt.nodes;

			const float aix = as_float(((n.head_w) & 0xffu) << 23) * idx; // :76-78
			const float aiy = as_float(((n.head_w >> 8) & 0xffu) << 23) * idy;
			const float aiz = as_float(((n.head_w >> 16) & 0xffu) << 23) * idz;
			const float aox = (n.px - ox) * idx, aoy = (n.py - oy) * idy, aoz = (n.pz - oz) * idz; // :79

			node_x = n.child_base; // :81-83
			tri_x = n.tri_base;
			tri_y = 0;
			uint32_t hitmask = 0;
			for (int half = 0; half < 2; ++half) { // :86-143 and :145-202
				const uint32_t meta4 = half ? n.meta_hi : n.meta_lo;
				const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
				const uint32_t bit_index4 = (meta4 ^ (octinv4 & ((is_inner4 >> 4) * 0xffu))) & 0x1f1f1f1fu;
				const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
				const uint32_t qlox = half ? n.lox_hi : n.lox_lo, qhix = half ? n.hix_hi : n.hix_lo;
				const uint32_t qloy = half ? n.loy_hi : n.loy_lo, qhiy = half ? n.hiy_hi : n.hiy_lo;
				const uint32_t qloz = half ? n.loz_hi : n.loz_lo, qhiz = half ? n.hiz_hi : n.hiz_lo;
				const uint32_t s_lox = (idx < 0) ? qhix : qlox, s_hix = (idx < 0) ? qlox : qhix; // :92-99
				const uint32_t s_loy = (idy < 0) ? qhiy : qloy, s_hiy = (idy < 0) ? qloy : qhiy;
				const uint32_t s_loz = (idz < 0) ? qhiz : qloz, s_hiz = (idz < 0) ? qloz : qhiz;
				for (int k = 0; k < 4; ++k) { // :101-142
					const uint32_t sh = 8u * k;
					const float txmin = fmaf((float)((s_lox >> sh) & 0xffu), aix, aox);
					const float tymin = fmaf((float)((s_loy >> sh) & 0xffu), aiy, aoy);
					const float tzmin = fmaf((float)((s_loz >> sh) & 0xffu), aiz, aoz);
					const float txmax = fmaf((float)((s_hix >> sh) & 0xffu), aix, aox);
					const float tymax = fmaf((float)((s_hiy >> sh) & 0xffu), aiy, aoy);
					const float tzmax = fmaf((float)((s_hiz >> sh) & 0xffu), aiz, aoz);
					const float ctmin = max2(max2(txmin, tymin), max2(tzmin, hit_tmin));
					const float ctmax = min2(min2(txmax, tymax), min2(tzmax, hit_t));
					if (ctmin <= ctmax)
						hitmask |= ((child_bits4 >> sh) & 0xffu) << ((bit_index4 >> sh) & 0xffu);
				}
			}
			node_y = (hitmask & 0xff000000u) | ((n.head_w >> 24) & 0xffu); // :204-205
			tri_y = hitmask & 0x00ffffffu;
		} else { // :207-211 (dead in practice, kept)
			tri_x = node_x;
			tri_y = node_y;
			node_x = node_y = 0;
		}

		while (tri_y != 0) { // :213-243
			uint32_t tridx = (uint32_t)__builtin_ctz(tri_y); // findLSB
			tri_y &= ~(1u << tridx);
			tridx += tri_x;
			const Woop &w = woop[tridx];
			++cnt.tris;
			const float toz = w.m0[3] - dot_fma(w.m0, ox, oy, oz);
			const float tidz = 1.0f / dot_fma(w.m0, dir.x, dir.y, dir.z);
			const float tt = toz * tidz;
			const float tox = w.m1[3] + dot_fma(w.m1, ox, oy, oz);
			const float tdx = dot_fma(w.m1, dir.x, dir.y, dir.z);
			const float tu = fmaf(tt, tdx, tox);
			const float toy = w.m2[3] + dot_fma(w.m2, ox, oy, oz);
			const float tdy = dot_fma(w.m2, dir.x, dir.y, dir.z);
			const float tv = fmaf(tt, tdy, toy);
			if (tt > hit_tmin && tt < hit_t)
				if (tu >= 0.0f && tu <= 1.0f)
					if (tv >= 0.0f && tu + tv <= 1.0f) {
						hit_t = tt;
						if (ANY) { // :480-483
							++cnt.hits;
							if (o_t) *o_t = hit_t;
							return true;
						}
						hit_u = tu;
						hit_v = tv;
						hit_idx = (int32_t)tridx;
					}
		}

		if (node_y <= 0x00ffffffu) { // :245-250
			if (stack_ptr == 0) break;
			--stack_ptr;
			node_x = stack_x[stack_ptr];
			node_y = stack_y[stack_ptr];
		}
	}
	if (ANY) return false;
	if (hit_idx != -1) ++cnt.hits;
	*o_tri = hit_idx; // leaf-reference index; caller applies uTriIndices (:253-254)
	*o_u = hit_u;
	*o_v = hit_v;
	if (o_t) *o_t = hit_t;
	return hit_idx != -1;
}

template <class F> void parallel_chunks(uint64_t n, int nthreads, F &&fn)
{
	if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
	if (nthreads < 1) nthreads = 1;
	const uint64_t chunk = 4096; // dynamic 4096-ray chunks off an atomic counter
	std::atomic<uint64_t> next{0};
	auto worker = [&](int tid) {
		for (;;) {
			uint64_t b = next.fetch_add(chunk);
			if (b >= n) break;
			fn(tid, b, std::min(n, b + chunk));
		}
	};
	if (nthreads == 1) { worker(0); return; }
	std::vector<std::thread> th;
	for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
	for (auto &t : th) t.join();
}

// ---------------------------------------------------------------------------------------------
// glm::inverse(mat4) (func_matrix.inl:294-351), column-major m[col][row]
void mat4_inverse(const float in[16], float out[16])
{
	float m[4][4];
	memcpy(m, in, 64);
	float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
	float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
	float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
	float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
	float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
	float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
	float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
	float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
	float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
	float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
	float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
	float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
	float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
	float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
	float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
	float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
	float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
	float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
	const float Fac0[4] = {Coef00, Coef00, Coef02, Coef03}, Fac1[4] = {Coef04, Coef04, Coef06, Coef07};
	const float Fac2[4] = {Coef08, Coef08, Coef10, Coef11}, Fac3[4] = {Coef12, Coef12, Coef14, Coef15};
	const float Fac4[4] = {Coef16, Coef16, Coef18, Coef19}, Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
	const float Vec0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, Vec1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
	const float Vec2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, Vec3_[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
	const float SignA[4] = {+1, -1, +1, -1}, SignB[4] = {-1, +1, -1, +1};
	float inv[4][4];
	for (int i = 0; i < 4; ++i) {
		float Inv0 = Vec1[i] * Fac0[i] - Vec2[i] * Fac1[i] + Vec3_[i] * Fac2[i];
		float Inv1 = Vec0[i] * Fac0[i] - Vec2[i] * Fac3[i] + Vec3_[i] * Fac4[i];
		float Inv2 = Vec0[i] * Fac1[i] - Vec1[i] * Fac3[i] + Vec3_[i] * Fac5[i];
		float Inv3 = Vec0[i] * Fac2[i] - Vec1[i] * Fac4[i] + Vec2[i] * Fac5[i];
		inv[0][i] = Inv0 * SignA[i];
		inv[1][i] = Inv1 * SignB[i];
		inv[2][i] = Inv2 * SignA[i];
		inv[3][i] = Inv3 * SignB[i];
	}
	const float d0 = m[0][0] * inv[0][0], d1 = m[0][1] * inv[1][0], d2 = m[0][2] * inv[2][0], d3 = m[0][3] * inv[3][0];
	const float Dot1 = (d0 + d1) + (d2 + d3);
	const float OneOverDeterminant = 1.0f / Dot1;
	float *o = out;
	for (int c = 0; c < 4; ++c)
		for (int r = 0; r < 4; ++r) *o++ = inv[c][r] * OneOverDeterminant;
}

// glm::rotate(m, angle, axis) (ext/matrix_transform.inl:18-46) for the two axis-aligned uses in Camera.cpp
void mat4_rotate(float m[4][4], float angle, Vec3 v)
{
	const float c = cosf(angle), s = sinf(angle);
	const Vec3 axis = normalize(v);
	const float temp[3] = {(1.0f - c) * axis.x, (1.0f - c) * axis.y, (1.0f - c) * axis.z};
	const float ax[3] = {axis.x, axis.y, axis.z};
	float R[3][3];
	R[0][0] = c + temp[0] * ax[0];
	R[0][1] = temp[0] * ax[1] + s * ax[2];
	R[0][2] = temp[0] * ax[2] - s * ax[1];
	R[1][0] = temp[1] * ax[0] - s * ax[2];
	R[1][1] = c + temp[1] * ax[1];
	R[1][2] = temp[1] * ax[2] + s * ax[0];
	R[2][0] = temp[2] * ax[0] + s * ax[1];
	R[2][1] = temp[2] * ax[1] - s * ax[0];
	R[2][2] = c + temp[2] * ax[2];
	float res[4][4];
	for (int col = 0; col < 3; ++col)
		for (int r = 0; r < 4; ++r) res[col][r] = m[0][r] * R[col][0] + m[1][r] * R[col][1] + m[2][r] * R[col][2];
	for (int r = 0; r < 4; ++r) res[3][r] = m[3][r];
	memcpy(m, res, 64);
}

// ---------------------------------------------------------------------------------------------
// Sobol (Sobol.cpp:5-21) with direction numbers regenerated from Joe-Kuo parameters
// The parameter table is DATA shared with the product (one copy of the published Joe-Kuo constants in the tree, five packed
// words per dimension; layout in tools/derive_sobol_params.py); the decoder and the recurrence below are the oracle's own.
const uint32_t kJoeKuoWords[] = {
#include "../adypt_b200/csrc/sobol_joe_kuo.inc"
};
constexpr int kSobolTableDims = (int)(sizeof(kJoeKuoWords) / 20);
// Sobol.hpp:9 / Sobol.inl:4: 10 005 rows declared, 10 000 listed -- the rest are zero rows (constant 0 output)
constexpr int kSobolMaxDim = 10005;
std::vector<uint32_t> g_sobol_rows; // [kSobolMaxDim][32]
uint32_t (*g_sobol_v)[32] = nullptr;
std::once_flag g_sobol_once;

struct BitReader { // little-endian bit stream over one 160-bit record
	const uint32_t *w;
	int pos = 0;
	uint32_t take(int n)
	{
		uint32_t v = 0;
		for (int i = 0; i < n; ++i, ++pos) v |= ((w[pos >> 5] >> (pos & 31)) & 1u) << i;
		return v;
	}
};

void sobol_init()
{
	std::call_once(g_sobol_once, [] {
		g_sobol_rows.assign((size_t)kSobolMaxDim * 32u, 0u);
		g_sobol_v = reinterpret_cast<uint32_t (*)[32]>(g_sobol_rows.data());
		for (int j = 0; j < kSobolTableDims; ++j) {
			BitReader r{kJoeKuoWords + 5 * j};
			const int s = (int)r.take(5);
			const uint32_t a = r.take(16);
			uint32_t m[32];
			if (s == 0) {
				for (int k = 0; k < 32; ++k) m[k] = 1;
			} else {
				for (int k = 0; k < s; ++k) m[k] = (r.take(k) << 1) | 1u;
				for (int k = s; k < 32; ++k) {
					uint32_t v = m[k - s] ^ (m[k - s] << s);
					for (int i = 1; i < s; ++i)
						if ((a >> (s - 1 - i)) & 1u) v ^= m[k - i] << i;
					m[k] = v;
				}
			}
			for (int k = 0; k < 32; ++k) g_sobol_v[j][k] = m[k] << (31 - k);
		}
	});
}

inline uint32_t first_zero_bit(uint32_t x) { return x == 0xffffffffu ? 32u : (uint32_t)__builtin_ctz(~x); } // Sobol.cpp:5-14

#include "oracle_pt.inc"

} // namespace

// =============================================================================================
extern "C" {

int oracle_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// counters (nullable): [0] nodes visited, [1] triangles tested, [2] max stack depth, [3] rays that hit
int oracle_trace_closest(const void *nodes, const int32_t *tri_indices, const float *woop, const float *rays,
                         uint64_t n, int32_t *out_tri, float *out_t, float *out_uv, uint64_t *counters, int nthreads)
{
	const Node *N = (const Node *)nodes;
	const Woop *W = (const Woop *)woop;
	const Ray *R = (const Ray *)rays;
	int nt = nthreads <= 0 ? (int)std::thread::hardware_concurrency() : nthreads;
	std::vector<Counters> cs((size_t)std::max(nt, 1));
	parallel_chunks(n, nt, [&](int tid, uint64_t b, uint64_t e) {
		Counters &c = cs[tid];
		for (uint64_t i = b; i < e; ++i) {
			int32_t tri;
			float u, v, t;
			trace_one<false>(N, W, R[i], &tri, &u, &v, &t, c);
			out_tri[i] = tri >= 0 ? tri_indices[tri] : -1; // :253-254
			if (out_uv) { out_uv[2 * i] = u; out_uv[2 * i + 1] = v; }
			if (out_t) out_t[i] = t;
		}
	});
	if (counters) {
		counters[0] = counters[1] = counters[2] = counters[3] = 0;
		for (auto &c : cs) {
			counters[0] += c.nodes; counters[1] += c.tris; counters[3] += c.hits;
			counters[2] = std::max(counters[2], c.max_stack);
		}
	}
	return 0;
}

int oracle_trace_any(const void *nodes, const float *woop, const float *rays, uint64_t n, uint8_t *out_occluded,
                     uint64_t *counters, int nthreads)
{
	const Node *N = (const Node *)nodes;
	const Woop *W = (const Woop *)woop;
	const Ray *R = (const Ray *)rays;
	int nt = nthreads <= 0 ? (int)std::thread::hardware_concurrency() : nthreads;
	std::vector<Counters> cs((size_t)std::max(nt, 1));
	parallel_chunks(n, nt, [&](int tid, uint64_t b, uint64_t e) {
		Counters &c = cs[tid];
		for (uint64_t i = b; i < e; ++i) {
			int32_t tri;
			float u, v;
			out_occluded[i] = trace_one<true>(N, W, R[i], &tri, &u, &v, nullptr, c) ? 1 : 0;
		}
	});
	if (counters) {
		counters[0] = counters[1] = counters[2] = counters[3] = 0;
		for (auto &c : cs) {
			counters[0] += c.nodes; counters[1] += c.tris; counters[3] += c.hits;
			counters[2] = std::max(counters[2], c.max_stack);
		}
	}
	return 0;
}

// O(n_refs) brute force over ALL leaf references in ascending order with the same Woop test: the
// independent check of the traversal restatement.
int oracle_brute_closest(const int32_t *tri_indices, const float *woop, uint32_t n_refs, const float *rays, uint64_t n,
                         int32_t *out_tri, float *out_t, float *out_uv, int nthreads)
{
	const Woop *W = (const Woop *)woop;
	const Ray *R = (const Ray *)rays;
	parallel_chunks(n, nthreads, [&](int, uint64_t b, uint64_t e) {
		for (uint64_t i = b; i < e; ++i) {
			const Ray &ray = R[i];
			const float ooeps = 5.42101086242752217e-20f;
			Vec3 dir{ray.dx, ray.dy, ray.dz};
			dir.x = fabsf(dir.x) > ooeps ? dir.x : (dir.x >= 0 ? ooeps : -ooeps);
			dir.y = fabsf(dir.y) > ooeps ? dir.y : (dir.y >= 0 ? ooeps : -ooeps);
			dir.z = fabsf(dir.z) > ooeps ? dir.z : (dir.z >= 0 ? ooeps : -ooeps);
			dir = normalize(dir);
			float hit_t = 1e9f, hu = 0, hv = 0;
			int32_t hit = -1;
			for (uint32_t k = 0; k < n_refs; ++k) {
				const Woop &w = W[k];
				const float toz = w.m0[3] - dot_fma(w.m0, ray.ox, ray.oy, ray.oz);
				const float tidz = 1.0f / dot_fma(w.m0, dir.x, dir.y, dir.z);
				const float tt = toz * tidz;
				const float tox = w.m1[3] + dot_fma(w.m1, ray.ox, ray.oy, ray.oz);
				const float tu = fmaf(tt, dot_fma(w.m1, dir.x, dir.y, dir.z), tox);
				const float toy = w.m2[3] + dot_fma(w.m2, ray.ox, ray.oy, ray.oz);
				const float tv = fmaf(tt, dot_fma(w.m2, dir.x, dir.y, dir.z), toy);
				if (tt > ray.tmin && tt < hit_t && tu >= 0.0f && tu <= 1.0f && tv >= 0.0f && tu + tv <= 1.0f) {
					hit_t = tt; hu = tu; hv = tv; hit = (int32_t)k;
				}
			}
			out_tri[i] = hit >= 0 ? tri_indices[hit] : -1;
			if (out_t) out_t[i] = hit_t;
			if (out_uv) { out_uv[2 * i] = hu; out_uv[2 * i + 1] = hv; }
		}
	});
	return 0;
}

void oracle_mat4_inverse(const float in[16], float out[16]) { mat4_inverse(in, out); }

// OglScene::init_triangles (OglScene.cpp:93-116): tris100 = reference Triangle[] (100-byte records, the
// first 36 bytes are the three corner positions), one 48-byte Woop per leaf reference.
void oracle_build_woop(const void *tris100, const int32_t *tri_indices, uint32_t n_refs, float *out)
{
	const uint8_t *T = (const uint8_t *)tris100;
	for (uint32_t i = 0; i < n_refs; ++i) {
		float p[9];
		memcpy(p, T + (size_t)tri_indices[i] * 100, 36);
		const float *v0 = p, *v1 = p + 3, *v2 = p + 6;
		const float e0[3] = {v0[0] - v2[0], v0[1] - v2[1], v0[2] - v2[2]};
		const float e1[3] = {v1[0] - v2[0], v1[1] - v2[1], v1[2] - v2[2]};
		// glm::cross (func_geometric.inl): (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
		const float cr[3] = {e0[1] * e1[2] - e1[1] * e0[2], e0[2] * e1[0] - e1[2] * e0[0], e0[0] * e1[1] - e1[0] * e0[1]};
		// mat4 constructor takes column-major scalars: column0 = (c0.x, c1.x, c2.x, c3.x) ...
		float mtx[16] = {e0[0], e1[0], cr[0], v2[0], e0[1], e1[1], cr[1], v2[1],
		                 e0[2], e1[2], cr[2], v2[2], 0.f,   0.f,   0.f,   1.f};
		float inv[16];
		mat4_inverse(mtx, inv);
		float *o = out + (size_t)i * 12;
		o[0] = inv[8];  o[1] = inv[9];  o[2] = inv[10];  o[3] = -inv[11]; // mtx[2][0..3], w negated
		o[4] = inv[0];  o[5] = inv[1];  o[6] = inv[2];   o[7] = inv[3];   // mtx[0][..]
		o[8] = inv[4];  o[9] = inv[5];  o[10] = inv[6];  o[11] = inv[7];  // mtx[1][..]
	}
}

// Camera::GetView/GetProjection (Camera.cpp:13-23) + inverses (OglPathTracer.cpp:27-32)
void oracle_camera_matrices(float fov_deg, float yaw_deg, float pitch_deg, int width, int height, float proj[16],
                            float view[16], float inv_proj[16], float inv_view[16])
{
	const float kDeg = 0.01745329251994329576923690768489f; // glm::radians
	const float aspect = width / (float)height;
	float v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
	mat4_rotate(v, -pitch_deg * kDeg, Vec3{1.f, 0.f, 0.f});
	mat4_rotate(v, -yaw_deg * kDeg, Vec3{0.f, 1.f, 0.f});
	// tweakedInfinitePerspective(fovy, aspect, zNear=0.01, ep=FLT_EPSILON) (ext/matrix_clip_space.inl:512-527)
	const float fovy = fov_deg * kDeg, zNear = 0.01f, ep = 1.1920928955078125e-07f;
	const float range = tanf(fovy / 2.0f) * zNear;
	const float left = -range * aspect, right = range * aspect, bottom = -range, top = range;
	float p[4][4];
	memset(p, 0, sizeof(p));
	p[0][0] = (2.0f * zNear) / (right - left);
	p[1][1] = (2.0f * zNear) / (top - bottom);
	p[2][2] = ep - 1.0f;
	p[2][3] = -1.0f;
	p[3][2] = (ep - 2.0f) * zNear;
	memcpy(proj, p, 64);
	memcpy(view, v, 64);
	mat4_inverse(proj, inv_proj);
	mat4_inverse(view, inv_view);
}

// Camera() of primaryray.glsl:39-44 (bias 0) / pathtracer.glsl:213-218 (bias = SubPixel()), one ray per
// pixel in row-major order (pixel index = y*width + x); origin_tmin = uOrigin_TMin.
void oracle_primary_rays(const float origin_tmin[4], const float inv_proj[16], const float inv_view[16], int width,
                         int height, float bias_x, float bias_y, float *rays)
{
	for (int y = 0; y < height; ++y)
		for (int x = 0; x < width; ++x) {
			Vec3 d = camera_dir(inv_proj, inv_view, width, height, x, y, bias_x, bias_y);
			float *r = rays + ((size_t)y * width + x) * 8;
			r[0] = origin_tmin[0]; r[1] = origin_tmin[1]; r[2] = origin_tmin[2]; r[3] = origin_tmin[3];
			r[4] = d.x; r[5] = d.y; r[6] = d.z; r[7] = 0.f;
		}
}

int oracle_sobol_max_dim() { return kSobolMaxDim; }

// the 32 direction numbers of dimension j (0-based), for pinning against Sobol.inl
int oracle_sobol_directions(uint32_t dim_index, uint32_t out[32])
{
	sobol_init();
	if ((int)dim_index >= kSobolMaxDim) return -1;
	memcpy(out, g_sobol_v[dim_index], 128);
	return 0;
}

// Sobol::Reset(dim) then n_calls x Next (Sobol.cpp:16-21), sequential Gray-code form
int oracle_sobol_sequence(uint32_t dim, uint32_t n_calls, float *out)
{
	sobol_init();
	if ((int)dim > kSobolMaxDim) return -1;
	std::vector<uint32_t> x(dim, 0u);
	for (uint32_t idx = 0; idx < n_calls; ++idx) {
		const uint32_t c = first_zero_bit(idx);
		for (uint32_t j = 0; j < dim; ++j) {
			x[j] ^= g_sobol_v[j][c & 31u];
			out[(size_t)idx * dim + j] = (float)(x[j] / 4294967296.0);
		}
	}
	return 0;
}

// closed form: the vector Next() writes on its (index+1)-th call
int oracle_sobol_at(uint32_t dim, uint32_t index, float *out)
{
	sobol_init();
	if ((int)dim > kSobolMaxDim) return -1;
	sobol_vector(dim, index, out);
	return 0;
}

#include "oracle_pt_api.inc"

} // extern "C"
