// CWBVH traversal for sm_100a: closest-hit and any-hit, one ray per lane, persistent warps.
//
// Semantics are those of BVHIntersection in shaders/traversal.glsl:14-255 (closest) and :257-494 (any),
// ray by ray: children are visited in descending hit-bit order (findMSB, :52), triangles in ascending
// order (findLSB, :215), deferred groups LIFO (:59-60, :245-250), strict t comparisons (:235). Lanes never
// share a ray, so the per-ray visit order -- and with it every tie-break -- is the reference's. The warp
// cooperates on SCHEDULING only: lanes whose ray has finished are refilled from a warp-local pool of ray
// indices (ballot + popc ranking), and the pool is topped up with one atomicAdd per CHUNK rays.
//
// FP policy: IEEE fp32, round-to-nearest, no implicit contraction (explicit __f*_rn
// intrinsics), with fused multiply-add exactly where the oracle has fmaf(): the 48 slab evaluations per
// node and the Woop dot chains. The product kernel (MODE 2) issues those two at a time with Blackwell's packed
// fp32 instructions (fma / sub / mul .rn.f32x2 -> FFMA2 / FADD2 / FMUL2): each half is the scalar IEEE operation.
//
// Memory: node = 3 x 256-bit loads of the scene's 96-byte node copy (MODE 2; 5 x LDG.128 of the reference's
// 80-byte node in the older variants) and Woop = 3 x LDG.128 through the read-only path (L1 + L2; C1/C2 BVHs are
// L2-resident on B200), traversal stack = kSmemStack entries per lane in shared memory (per warp [entry][lane],
// so a warp's 8-byte accesses are conflict-free) with a local-memory overflow that real scenes never reach.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "layouts.h"

namespace adypt {

struct TraceParams {
	const uint4 *__restrict__ nodes;        // 5 per node
	const uint4 *__restrict__ nodes_wide;   // the derived 96-byte layout (scene.cu build_wide_nodes), MODE 2 kernels
	const uint4 *__restrict__ nodes_wide128; // experiment (MODE 3), built on demand
	const float4 *__restrict__ woop64;       // experiment (MODE 5), built on demand: 64-byte rows (48 bytes of Woop + 16 spare), 4 float4 per reference
	const float4 *__restrict__ woop;        // 3 per leaf reference
	const int32_t *__restrict__ tri_indices;
	const float4 *__restrict__ rays;        // origin + tmin of ray r at rays[r * ray_stride]
	const float4 *__restrict__ dirs;        // direction (+ pad) of ray r at dirs[r * ray_stride]; batch ABI (32-byte rays): dirs = rays + 1, stride 2;
	uint32_t ray_stride;                    // the wavefront's queues keep origins and directions in two arrays: stride 1
	unsigned long long n;
	const unsigned long long *n_ptr;        // when non-null the ray count is read from device memory (wavefront queues)
	int32_t *__restrict__ out_tri;          // closest
	float *__restrict__ out_t;              // closest, nullable
	float2 *__restrict__ out_uv;            // closest, nullable
	uint8_t *__restrict__ out_occ;          // any
	unsigned long long *counter;            // zeroed before launch
	unsigned long long *stats;              // STATS kernels only: [0] nodes visited [1] triangles tested [2] hits [3] max stack depth [4] rays
	int refill_threshold;                   // leave the traversal loop when fewer lanes than this are busy
	uint32_t magic;                         // 0x4B000000, passed as data so ptxas keeps it in a register (see byte_to_float)
	uint32_t pool_chunk;                    // most ray indices a warp takes per atomicAdd (multiple of 32, <= kPoolChunk)
	uint32_t guided_shift;                  // a warp asks for (rays not yet handed out) >> guided_shift, within [32, pool_chunk]
};

constexpr int kTraceBlock = 128;      // threads per CTA
#ifndef ADYPT_SMEM_STACK
#define ADYPT_SMEM_STACK 8
#endif
constexpr int kSmemStack = ADYPT_SMEM_STACK;  // stack entries per lane kept in shared memory
constexpr int kLocalStack = 64 - kSmemStack;  // overflow entries (local memory); total 64 like the oracle
constexpr unsigned kPoolChunk = 256;  // most ray indices a warp takes per atomicAdd (small batches take fewer, see launch_trace)
constexpr unsigned kFullMask = 0xffffffffu;


__device__ __forceinline__ void sts128(uint32_t addr, float4 v)
{
	asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
	return v;
}

// Guided self-scheduling of the ray pools: with a fixed 256-ray pool the warps of a launch finish up to one pool (~0.25 ms
// on C2) apart and the SMs idle through that tail; asking for a share of what is LEFT makes the last pools 32 rays long.
__device__ __forceinline__ uint32_t next_chunk(unsigned long long left, uint32_t shift, uint32_t most)
{
	const unsigned long long share = (left >> shift) & ~31ull;
	return share >= most ? most : (share < 32ull ? 32u : (uint32_t)share);
}

__device__ __forceinline__ float dot3_fma(float ax, float ay, float az, const float4 m)
{
	return __fmaf_rn(az, m.z, __fmaf_rn(ay, m.y, __fmul_rn(ax, m.x)));
}

// exact uint8 -> float, two interchangeable forms (both equal (float)byte bit for bit):
//  * MAGIC: splice the byte into the mantissa of 2^23 (PRMT with an IMMEDIATE selector against a register that
//    holds 0x4B000000) and subtract 2^23 (FADD): alu pipe + fma pipe, no conversion-pipe instruction. `magic`
//    arrives as a kernel parameter so ptxas cannot fold it: folded, it becomes PRMT's immediate and the four
//    selectors get re-materialised in registers for every use (45 IMAD.U32 per node step in profile r1a).
//  * CVT: I2F.U8 with a byte selector: one instruction, but on the quarter-rate conversion pipe.
// CVT_PLANES (0..6) of the six quantised planes use CVT, the rest MAGIC, to balance the pipes. Measured on B200
// (profiles/README.md, r1n_tune_tribatch.log): with the bounded triangle batch, 4 CVT planes at 8 CTAs/SM (64
// registers) is fastest; all-MAGIC and all-CVT both lose.
template <int K>
__device__ __forceinline__ float byte_to_float_magic(uint32_t word, uint32_t magic)
{
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(word), "r"(magic), "n"(0x7540 | K));
	return __fsub_rn(__uint_as_float(r), 8388608.0f);
}
template <int K>
__device__ __forceinline__ float byte_to_float_cvt(uint32_t word)
{
	return (float)((word >> (8 * K)) & 0xffu);
}
template <int K, bool CVT>
__device__ __forceinline__ float byte_to_float(uint32_t word, uint32_t magic)
{
	return CVT ? byte_to_float_cvt<K>(word) : byte_to_float_magic<K>(word, magic);
}

// one group of four children (one 32-bit lane of each quantised plane), traversal.glsl:86-143 / :145-202
template <int K, int CVT_PLANES, bool HM_LUT>
__device__ __forceinline__ uint32_t test_child(uint32_t meta_oct4, uint32_t s_lox, uint32_t s_loy, uint32_t s_loz,
                                               uint32_t s_hix, uint32_t s_hiy, uint32_t s_hiz, float aix, float aiy, float aiz,
                                               float aox, float aoy, float aoz, float tmin, float hit_t, uint32_t magic, const uint32_t *lut)
{
	const float txmin = __fmaf_rn(byte_to_float<K, (CVT_PLANES > 0)>(s_lox, magic), aix, aox);
	const float tymin = __fmaf_rn(byte_to_float<K, (CVT_PLANES > 2)>(s_loy, magic), aiy, aoy);
	const float tzmin = __fmaf_rn(byte_to_float<K, (CVT_PLANES > 4)>(s_loz, magic), aiz, aoz);
	const float txmax = __fmaf_rn(byte_to_float<K, (CVT_PLANES > 1)>(s_hix, magic), aix, aox);
	const float tymax = __fmaf_rn(byte_to_float<K, (CVT_PLANES > 3)>(s_hiy, magic), aiy, aoy);
	const float tzmax = __fmaf_rn(byte_to_float<K, (CVT_PLANES > 5)>(s_hiz, magic), aiz, aoz);
	const float ctmin = fmaxf(fmaxf(txmin, tymin), fmaxf(tzmin, tmin));
	const float ctmax = fminf(fminf(txmax, tymax), fminf(tzmax, hit_t));
	if (ctmin <= ctmax) {
		// the child's meta byte (count bits 7..5, bit index 4..0, already XOR-ed with the octant for inner children),
		// zero-extended by one PRMT; the funnel shift takes its amount modulo 32, so the index needs no mask
		const uint32_t b = __byte_perm(meta_oct4, 0u, 0x4440u | K);
		if (HM_LUT) {
			// the same value from a 256-entry table in shared memory: one LDS (and an address IMAD on the fma pipe) instead of
			// two shifts on the alu pipe, which is the busiest pipe of the node step
			return lut[b];
		}
		return __funnelshift_l(0u, b >> 5, b);
	}
	return 0u;
}

template <int CVT_PLANES, bool HM_LUT>
__device__ __forceinline__ uint32_t test_children4(uint32_t meta4, uint32_t octinv4, uint32_t s_lox, uint32_t s_loy,
                                                   uint32_t s_loz, uint32_t s_hix, uint32_t s_hiy, uint32_t s_hiz,
                                                   float aix, float aiy, float aiz, float aox, float aoy, float aoz,
                                                   float tmin, float hit_t, uint32_t magic, const uint32_t *lut)
{
	const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
	const uint32_t meta_oct4 = meta4 ^ (octinv4 & ((is_inner4 >> 4) * 0xffu)); // octinv only touches index bits 2..0 of inner children
#define ADYPT_CHILD(K) test_child<K, CVT_PLANES, HM_LUT>(meta_oct4, s_lox, s_loy, s_loz, s_hix, s_hiy, s_hiz, aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic, lut)
	return ADYPT_CHILD(0) | ADYPT_CHILD(1) | ADYPT_CHILD(2) | ADYPT_CHILD(3);
#undef ADYPT_CHILD
}

// The same four children with the slab evaluations issued two at a time: Blackwell's packed fp32 instructions (fma.rn.f32x2 /
// sub.rn.f32x2 -> FFMA2 / FADD2) take a register PAIR for the two children's quantised coordinates and the ray's scaled inverse
// direction and origin as scalars that SASS broadcasts (`R4.F32`), so a node step issues 24 FFMA2 + 8 FADD2 where it issued
// 48 FFMA + 16 FADD. Each half is the IEEE round-to-nearest fused multiply-add / subtraction of the scalar form: same bits.
__device__ __forceinline__ unsigned long long pack2(float a, float b)
{
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
	return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &a, float &b)
{
	asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
// (o . m.xyz, d . m.xyz) with the two dot products side by side: three packed instructions instead of six, each half evaluated
// exactly like dot3_fma (fma(z, m.z, fma(y, m.y, x * m.x)))
__device__ __forceinline__ void dot3_fma_pair(float ox, float dx, float oy, float dy, float oz, float dz, const float4 m, float &dot_o, float &dot_d)
{
	unsigned long long r;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(ox, dx)), "l"(pack2(m.x, m.x)));
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pack2(oy, dy)), "l"(pack2(m.y, m.y)), "l"(r));
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pack2(oz, dz)), "l"(pack2(m.z, m.z)), "l"(r));
	unpack2(r, dot_o, dot_d);
}
// t, u, v of the Woop test (:221-233) for rows M0..M2
template <bool PACKED>
__device__ __forceinline__ void woop_eval(float ox, float oy, float oz, float dx, float dy, float dz, const float4 M0, const float4 M1, const float4 M2,
                                          float &tt, float &tu, float &tv)
{
	float o0, d0, o1, d1, o2, d2;
	if (PACKED) {
		dot3_fma_pair(ox, dx, oy, dy, oz, dz, M0, o0, d0);
		dot3_fma_pair(ox, dx, oy, dy, oz, dz, M1, o1, d1);
		dot3_fma_pair(ox, dx, oy, dy, oz, dz, M2, o2, d2);
	} else {
		o0 = dot3_fma(ox, oy, oz, M0); d0 = dot3_fma(dx, dy, dz, M0);
		o1 = dot3_fma(ox, oy, oz, M1); d1 = dot3_fma(dx, dy, dz, M1);
		o2 = dot3_fma(ox, oy, oz, M2); d2 = dot3_fma(dx, dy, dz, M2);
	}
	const float toz = __fsub_rn(M0.w, o0);
	const float tidz = __frcp_rn(d0);
	tt = __fmul_rn(toz, tidz);
	const float tox = __fadd_rn(M1.w, o1);
	tu = __fmaf_rn(tt, d1, tox);
	const float toy = __fadd_rn(M2.w, o2);
	tv = __fmaf_rn(tt, d2, toy);
}

// (float)byte K and K + 1 of `word`, times a, plus c
template <int K, bool CVT>
__device__ __forceinline__ void plane2(uint32_t word, uint32_t magic, float a, float c, float &t0, float &t1)
{
	unsigned long long q;
	if (CVT)
		q = pack2(byte_to_float_cvt<K>(word), byte_to_float_cvt<K + 1>(word));
	else {
		uint32_t r0, r1;
		asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r0) : "r"(word), "r"(magic), "n"(0x7540 | K));
		asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r1) : "r"(word), "r"(magic), "n"(0x7540 | (K + 1)));
		const float m = __uint_as_float(magic); // 2^23
		asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(q) : "l"(pack2(__uint_as_float(r0), __uint_as_float(r1))), "l"(pack2(m, m)));
	}
	unsigned long long r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(q), "l"(pack2(a, a)), "l"(pack2(c, c)));
	unpack2(r, t0, t1);
}
template <int K>
__device__ __forceinline__ uint32_t finish_child(uint32_t meta_oct4, float txmin, float tymin, float tzmin, float txmax, float tymax, float tzmax,
                                                 float tmin, float hit_t)
{
	const float ctmin = fmaxf(fmaxf(txmin, tymin), fmaxf(tzmin, tmin));
	const float ctmax = fminf(fminf(txmax, tymax), fminf(tzmax, hit_t));
	if (ctmin <= ctmax) {
		const uint32_t b = __byte_perm(meta_oct4, 0u, 0x4440u | K);
		return __funnelshift_l(0u, b >> 5, b);
	}
	return 0u;
}
template <int K, int CVT_PLANES>
__device__ __forceinline__ uint32_t test_child_pair(uint32_t meta_oct4, uint32_t s_lox, uint32_t s_loy, uint32_t s_loz, uint32_t s_hix, uint32_t s_hiy,
                                                    uint32_t s_hiz, float aix, float aiy, float aiz, float aox, float aoy, float aoz, float tmin, float hit_t,
                                                    uint32_t magic)
{
	float x0, x1, y0, y1, z0, z1, X0, X1, Y0, Y1, Z0, Z1;
	plane2<K, (CVT_PLANES > 0)>(s_lox, magic, aix, aox, x0, x1);
	plane2<K, (CVT_PLANES > 2)>(s_loy, magic, aiy, aoy, y0, y1);
	plane2<K, (CVT_PLANES > 4)>(s_loz, magic, aiz, aoz, z0, z1);
	plane2<K, (CVT_PLANES > 1)>(s_hix, magic, aix, aox, X0, X1);
	plane2<K, (CVT_PLANES > 3)>(s_hiy, magic, aiy, aoy, Y0, Y1);
	plane2<K, (CVT_PLANES > 5)>(s_hiz, magic, aiz, aoz, Z0, Z1);
	return finish_child<K>(meta_oct4, x0, y0, z0, X0, Y0, Z0, tmin, hit_t) | finish_child<K + 1>(meta_oct4, x1, y1, z1, X1, Y1, Z1, tmin, hit_t);
}
template <int CVT_PLANES>
__device__ __forceinline__ uint32_t test_children4_packed(uint32_t meta4, uint32_t octinv4, uint32_t s_lox, uint32_t s_loy, uint32_t s_loz, uint32_t s_hix,
                                                          uint32_t s_hiy, uint32_t s_hiz, float aix, float aiy, float aiz, float aox, float aoy, float aoz,
                                                          float tmin, float hit_t, uint32_t magic)
{
	const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
	const uint32_t meta_oct4 = meta4 ^ (octinv4 & ((is_inner4 >> 4) * 0xffu));
	return test_child_pair<0, CVT_PLANES>(meta_oct4, s_lox, s_loy, s_loz, s_hix, s_hiy, s_hiz, aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic) |
	       test_child_pair<2, CVT_PLANES>(meta_oct4, s_lox, s_loy, s_loz, s_hix, s_hiy, s_hiz, aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic);
}

// Wide nodes (MODE 2): the scene keeps a 96-byte, 32-byte aligned copy of every node -- the reference's 80 bytes followed by the three
// plane scales 2^ex, 2^ey, 2^ez as floats -- so that a visit fetches it with THREE 256-bit loads (Blackwell's ld.global.v8.b32) instead of
// five 128-bit ones, and does not rebuild the scales from their exponent bytes. Divergent lanes cost the L1 a tag look-up per lane and
// LOAD, and the L1 is the kernel's second-busiest unit (63 %) after instruction issue.
struct Words8 { uint32_t v[8]; };
__device__ __forceinline__ Words8 ldg256(const void *p)
{
	Words8 r;
	asm("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
	    : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
	return r;
}

// EXPERIMENT (MODE 3): 128-byte nodes that also spell out what each child adds to the hit mask (c[k] = (meta[k] >> 5) << (meta[k] & 31)),
// so a visit ORs words in and applies the octant once (three delta swaps on the byte of inner hits) instead of decoding eight meta bytes
// (~55 -> ~25 instructions). Rows of 32 bytes: [p, imask | scales, child_base] [lox loy loz hix] [hiy hiz | c0..c3] [c4..c7 | tri_base].
template <int K, int CVT_PLANES>
__device__ __forceinline__ uint32_t test_child_pair_wide(uint32_t c0, uint32_t c1, uint32_t s_lox, uint32_t s_loy, uint32_t s_loz, uint32_t s_hix,
                                                         uint32_t s_hiy, uint32_t s_hiz, float aix, float aiy, float aiz, float aox, float aoy, float aoz,
                                                         float tmin, float hit_t, uint32_t magic)
{
	float x0, x1, y0, y1, z0, z1, X0, X1, Y0, Y1, Z0, Z1;
	plane2<K, (CVT_PLANES > 0)>(s_lox, magic, aix, aox, x0, x1);
	plane2<K, (CVT_PLANES > 2)>(s_loy, magic, aiy, aoy, y0, y1);
	plane2<K, (CVT_PLANES > 4)>(s_loz, magic, aiz, aoz, z0, z1);
	plane2<K, (CVT_PLANES > 1)>(s_hix, magic, aix, aox, X0, X1);
	plane2<K, (CVT_PLANES > 3)>(s_hiy, magic, aiy, aoy, Y0, Y1);
	plane2<K, (CVT_PLANES > 5)>(s_hiz, magic, aiz, aoz, Z0, Z1);
	const float lo0 = fmaxf(fmaxf(x0, y0), fmaxf(z0, tmin)), hi0 = fminf(fminf(X0, Y0), fminf(Z0, hit_t));
	const float lo1 = fmaxf(fmaxf(x1, y1), fmaxf(z1, tmin)), hi1 = fminf(fminf(X1, Y1), fminf(Z1, hit_t));
	return (lo0 <= hi0 ? c0 : 0u) | (lo1 <= hi1 ? c1 : 0u);
}
__device__ __forceinline__ uint32_t permute_inner_slots(uint32_t h, uint32_t octinv)
{
#pragma unroll
	for (uint32_t j = 1u; j <= 4u; j <<= 1) {
		const uint32_t sh = octinv & j; // 0 or j: a delta swap by 0 is the identity
		const uint32_t m = (j == 1u ? 0x55u : j == 2u ? 0x33u : 0x0fu) << 24;
		const uint32_t t = ((h >> sh) ^ h) & m;
		h ^= t | (t << sh);
	}
	return h;
}

// TRI_BATCH: 0 = a lane tests all triangles of its group before the warp moves on (the GLSL's loop shape);
// K > 0 = at most K triangle tests per lane and round, lanes with triangles left skip their next node step until
// the group is empty; 12 = K 2 with both triangles' rows fetched before the first test. Only the warp-level
// interleaving differs: each ray's own sequence of node steps and triangle tests -- and therefore every result and
// counter -- is the same. A warp no longer waits for its one lane with nine triangles: -12 % time on C2.
//
// STAGED: ray set-up (two 16-byte loads, clamp, IEEE normalise, three IEEE reciprocals: ~100 instructions) is done by the
// whole warp for the next 32 rays of its pool at once and parked in shared memory (48 B per ray); an idle lane then picks
// its ray up with three LDS.128. Unstaged, the same code runs at every refill for the ~6 lanes that happen to be idle, and
// the warp waits on the (cold, HBM) ray loads five times as often. Same arithmetic per ray, so same results.
template <bool ANY, bool STATS = false, int CVT_PLANES = 4, int MIN_CTAS = 8, int TRI_BATCH = 12, bool STAGED = true, bool HM_LUT = false, int MODE = 0>
__global__ void __launch_bounds__(kTraceBlock, MIN_CTAS) trace_kernel(const TraceParams p)
{
	constexpr bool PACKED = MODE >= 1; // slab and Woop evaluations as packed fp32 pairs
	constexpr bool WIDE = MODE == 2 || MODE == 5; // 96-byte nodes fetched with three 256-bit loads
	constexpr bool WOOP64 = MODE == 5; // experiment: 64-byte Woop rows fetched with two 256-bit loads
	constexpr bool WIDE128 = MODE == 3; // experiment: 128-byte nodes with hit-mask words, four 256-bit loads

	const uint32_t magic = p.magic;
	constexpr int NODE_REPS = (ANY && TRI_BATCH > 0) ? 2 : 1; // the unbounded triangle loop (TRI_BATCH 0) has no "triangles left over" state to skip a node step with
	// Shared memory, one region per warp: [stack: kSmemStack entries x 32 lanes x 8 B][stage: 3 rows x 32 entries x 16 B].
	// Every address is formed from ONE register, lane_addr = region + 8 * lane (stack entry sp of this lane is at
	// lane_addr + 256 * sp; 8-byte accesses of a warp are conflict-free), and so are the lane number and its lt-mask:
	// left to itself ptxas re-derives all of these from S2R tid / cta-id reads at every push, pop and refill.
	constexpr unsigned kWarpStack = kSmemStack * 256u, kWarpRegion = kWarpStack + (STAGED ? 3u * 512u : 0u);
	static_assert(kWarpRegion % 256u == 0u, "lane bits are taken from the low byte of the address");
	__shared__ __align__(256) unsigned char s_mem[(kTraceBlock / 32) * kWarpRegion];
	uint2 l_stack[kLocalStack];
	__shared__ uint32_t s_hm_lut[HM_LUT ? 256 : 1]; // meta byte -> (count bits) << (bit index), see test_child
	if (HM_LUT) {
		for (unsigned i = threadIdx.x; i < 256u; i += kTraceBlock) s_hm_lut[i] = (i >> 5) << (i & 31u);
		__syncthreads();
	}

	uint32_t lane_addr = (uint32_t)__cvta_generic_to_shared(s_mem) + (threadIdx.x >> 5) * kWarpRegion + (threadIdx.x & 31u) * 8u;
	asm volatile("mov.u32 %0, %0;" : "+r"(lane_addr)); // opaque, so that it is held in a register (-3.6 % time)
	const unsigned lane = (lane_addr >> 3) & 31u;
	const unsigned lt_mask = ~(0xffffffffu << lane);
	const unsigned long long n_rays = p.n_ptr ? *p.n_ptr : p.n;
	const uint4 *nodes_base = WIDE ? p.nodes_wide : WIDE128 ? p.nodes_wide128 : p.nodes;
	const float4 *woop_base = WOOP64 ? p.woop64 : p.woop;
#ifdef ADYPT_PTR_REGS
	asm volatile("mov.u64 %0, %0;" : "+l"(nodes_base));
#if ADYPT_PTR_REGS > 1
	asm volatile("mov.u64 %0, %0;" : "+l"(woop_base));
#endif
#endif

	// warp-uniform pool of ray indices
	unsigned long long pool_next = 0, pool_end = 0;
	bool exhausted = false;
	// STAGED: entries [stage_pos, stage_cnt) of the warp's 32-entry stage hold rays stage_base + entry, set up and untaken
	unsigned long long stage_base = 0;
	unsigned stage_pos = 0, stage_cnt = 0;

	// per-lane ray state. sp doubles as the lane's status: >= 0 traversing (stack depth), kIdle = no ray,
	// kUnsaved = ray finished, result still in registers (stored at the next refill)
	constexpr int kIdle = -1, kUnsaved = -2;
	unsigned long long ray_idx = 0;
	float ox = 0, oy = 0, oz = 0, tmin = 0, dx = 0, dy = 0, dz = 0, idx = 0, idy = 0, idz = 0;
	uint32_t octinv = 0;
	float hit_t = 0, hit_u = 0, hit_v = 0;
	int32_t hit_idx = -1;
	uint2 ng = make_uint2(0, 0), tg = make_uint2(0, 0);
	int sp = kIdle;
	unsigned long long st_nodes = 0, st_tris = 0, st_hits = 0, st_depth = 0; // dead code unless STATS

	for (;;) {
		// ---------------------------------------------------------------- refill idle lanes
		unsigned idle = __ballot_sync(kFullMask, sp < 0);
		// Results of rays that finished since the last refill are written here, by all such lanes together
		// (about 6 per pass), instead of by 1-3 lanes at the moment each ray ends.
		if (sp == kUnsaved) {
			sp = kIdle;
			if (STATS && hit_t < 1e9f) ++st_hits;
			if (ANY) {
				p.out_occ[ray_idx] = (hit_t < 1e9f) ? 1 : 0;
			} else {
				p.out_tri[ray_idx] = hit_idx >= 0 ? __ldg(p.tri_indices + hit_idx) : -1; // :253-254
				if (p.out_t) p.out_t[ray_idx] = hit_t;
				if (p.out_uv) p.out_uv[ray_idx] = make_float2(hit_u, hit_v);
			}
		}
		if (STAGED) {
			while (idle != 0) {
				if (stage_pos == stage_cnt) {
					if (exhausted) break;
					if (pool_next >= pool_end) {
						unsigned long long b = 0;
						const uint32_t request = next_chunk(n_rays - pool_end, p.guided_shift, p.pool_chunk); // pool_end: where this warp's last pool ended
						if (lane == 0) b = atomicAdd(p.counter, (unsigned long long)request);
						b = __shfl_sync(kFullMask, b, 0);
						if (b >= n_rays) { exhausted = true; break; }
						pool_next = b;
						pool_end = (b + request < n_rays) ? b + request : n_rays;
					}
					const unsigned long long left = pool_end - pool_next;
					stage_cnt = left < 32ull ? (unsigned)left : 32u;
					stage_base = pool_next;
					pool_next += stage_cnt;
					stage_pos = 0;
					__syncwarp(); // the previous stage's readers are done
					if (lane < stage_cnt) {
						// ray setup, traversal.glsl:16-35
						const unsigned long long r = stage_base + lane;
						const float4 r0 = __ldg(p.rays + r * p.ray_stride), r1 = __ldg(p.dirs + r * p.ray_stride);
						const float ooeps = 5.42101086242752217e-20f; // exp2(-64)
						float sx = fabsf(r1.x) > ooeps ? r1.x : (r1.x >= 0.0f ? ooeps : -ooeps);
						float sy = fabsf(r1.y) > ooeps ? r1.y : (r1.y >= 0.0f ? ooeps : -ooeps);
						float sz = fabsf(r1.z) > ooeps ? r1.z : (r1.z >= 0.0f ? ooeps : -ooeps);
						const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy)), __fmul_rn(sz, sz));
						const float inv = __frcp_rn(__fsqrt_rn(len2)); // 1/x correctly rounded == IEEE 1.0f / x
						sx = __fmul_rn(sx, inv); sy = __fmul_rn(sy, inv); sz = __fmul_rn(sz, inv);
						const uint32_t oi = 7u - ((sx < 0.0f ? 1u : 0u) | (sy < 0.0f ? 2u : 0u) | (sz < 0.0f ? 4u : 0u));
						const uint32_t slot = lane_addr + (lane_addr & 0xffu) + kWarpStack; // region + stack + 16 * lane
						if (PACKED) { // origin and direction interleaved: the lane that takes the ray gets (ox, dx), (oy, dy), (oz, dz) as register pairs
							sts128(slot, make_float4(r0.x, sx, r0.y, sy));
							sts128(slot + 512u, make_float4(r0.z, sz, r0.w, __uint_as_float(oi)));
						} else {
							sts128(slot, r0);
							sts128(slot + 512u, make_float4(sx, sy, sz, __uint_as_float(oi)));
						}
						sts128(slot + 1024u, make_float4(__frcp_rn(sx), __frcp_rn(sy), __frcp_rn(sz), 0.0f));
					}
					__syncwarp();
				}
				const unsigned cand = stage_pos + (unsigned)__popc(idle & lt_mask);
				const bool take = sp < 0 && cand < stage_cnt;
				if (take) {
					const uint32_t slot = (lane_addr & ~0xffu) + kWarpStack + cand * 16u;
					const float4 a = lds128(slot), b = lds128(slot + 512u), c = lds128(slot + 1024u);
					ray_idx = stage_base + cand;
					if (PACKED) {
						ox = a.x; dx = a.y; oy = a.z; dy = a.w;
						oz = b.x; dz = b.y; tmin = b.z; octinv = __float_as_uint(b.w);
					} else {
						ox = a.x; oy = a.y; oz = a.z; tmin = a.w;
						dx = b.x; dy = b.y; dz = b.z; octinv = __float_as_uint(b.w);
					}
					idx = c.x; idy = c.y; idz = c.z;
					hit_t = 1e9f; hit_idx = -1; hit_u = 0.0f; hit_v = 0.0f;
					ng = make_uint2(0u, 0x80000000u);
					tg = make_uint2(0u, 0u);
					sp = 0;
				}
				const unsigned took = __ballot_sync(kFullMask, take);
				stage_pos += (unsigned)__popc(took);
				idle &= ~took;
			}
		} else
		while (idle != 0 && !exhausted) {
			if (pool_next >= pool_end) {
				unsigned long long b = 0;
				const uint32_t request = next_chunk(n_rays - pool_end, p.guided_shift, p.pool_chunk);
				if (lane == 0) b = atomicAdd(p.counter, (unsigned long long)request);
				b = __shfl_sync(kFullMask, b, 0);
				if (b >= n_rays) { exhausted = true; break; }
				pool_next = b;
				pool_end = (b + request < n_rays) ? b + request : n_rays;
			}
			const unsigned long long cand = pool_next + __popc(idle & lt_mask);
			const bool take = sp < 0 && cand < pool_end;
			if (take) {
				// ray setup, traversal.glsl:16-35
				ray_idx = cand;
				const float4 r0 = __ldg(p.rays + cand * p.ray_stride), r1 = __ldg(p.dirs + cand * p.ray_stride);
				ox = r0.x; oy = r0.y; oz = r0.z; tmin = r0.w;
				const float ooeps = 5.42101086242752217e-20f; // exp2(-64)
				dx = fabsf(r1.x) > ooeps ? r1.x : (r1.x >= 0.0f ? ooeps : -ooeps);
				dy = fabsf(r1.y) > ooeps ? r1.y : (r1.y >= 0.0f ? ooeps : -ooeps);
				dz = fabsf(r1.z) > ooeps ? r1.z : (r1.z >= 0.0f ? ooeps : -ooeps);
				const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
				const float inv = __frcp_rn(__fsqrt_rn(len2)); // 1/x correctly rounded == IEEE 1.0f / x
				dx = __fmul_rn(dx, inv); dy = __fmul_rn(dy, inv); dz = __fmul_rn(dz, inv);
				idx = __frcp_rn(dx); idy = __frcp_rn(dy); idz = __frcp_rn(dz);
				octinv = 7u - ((dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 4u : 0u));
				hit_t = 1e9f; hit_idx = -1; hit_u = 0.0f; hit_v = 0.0f;
				ng = make_uint2(0u, 0x80000000u);
				tg = make_uint2(0u, 0u);
				sp = 0;
			}
			const unsigned took = __ballot_sync(kFullMask, take);
			pool_next += __popc(took);
			idle &= ~took;
		}
		if (!__any_sync(kFullMask, sp >= 0)) break;

		// ---------------------------------------------------------------- traverse
		unsigned busy;
		do {
			if (sp >= 0) {
				bool finished = false;
				// NODE_REPS node steps (with the pop a lane needs in between), then ONE triangle section: any-hit rays test
				// few triangles, so the section runs for half as many rounds at twice the lanes (+7 %); for closest-hit
				// rays the lanes that wait with their triangles through the second node step cost as much as that saves
				// (measured 1.80 vs 1.76 ms), so they keep one node step per round.
#pragma unroll 1
				for (int rep = 0; rep < NODE_REPS; ++rep) {
				if (TRI_BATCH > 0 && tg.y != 0u) {
					// triangles left over from the previous round: no node step yet
				} else if (ng.y > 0x00ffffffu) {
					// n <- closest child of G (:50-67)
					const uint32_t imask = ng.y;
					const uint32_t bit = 31u - (uint32_t)__clz((int)ng.y);
					const uint32_t base = ng.x;
					ng.y &= ~(1u << bit);
					if (ng.y > 0x00ffffffu) {
						if (sp < kSmemStack) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(lane_addr + (uint32_t)sp * 256u), "r"(ng.x), "r"(ng.y) : "memory");
						else if (sp < kSmemStack + kLocalStack) l_stack[sp - kSmemStack] = ng;
						++sp;
						if (STATS && (unsigned long long)sp > st_depth) st_depth = sp;
					}
					if (STATS) ++st_nodes;
					const uint32_t slot = (bit - 24u) ^ octinv;
					const uint32_t rel = (uint32_t)__popc(imask & ~(0xffffffffu << slot));
					if (WIDE128) {
						const uint8_t *np = reinterpret_cast<const uint8_t *>(nodes_base) + (size_t)(base + rel) * 128u;
						const Words8 a = ldg256(np), b = ldg256(np + 32), c = ldg256(np + 64), d = ldg256(np + 96);
						const float aix = __fmul_rn(__uint_as_float(a.v[4]), idx);
						const float aiy = __fmul_rn(__uint_as_float(a.v[5]), idy);
						const float aiz = __fmul_rn(__uint_as_float(a.v[6]), idz);
						const float aox = __fmul_rn(__fsub_rn(__uint_as_float(a.v[0]), ox), idx);
						const float aoy = __fmul_rn(__fsub_rn(__uint_as_float(a.v[1]), oy), idy);
						const float aoz = __fmul_rn(__fsub_rn(__uint_as_float(a.v[2]), oz), idz);
						ng.x = a.v[7];
						tg.x = d.v[4];
						const bool nx = idx < 0.0f, ny = idy < 0.0f, nz = idz < 0.0f;
						// planes: b = (lox.lo, lox.hi, loy.lo, loy.hi, loz.lo, loz.hi, hix.lo, hix.hi), c.v[0..3] = (hiy.lo, hiy.hi, hiz.lo, hiz.hi)
						const uint32_t lx0 = nx ? b.v[6] : b.v[0], ly0 = ny ? c.v[0] : b.v[2], lz0 = nz ? c.v[2] : b.v[4];
						const uint32_t hx0 = nx ? b.v[0] : b.v[6], hy0 = ny ? b.v[2] : c.v[0], hz0 = nz ? b.v[4] : c.v[2];
						const uint32_t lx1 = nx ? b.v[7] : b.v[1], ly1 = ny ? c.v[1] : b.v[3], lz1 = nz ? c.v[3] : b.v[5];
						const uint32_t hx1 = nx ? b.v[1] : b.v[7], hy1 = ny ? b.v[3] : c.v[1], hz1 = nz ? b.v[5] : c.v[3];
						uint32_t hitmask = test_child_pair_wide<0, CVT_PLANES>(c.v[4], c.v[5], lx0, ly0, lz0, hx0, hy0, hz0, aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic);
						hitmask |= test_child_pair_wide<2, CVT_PLANES>(c.v[6], c.v[7], lx0, ly0, lz0, hx0, hy0, hz0, aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic);
						hitmask |= test_child_pair_wide<0, CVT_PLANES>(d.v[0], d.v[1], lx1, ly1, lz1, hx1, hy1, hz1, aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic);
						hitmask |= test_child_pair_wide<2, CVT_PLANES>(d.v[2], d.v[3], lx1, ly1, lz1, hx1, hy1, hz1, aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic);
						hitmask = permute_inner_slots(hitmask, octinv);
						ng.y = (hitmask & 0xff000000u) | a.v[3];
						tg.y = hitmask & 0x00ffffffu;
					} else {
					uint4 n0, n1, n2, n3, n4;
					float scx, scy, scz; // 2^ex, 2^ey, 2^ez
					if (WIDE) {
						const uint8_t *np = reinterpret_cast<const uint8_t *>(nodes_base) + (size_t)(base + rel) * 96u;
						const Words8 a = ldg256(np), b = ldg256(np + 32), c = ldg256(np + 64);
						n0 = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]); n1 = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
						n2 = make_uint4(b.v[0], b.v[1], b.v[2], b.v[3]); n3 = make_uint4(b.v[4], b.v[5], b.v[6], b.v[7]);
						n4 = make_uint4(c.v[0], c.v[1], c.v[2], c.v[3]);
						scx = __uint_as_float(c.v[4]); scy = __uint_as_float(c.v[5]); scz = __uint_as_float(c.v[6]);
					} else {
						const uint4 *np = nodes_base + (size_t)(base + rel) * 5u;
						n0 = __ldg(np); n1 = __ldg(np + 1); n2 = __ldg(np + 2); n3 = __ldg(np + 3); n4 = __ldg(np + 4);
						scx = __uint_as_float((n0.w & 0xffu) << 23);
						scy = __uint_as_float(((n0.w >> 8) & 0xffu) << 23);
						scz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23);
					}
					const float aix = __fmul_rn(scx, idx);
					const float aiy = __fmul_rn(scy, idy);
					const float aiz = __fmul_rn(scz, idz);
					const float aox = __fmul_rn(__fsub_rn(__uint_as_float(n0.x), ox), idx);
					const float aoy = __fmul_rn(__fsub_rn(__uint_as_float(n0.y), oy), idy);
					const float aoz = __fmul_rn(__fsub_rn(__uint_as_float(n0.z), oz), idz);

					ng.x = n1.x;
					tg.x = n1.y;
					const uint32_t octinv4 = octinv * 0x01010101u;
					const bool nx = idx < 0.0f, ny = idy < 0.0f, nz = idz < 0.0f;
					// planes: n2 = (lox.lo, lox.hi, loy.lo, loy.hi) n3 = (loz.lo, loz.hi, hix.lo, hix.hi)
					//         n4 = (hiy.lo, hiy.hi, hiz.lo, hiz.hi)
					uint32_t hitmask;
					if (PACKED) {
						hitmask = test_children4_packed<CVT_PLANES>(n1.z, octinv4,
							nx ? n3.z : n2.x, ny ? n4.x : n2.z, nz ? n4.z : n3.x,
							nx ? n2.x : n3.z, ny ? n2.z : n4.x, nz ? n3.x : n4.z,
							aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic);
						hitmask |= test_children4_packed<CVT_PLANES>(n1.w, octinv4,
							nx ? n3.w : n2.y, ny ? n4.y : n2.w, nz ? n4.w : n3.y,
							nx ? n2.y : n3.w, ny ? n2.w : n4.y, nz ? n3.y : n4.w,
							aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic);
					} else {
					hitmask = test_children4<CVT_PLANES, HM_LUT>(n1.z, octinv4,
						nx ? n3.z : n2.x, ny ? n4.x : n2.z, nz ? n4.z : n3.x,
						nx ? n2.x : n3.z, ny ? n2.z : n4.x, nz ? n3.x : n4.z,
						aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic, s_hm_lut);
					hitmask |= test_children4<CVT_PLANES, HM_LUT>(n1.w, octinv4,
						nx ? n3.w : n2.y, ny ? n4.y : n2.w, nz ? n4.w : n3.y,
						nx ? n2.y : n3.w, ny ? n2.w : n4.y, nz ? n3.y : n4.w,
						aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic, s_hm_lut);
					}
					ng.y = (hitmask & 0xff000000u) | (n0.w >> 24);
					tg.y = hitmask & 0x00ffffffu;
					}
				}
				// The GLSL's else branch (:207-211, "G is a triangle group": tg = ng, ng = 0) cannot be reached: ng.y is
				// above 0x00ffffff at ray start and after every pop (only groups with inner hits are pushed), and a
				// group that loses its last inner hit is replaced at the bottom of the same round.

				if (rep + 1 < NODE_REPS && !finished && tg.y == 0u && ng.y <= 0x00ffffffu) { // pop between node steps
					if (sp == 0) { finished = true; ng.y = 0u; }
					else {
						--sp;
						if (sp < kSmemStack) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ng.x), "=r"(ng.y) : "r"(lane_addr + (uint32_t)sp * 256u) : "memory");
						else ng = l_stack[(sp - kSmemStack) < kLocalStack ? (sp - kSmemStack) : (kLocalStack - 1)];
					}
				}
				}
				// Woop test of leaf reference TR with rows M0..M2 (:221-241)
#define ADYPT_WOOP_TEST(TR, M0, M1, M2) \
	do { \
		if (STATS) ++st_tris; \
		float tt, tu, tv; \
		woop_eval<PACKED>(ox, oy, oz, dx, dy, dz, M0, M1, M2, tt, tu, tv); \
		if (tt > tmin && tt < hit_t && tu >= 0.0f && tv >= 0.0f && __fadd_rn(tu, tv) <= 1.0f) { /* u <= 1 is implied, see below */ \
			hit_t = tt; \
			if (ANY) finished = true; /* :480-483 */ \
			else { hit_u = tu; hit_v = tv; hit_idx = (int32_t)(TR); } \
		} \
	} while (0)
#define ADYPT_WOOP_EVAL(M0, M1, M2, TT, TU, TV) woop_eval<PACKED>(ox, oy, oz, dx, dy, dz, M0, M1, M2, TT, TU, TV)
	// The reference also tests u <= 1 (traversal.glsl:235). It is implied by the others for every input -- v >= 0 makes u + v >= u, rounding is
	// monotonic, so u <= fl(u + v) <= 1, and a NaN u already fails u >= 0 -- so the acceptance below leaves it out (same results, one compare less).
#define ADYPT_WOOP_ACCEPT(TR, TT, TU, TV) \
	do { \
		if (STATS) ++st_tris; \
		if (TT > tmin && TT < hit_t && TU >= 0.0f && TV >= 0.0f && __fadd_rn(TU, TV) <= 1.0f) { \
			hit_t = TT; \
			if (ANY) finished = true; /* :480-483 */ \
			else { hit_u = TU; hit_v = TV; hit_idx = (int32_t)(TR); } \
		} \
	} while (0)
				if (TRI_BATCH == 12) {
					// batch of two with both fetches in flight before the first test
					if (tg.y != 0u) {
						const uint32_t tr0 = tg.x + (uint32_t)(__ffs((int)tg.y) - 1);
						tg.y &= tg.y - 1u;
						const bool two = tg.y != 0u;
						const uint32_t tr1 = tg.x + (uint32_t)(__ffs((int)tg.y) - 1);
						tg.y &= tg.y - 1u; // no-op on 0
						float4 a0, a1, a2, b0, b1, b2;
						if (WOOP64) {
							const float4 *wa = woop_base + (size_t)tr0 * 4u;
							const float4 *wb = woop_base + (size_t)(two ? tr1 : tr0) * 4u;
							const Words8 pa = ldg256(wa), qa = ldg256(wa + 2), pb = ldg256(wb), qb = ldg256(wb + 2);
							a0 = make_float4(__uint_as_float(pa.v[0]), __uint_as_float(pa.v[1]), __uint_as_float(pa.v[2]), __uint_as_float(pa.v[3]));
							a1 = make_float4(__uint_as_float(pa.v[4]), __uint_as_float(pa.v[5]), __uint_as_float(pa.v[6]), __uint_as_float(pa.v[7]));
							a2 = make_float4(__uint_as_float(qa.v[0]), __uint_as_float(qa.v[1]), __uint_as_float(qa.v[2]), __uint_as_float(qa.v[3]));
							b0 = make_float4(__uint_as_float(pb.v[0]), __uint_as_float(pb.v[1]), __uint_as_float(pb.v[2]), __uint_as_float(pb.v[3]));
							b1 = make_float4(__uint_as_float(pb.v[4]), __uint_as_float(pb.v[5]), __uint_as_float(pb.v[6]), __uint_as_float(pb.v[7]));
							b2 = make_float4(__uint_as_float(qb.v[0]), __uint_as_float(qb.v[1]), __uint_as_float(qb.v[2]), __uint_as_float(qb.v[3]));
						} else {
							const float4 *wa = woop_base + (size_t)tr0 * 3u;
							const float4 *wb = woop_base + (size_t)(two ? tr1 : tr0) * 3u;
							a0 = __ldg(wa); a1 = __ldg(wa + 1); a2 = __ldg(wa + 2);
							b0 = __ldg(wb); b1 = __ldg(wb + 1); b2 = __ldg(wb + 2);
						}
						if (!ANY) {
							// both triangles' t, u, v first (they do not depend on hit_t, so ptxas can interleave the two chains;
							// a lane with one triangle evaluates it twice, wb == wa), then the acceptance tests in the reference's
							// order. -1 % time; any-hit rays mostly stop at their first triangle and keep the sequential form.
							float tt0, tu0, tv0, tt1, tu1, tv1;
							ADYPT_WOOP_EVAL(a0, a1, a2, tt0, tu0, tv0);
							ADYPT_WOOP_EVAL(b0, b1, b2, tt1, tu1, tv1);
							ADYPT_WOOP_ACCEPT(tr0, tt0, tu0, tv0);
							if (two) ADYPT_WOOP_ACCEPT(tr1, tt1, tu1, tv1);
						} else {
							ADYPT_WOOP_TEST(tr0, a0, a1, a2);
							if (two && !finished) ADYPT_WOOP_TEST(tr1, b0, b1, b2);
						}
					}
				} else
				for (int batch = 0; tg.y != 0u && (TRI_BATCH == 0 || batch < TRI_BATCH); ++batch) { // :213-243
					const uint32_t tr = tg.x + (uint32_t)(__ffs((int)tg.y) - 1);
					tg.y &= tg.y - 1u;
					const float4 *wp = woop_base + (size_t)tr * 3u;
					const float4 m0 = __ldg(wp), m1 = __ldg(wp + 1), m2 = __ldg(wp + 2);
					ADYPT_WOOP_TEST(tr, m0, m1, m2);
					if (ANY && finished) break;
				}
#undef ADYPT_WOOP_TEST
#undef ADYPT_WOOP_EVAL
#undef ADYPT_WOOP_ACCEPT

				if (!finished && (TRI_BATCH == 0 || tg.y == 0u) && ng.y <= 0x00ffffffu) { // :245-250
					if (sp == 0) finished = true;
					else {
						--sp;
						if (sp < kSmemStack) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ng.x), "=r"(ng.y) : "r"(lane_addr + (uint32_t)sp * 256u) : "memory");
						else ng = l_stack[(sp - kSmemStack) < kLocalStack ? (sp - kSmemStack) : (kLocalStack - 1)];
					}
				}

				if (finished) sp = kUnsaved;
			}
			busy = __ballot_sync(kFullMask, sp >= 0);
		} while (busy != 0u && (exhausted || __popc(busy) >= p.refill_threshold));
	}
	if (STATS) {
		atomicAdd(p.stats + 0, st_nodes);
		atomicAdd(p.stats + 1, st_tris);
		atomicAdd(p.stats + 2, st_hits);
		atomicMax(p.stats + 3, st_depth);
		if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.stats + 4, n_rays);
	}
}

} // namespace adypt
