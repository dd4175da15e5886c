// EXPERIMENT (round 2): CWBVH traversal with the rays of a CTA held in SHARED MEMORY and bound to lanes one round
// at a time -- "moving rays instead of work". The product kernel (traverse.cuh) keeps a ray in its lane for life, so every warp
// runs its node step at ~28/32 lanes and its triangle section at ~6.5/32. Here a warp's round is: claim up to 32 ray slots from
// one of two queues, run ONE kind of work for all of them, write the state back, and queue every ray for what it needs next:
//   node round    : the rays that need a node step          (traversal.glsl:47-205)
//   triangle round: the rays that have triangles to test    (traversal.glsl:213-243; two tests per round, like TRI_BATCH 12)
//   set-up round  : 32 new rays into 32 free slots          (traversal.glsl:16-35)
// so that, away from the tail of a batch, every round runs at 32/32 lanes. A ray's own sequence of node steps and triangle
// tests -- and with it every result -- is exactly the product kernel's; only which lane executes a step differs.
//
// Shared memory per CTA (256 threads, S = 320 slots): ray state as structure of arrays (88 B per ray), an 8-entry traversal
// stack per ray, three rings of slot numbers (needs-node, needs-triangles, free) with reserve-then-publish semantics
// (tail is bumped with one atomic per warp, each entry carries a valid bit the consumer waits for and clears).
#pragma once
#include "traverse.cuh"

namespace adypt {

constexpr int kPoolBlock = 256;        // threads per CTA
constexpr int kPoolSlots = 320;        // rays resident per CTA
constexpr int kPoolStack = 8;          // stack entries per ray in shared memory (deeper rays: not supported by the experiment, traps)
constexpr unsigned kPoolRing = 512;    // ring capacity (power of two >= kPoolSlots)

struct PoolShared {
	float4 o[kPoolSlots];              // origin, tmin
	float4 d[kPoolSlots];              // direction, octinv bits
	float4 i[kPoolSlots];              // 1 / direction, low half of the ray index
	float4 h[kPoolSlots];              // hit_t, hit_u, hit_v, hit_idx bits
	uint4 g[kPoolSlots];               // node group (x, y), triangle group (z, w)
	uint2 m[kPoolSlots];               // stack depth, high half of the ray index
	uint2 stack[kPoolStack][kPoolSlots];
	unsigned short ring[3][kPoolRing]; // 0: needs a node step, 1: needs triangle tests, 2: free slots; entry = slot | 0x8000
	unsigned head[3], tail[3];
	int live;                          // rays resident in slots
};

enum { kRingNode = 0, kRingTri = 1, kRingFree = 2 };

// lane 0: claims up to `want` entries of ring r; returns how many and the first position
__device__ __forceinline__ unsigned pool_claim(PoolShared *S, int r, unsigned want, unsigned *base)
{
	volatile unsigned *head = &S->head[r], *tail = &S->tail[r];
	for (;;) {
		const unsigned hd = *head, tl = *tail;
		const unsigned avail = tl - hd;
		const unsigned k = avail < want ? avail : want;
		if (k == 0u) return 0u;
		if (atomicCAS(&S->head[r], hd, hd + k) == hd) {
			*base = hd;
			return k;
		}
	}
}

// the slot number at position pos of ring r (waits for the producer to publish it, then clears the entry)
__device__ __forceinline__ unsigned pool_take(PoolShared *S, int r, unsigned pos)
{
	volatile unsigned short *e = &S->ring[r][pos & (kPoolRing - 1u)];
	unsigned v;
	while (((v = *e) & 0x8000u) == 0u) {}
	*e = 0;
	__threadfence_block();
	return v & 0x7fffu;
}

// every lane with `want` appends its slot to ring r (one atomic per warp)
__device__ __forceinline__ void pool_push(PoolShared *S, int r, bool want, unsigned slot, unsigned lane)
{
	const unsigned m = __ballot_sync(kFullMask, want);
	if (m == 0u) return;
	const unsigned leader = (unsigned)__ffs((int)m) - 1u;
	unsigned base = 0;
	if (lane == leader) base = atomicAdd(&S->tail[r], (unsigned)__popc(m));
	base = __shfl_sync(kFullMask, base, (int)leader);
	if (want) {
		const unsigned pos = base + (unsigned)__popc(m & ((1u << lane) - 1u));
		volatile unsigned short *e = &S->ring[r][pos & (kPoolRing - 1u)];
		while (*e != 0) {}       // the previous lap's consumer has taken it
		__threadfence_block();   // the ray's state is visible before the entry is
		*e = (unsigned short)(slot | 0x8000u);
	}
}

template <bool ANY, int CVT_PLANES = 4>
__global__ void __launch_bounds__(kPoolBlock, 3) trace_pool_kernel(const TraceParams p)
{
	extern __shared__ __align__(16) unsigned char pool_raw[];
	PoolShared *S = reinterpret_cast<PoolShared *>(pool_raw);
	const uint32_t magic = p.magic;
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lt_mask = (1u << lane) - 1u;
	const unsigned long long n_rays = p.n_ptr ? *p.n_ptr : p.n;

	for (unsigned k = threadIdx.x; k < 3u * kPoolRing; k += kPoolBlock) (&S->ring[0][0])[k] = 0;
	if (threadIdx.x < 3) S->head[threadIdx.x] = 0u;
	if (threadIdx.x < 2) S->tail[threadIdx.x] = 0u;
	if (threadIdx.x == 0) {
		S->tail[kRingFree] = (unsigned)kPoolSlots;
		S->live = 0;
	}
	__syncthreads();
	for (unsigned k = threadIdx.x; k < (unsigned)kPoolSlots; k += kPoolBlock) S->ring[kRingFree][k] = (unsigned short)(k | 0x8000u);
	__syncthreads();

	unsigned long long pool_next = 0, pool_end = 0; // warp-uniform pool of ray indices (guided self-scheduling as in traverse.cuh)
	bool exhausted = false;
	unsigned idle_spins = 0;

	for (;;) {
		// ------------------------------------------------------------- what does this warp do next?
		int action = 0; // 1 triangle round, 2 set-up round, 3 node round, 4 exit
		unsigned count = 0, base = 0;
		if (lane == 0) {
			volatile unsigned *hd = S->head, *tl = S->tail;
			const unsigned tri_avail = tl[kRingTri] - hd[kRingTri], node_avail = tl[kRingNode] - hd[kRingNode], free_avail = tl[kRingFree] - hd[kRingFree];
			const bool patient = idle_spins < 8u; // wait a little for a full round before running a partial one
			if (tri_avail >= 32u) action = 1;
			else if (!exhausted && free_avail >= 32u) action = 2;
			else if (node_avail >= 32u) action = 3;
			else if (!patient && tri_avail > 0u) action = 1;
			else if (!patient && node_avail > 0u) action = 3;
			else if (!patient && !exhausted && free_avail > 0u) action = 2;
			else if (exhausted && *(volatile int *)&S->live == 0) action = 4;
			if (action == 1) count = pool_claim(S, kRingTri, 32u, &base);
			else if (action == 3) count = pool_claim(S, kRingNode, 32u, &base);
		}
		action = __shfl_sync(kFullMask, action, 0);
		count = __shfl_sync(kFullMask, count, 0);
		base = __shfl_sync(kFullMask, base, 0);
		if (action == 4) break;
		if (action == 0 || ((action == 1 || action == 3) && count == 0u)) {
			++idle_spins;
			__nanosleep(64);
			continue;
		}
		idle_spins = 0;

		if (action == 2) {
			// --------------------------------------------------------- set-up round (traversal.glsl:16-35)
			if (pool_next >= pool_end) {
				unsigned long long b = 0;
				const uint32_t request = next_chunk(n_rays - pool_end, p.guided_shift, p.pool_chunk);
				if (lane == 0) b = atomicAdd(p.counter, (unsigned long long)request);
				b = __shfl_sync(kFullMask, b, 0);
				if (b >= n_rays) {
					exhausted = true;
					continue;
				}
				pool_next = b;
				pool_end = (b + request < n_rays) ? b + request : n_rays;
			}
			const unsigned long long left = pool_end - pool_next;
			unsigned want = left < 32ull ? (unsigned)left : 32u;
			if (lane == 0) count = pool_claim(S, kRingFree, want, &base);
			count = __shfl_sync(kFullMask, count, 0);
			base = __shfl_sync(kFullMask, base, 0);
			if (count == 0u) continue;
			const bool mine = lane < count;
			unsigned slot = 0;
			if (mine) {
				slot = pool_take(S, kRingFree, base + lane);
				const unsigned long long r = pool_next + lane;
				const float4 r0 = __ldg(p.rays + r * p.ray_stride), r1 = __ldg(p.dirs + r * p.ray_stride);
				const float ooeps = 5.42101086242752217e-20f; // exp2(-64)
				float sx = fabsf(r1.x) > ooeps ? r1.x : (r1.x >= 0.0f ? ooeps : -ooeps);
				float sy = fabsf(r1.y) > ooeps ? r1.y : (r1.y >= 0.0f ? ooeps : -ooeps);
				float sz = fabsf(r1.z) > ooeps ? r1.z : (r1.z >= 0.0f ? ooeps : -ooeps);
				const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy)), __fmul_rn(sz, sz));
				const float inv = __frcp_rn(__fsqrt_rn(len2));
				sx = __fmul_rn(sx, inv); sy = __fmul_rn(sy, inv); sz = __fmul_rn(sz, inv);
				const uint32_t oi = 7u - ((sx < 0.0f ? 1u : 0u) | (sy < 0.0f ? 2u : 0u) | (sz < 0.0f ? 4u : 0u));
				S->o[slot] = r0;
				S->d[slot] = make_float4(sx, sy, sz, __uint_as_float(oi));
				S->i[slot] = make_float4(__frcp_rn(sx), __frcp_rn(sy), __frcp_rn(sz), __uint_as_float((uint32_t)r));
				S->h[slot] = make_float4(1e9f, 0.0f, 0.0f, __int_as_float(-1));
				S->g[slot] = make_uint4(0u, 0x80000000u, 0u, 0u);
				S->m[slot] = make_uint2(0u, (uint32_t)(r >> 32));
			}
			pool_next += count;
			if (lane == 0) atomicAdd(&S->live, (int)count);
			pool_push(S, kRingNode, mine, slot, lane);
			continue;
		}

		const bool mine = lane < count;
		unsigned slot = 0;
		int dest = -1; // ring the ray goes to next; 3 = finished
		if (action == 3) {
			// --------------------------------------------------------- node round (traversal.glsl:47-205)
			if (mine) {
				slot = pool_take(S, kRingNode, base + lane);
				const float4 O = S->o[slot], D = S->d[slot], I = S->i[slot];
				uint4 G = S->g[slot];
				const float hit_t = S->h[slot].x;
				unsigned sp = S->m[slot].x;
				const float ox = O.x, oy = O.y, oz = O.z, tmin = O.w, idx = I.x, idy = I.y, idz = I.z;
				const uint32_t octinv = __float_as_uint(D.w);
				uint2 ng = make_uint2(G.x, G.y), tg;
				// n <- closest child of G (:50-67)
				const uint32_t imask = ng.y;
				const uint32_t bit = 31u - (uint32_t)__clz((int)ng.y);
				const uint32_t nbase = ng.x;
				ng.y &= ~(1u << bit);
				if (ng.y > 0x00ffffffu) {
					if (sp >= (unsigned)kPoolStack) __trap(); // experiment: no deep-stack path
					S->stack[sp][slot] = ng;
					++sp;
				}
				const uint32_t cslot = (bit - 24u) ^ octinv;
				const uint32_t rel = (uint32_t)__popc(imask & ~(0xffffffffu << cslot));
				const uint4 *np = p.nodes + (size_t)(nbase + rel) * 5u;
				const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
				const float aix = __fmul_rn(__uint_as_float((n0.w & 0xffu) << 23), idx);
				const float aiy = __fmul_rn(__uint_as_float(((n0.w >> 8) & 0xffu) << 23), idy);
				const float aiz = __fmul_rn(__uint_as_float(((n0.w >> 16) & 0xffu) << 23), idz);
				const float aox = __fmul_rn(__fsub_rn(__uint_as_float(n0.x), ox), idx);
				const float aoy = __fmul_rn(__fsub_rn(__uint_as_float(n0.y), oy), idy);
				const float aoz = __fmul_rn(__fsub_rn(__uint_as_float(n0.z), oz), idz);
				ng.x = n1.x;
				tg.x = n1.y;
				const uint32_t octinv4 = octinv * 0x01010101u;
				const bool nx = idx < 0.0f, ny = idy < 0.0f, nz = idz < 0.0f;
				uint32_t hitmask = test_children4<CVT_PLANES, false>(n1.z, octinv4,
					nx ? n3.z : n2.x, ny ? n4.x : n2.z, nz ? n4.z : n3.x,
					nx ? n2.x : n3.z, ny ? n2.z : n4.x, nz ? n3.x : n4.z,
					aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic, nullptr);
				hitmask |= test_children4<CVT_PLANES, false>(n1.w, octinv4,
					nx ? n3.w : n2.y, ny ? n4.y : n2.w, nz ? n4.w : n3.y,
					nx ? n2.y : n3.w, ny ? n2.w : n4.y, nz ? n3.y : n4.w,
					aix, aiy, aiz, aox, aoy, aoz, tmin, hit_t, magic, nullptr);
				ng.y = (hitmask & 0xff000000u) | (n0.w >> 24);
				tg.y = hitmask & 0x00ffffffu;
				if (tg.y != 0u) dest = kRingTri;
				else if (ng.y > 0x00ffffffu) dest = kRingNode;
				else if (sp == 0u) dest = 3; // :245-250
				else {
					--sp;
					ng = S->stack[sp][slot];
					dest = kRingNode;
				}
				S->g[slot] = make_uint4(ng.x, ng.y, tg.x, tg.y);
				S->m[slot].x = sp;
			}
		} else {
			// --------------------------------------------------------- triangle round (traversal.glsl:213-243)
			if (mine) {
				slot = pool_take(S, kRingTri, base + lane);
				const float4 O = S->o[slot], D = S->d[slot];
				float4 H = S->h[slot];
				uint4 G = S->g[slot];
				const float ox = O.x, oy = O.y, oz = O.z, tmin = O.w, dx = D.x, dy = D.y, dz = D.z;
				float hit_t = H.x, hit_u = H.y, hit_v = H.z;
				int32_t hit_idx = __float_as_int(H.w);
				uint2 ng = make_uint2(G.x, G.y), tg = make_uint2(G.z, G.w);
				bool finished = false;
				const uint32_t tr0 = tg.x + (uint32_t)(__ffs((int)tg.y) - 1);
				tg.y &= tg.y - 1u;
				const bool two = tg.y != 0u;
				const uint32_t tr1 = tg.x + (uint32_t)(__ffs((int)tg.y) - 1);
				tg.y &= tg.y - 1u; // no-op on 0
				const float4 *wa = p.woop + (size_t)tr0 * 3u;
				const float4 *wb = p.woop + (size_t)(two ? tr1 : tr0) * 3u;
				const float4 a0 = __ldg(wa), a1 = __ldg(wa + 1), a2 = __ldg(wa + 2);
				const float4 b0 = __ldg(wb), b1 = __ldg(wb + 1), b2 = __ldg(wb + 2);
#define ADYPT_POOL_EVAL(M0, M1, M2, TT, TU, TV) \
	do { \
		const float toz = __fsub_rn(M0.w, dot3_fma(ox, oy, oz, M0)); \
		const float tidz = __frcp_rn(dot3_fma(dx, dy, dz, M0)); \
		TT = __fmul_rn(toz, tidz); \
		const float tox = __fadd_rn(M1.w, dot3_fma(ox, oy, oz, M1)); \
		TU = __fmaf_rn(TT, dot3_fma(dx, dy, dz, M1), tox); \
		const float toy = __fadd_rn(M2.w, dot3_fma(ox, oy, oz, M2)); \
		TV = __fmaf_rn(TT, dot3_fma(dx, dy, dz, M2), toy); \
	} while (0)
#define ADYPT_POOL_ACCEPT(TR, TT, TU, TV) \
	do { \
		if (TT > tmin && TT < hit_t && TU >= 0.0f && TU <= 1.0f && TV >= 0.0f && __fadd_rn(TU, TV) <= 1.0f) { \
			hit_t = TT; \
			if (ANY) finished = true; /* :480-483 */ \
			else { hit_u = TU; hit_v = TV; hit_idx = (int32_t)(TR); } \
		} \
	} while (0)
				float tt0, tu0, tv0, tt1, tu1, tv1;
				ADYPT_POOL_EVAL(a0, a1, a2, tt0, tu0, tv0);
				ADYPT_POOL_EVAL(b0, b1, b2, tt1, tu1, tv1);
				ADYPT_POOL_ACCEPT(tr0, tt0, tu0, tv0);
				if (two && !finished) ADYPT_POOL_ACCEPT(tr1, tt1, tu1, tv1);
#undef ADYPT_POOL_EVAL
#undef ADYPT_POOL_ACCEPT
				if (finished) dest = 3;
				else if (tg.y != 0u) dest = kRingTri;
				else if (ng.y > 0x00ffffffu) dest = kRingNode;
				else {
					unsigned sp = S->m[slot].x;
					if (sp == 0u) dest = 3;
					else {
						--sp;
						ng = S->stack[sp][slot];
						S->m[slot].x = sp;
						dest = kRingNode;
					}
				}
				S->h[slot] = make_float4(hit_t, hit_u, hit_v, __int_as_float(hit_idx));
				S->g[slot] = make_uint4(ng.x, ng.y, tg.x, tg.y);
			}
		}
		// ------------------------------------------------------------- results of finished rays, then everybody moves on
		const unsigned fin = __ballot_sync(kFullMask, dest == 3);
		if (dest == 3) {
			const float4 H = S->h[slot];
			const unsigned long long ray_idx = ((unsigned long long)S->m[slot].y << 32) | (unsigned long long)__float_as_uint(S->i[slot].w);
			const int32_t hit_idx = __float_as_int(H.w);
			if (ANY) {
				p.out_occ[ray_idx] = (H.x < 1e9f) ? 1 : 0;
			} else {
				p.out_tri[ray_idx] = hit_idx >= 0 ? __ldg(p.tri_indices + hit_idx) : -1; // :253-254
				if (p.out_t) p.out_t[ray_idx] = H.x;
				if (p.out_uv) p.out_uv[ray_idx] = make_float2(H.y, H.z);
			}
		}
		pool_push(S, kRingNode, dest == kRingNode, slot, lane);
		pool_push(S, kRingTri, dest == kRingTri, slot, lane);
		pool_push(S, kRingFree, dest == 3, slot, lane);
		if (fin != 0u && lane == 0) atomicSub(&S->live, __popc(fin));
		(void)lt_mask;
	}
}

} // namespace adypt
