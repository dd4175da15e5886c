// Wavefront path tracer + AOV viewer: the CUDA stand-in for OglPathTracer (src/Tracer/OglPathTracer.*)
// and for the two GLSL programs it dispatches (shaders/primaryray.glsl, shaders/pathtracer.glsl).
//
// The reference runs one megakernel thread per pixel per sample (pathtracer.glsl:220-227). Here one
// "batch" is S consecutive samples of every pixel (S <= tmpLifetime, all sharing one primary hit and one
// sub-pixel stratum, pathtracer.glsl:113-127,206-211) and the integrator is split into queue-driven
// stages so that every traversal launch sees a dense array of live rays:
//   generate  : camera rays for the batch's stratum                      (Camera(), SubPixel())
//   extend    : closest-hit traversal of a ray queue (trace_kernel, traverse.cuh)
//   shade     : FetchInfo + material switch + next-ray sampling, appends survivors to the next queue
//   accumulate: clamp + running mean in sample order                      (main(), :224-226)
// Queue lengths stay on the device (kernels read them through pointers), so a whole batch is enqueued
// without a host round trip. Per-path results are independent of queue order, and the accumulate stage
// applies samples in ascending spp, so images are bit-reproducible run to run and identical to
// dispatching the reference's megakernel once per sample.
//
// FP policy: every GLSL expression is evaluated un-fused, left to right in IEEE fp32
// (the library is built with -fmad=false), sqrt and division are IEEE, and sin/cos/pow are the deterministic
// double-precision recipes of detmath.cuh -- so images match the CPU oracle bit for bit.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "detmath.cuh"
#include "hostmath.h"
#include "scene.h"
#include "traverse.cuh"

namespace adypt {

struct CameraArgs { // uuCamera (pathtracer.glsl:27-32), by value
	float origin[3], tmin;
	float inv_proj[16], inv_view[16];
};

struct PTArgs { // uuPT (pathtracer.glsl:34-38) + batch bookkeeping
	int32_t max_bounce, subpixel, tmp_lifetime;
	float clamp, sun[3];
	int32_t width, height;
	int32_t first_spp; // spp of the batch's first sample
	int32_t n_samples; // S
	int32_t dims;      // Sobol dimensions per sample: 2*max_bounce, +max_bounce roulette draws when rr_start >= 0
	int32_t rr_start;  // >= 0: Russian roulette from this bounce on (opt-in extension, -1 = the reference's behaviour)
	uint32_t zero;     // always 0, but only known at run time: lets a kernel tie an instruction's operand to a value it must wait for
	uint32_t early_slots; // tuning (ADYPT_EARLY_SLOTS): entries whose class always goes on ask for their queue slot before the gathers
};

// ---------------------------------------------------------------------------------------------
// GLSL-shaped helpers
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 x, V3 y) { return v3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
__device__ __forceinline__ V3 normalize(V3 v)
{
	const float inv = __fdiv_rn(1.0f, __fsqrt_rn(dot(v, v)));
	return v * inv;
}
__device__ __forceinline__ V3 reflect(V3 I, V3 N) { return I - N * (2.0f * dot(N, I)); }
__device__ __forceinline__ float glsl_min(float x, float y) { return y < x ? y : x; }
__device__ __forceinline__ float glsl_max(float x, float y) { return x < y ? y : x; }
__device__ __forceinline__ float fract(float x) { return x - floorf(x); }

// Camera(bias): primaryray.glsl:39-44 (bias 0) / pathtracer.glsl:213-218
__device__ __forceinline__ V3 camera_dir(const CameraArgs &c, int w, int h, int px, int py, float bx, float by)
{
	const float sx = 2.0f * ((float)px + bx) / (float)w - 1.0f;
	const float sy = -(2.0f * ((float)py + by) / (float)h - 1.0f);
	const float *ip = c.inv_proj, *iv = c.inv_view;
	const float vx = ip[0] * sx + ip[4] * sy + ip[8] * 1.0f + ip[12] * 1.0f;
	const float vy = ip[1] * sx + ip[5] * sy + ip[9] * 1.0f + ip[13] * 1.0f;
	const float vz = ip[2] * sx + ip[6] * sy + ip[10] * 1.0f + ip[14] * 1.0f;
	return normalize(v3(iv[0] * vx + iv[4] * vy + iv[8] * vz, iv[1] * vx + iv[5] * vy + iv[9] * vz, iv[2] * vx + iv[6] * vy + iv[10] * vz));
}

__device__ __forceinline__ void subpixel_bias(int subpixel, int tmp_lifetime, int spp, float *bx, float *by) // SubPixel(), :206-211
{
	const int idx = (spp / tmp_lifetime) % (subpixel * subpixel);
	const float unit = 1.0f / (float)subpixel;
	*bx = (float)(idx / subpixel) * unit;
	*by = (float)(idx % subpixel) * unit;
}

__global__ void k_generate(CameraArgs cam, int w, int h, float bx, float by, float4 *__restrict__ rays)
{
	const unsigned npix = (unsigned)w * (unsigned)h;
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
		const V3 d = camera_dir(cam, w, h, (int)(i % (unsigned)w), (int)(i / (unsigned)w), bx, by);
		rays[2 * (size_t)i] = make_float4(cam.origin[0], cam.origin[1], cam.origin[2], cam.tmin);
		rays[2 * (size_t)i + 1] = make_float4(d.x, d.y, d.z, 0.0f);
	}
}

__global__ void k_unorm8_table(float *__restrict__ out)
{
	out[threadIdx.x] = (float)threadIdx.x / 255.0f; // the same IEEE division the shading code would do per segment
}

// ---------------------------------------------------------------------------------------------
// Sobol: uSobol for every sample of the batch, computed on the device from the direction numbers
// (Sobol.cpp:16-21 in closed form: state after n calls = XOR of columns selected by gray(n)).
__global__ void k_sobol(const uint32_t *__restrict__ dirs, int dims, int first_spp, int n_samples, float *__restrict__ out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= dims * n_samples) return;
	const int s = i / dims, j = i % dims;
	const uint32_t n = (uint32_t)(first_spp + s) + 1u;
	uint32_t gray = n ^ (n >> 1), x = 0;
	for (int k = 0; gray; gray >>= 1, ++k)
		if (gray & 1u) x ^= dirs[j * 32 + k];
	out[i] = (float)((double)x / 4294967296.0);
}

// ---------------------------------------------------------------------------------------------
// shading

struct ShadeBuffers {
	const float4 *__restrict__ shade;        // 8 x float4 per scene triangle: the Triangle record on a 128-byte line (scene.cu)
	const uint32_t *__restrict__ tri_class;  // per triangle: shading branch of its material (kClass*) << 24 | material id, for the regrouping pass
	const Material *__restrict__ mats;
	const uchar4 *__restrict__ texels;       // diffuse textures, RGBX8 (nullptr when TEXTURE_COUNT == 0)
	const int4 *__restrict__ tex_table;      // (first texel, width, height, -)
	const uchar2 *__restrict__ bias;         // uSobolBiasImg
	const float *__restrict__ unorm8;        // the 256 values (float)b / 255.0f an RG8 UNORM texel can take, made by that very division on the device
	const float *__restrict__ sobol;         // [S][dims]
	// primary hit cache (uPrimaryTmpImg): x = tri id bits, zw = uv
	const int32_t *__restrict__ prim_tri;
	const float2 *__restrict__ prim_uv;
	// current queue (bounce >= 1), structure of arrays: ray origins (+ tmin) and directions (.w = path id bits) as the traversal
	// kernel takes them, the hits it wrote, and per entry the path's throughput + its pixel's bias bytes and sample number
	// (xyz = colour, w = bias.x | bias.y << 8 | sample in the batch << 16). Shading reads the direction but not the origin, so the two are kept apart.
	const float4 *__restrict__ in_org;
	const float4 *__restrict__ in_dir;
	const int32_t *__restrict__ in_tri;
	const float2 *__restrict__ in_uv;
	const float4 *__restrict__ in_state;
	const unsigned long long *in_count;
	// next queue
	float4 *__restrict__ out_org;
	float4 *__restrict__ out_dir;
	float4 *__restrict__ out_state;
	unsigned long long *out_count;
	// per-path radiance so far (final once the path ends); written by bounce 0, then only touched when something is added
	float4 *__restrict__ ret;
	unsigned long long *segments;            // statistics
	// connect stage (optional): shadow rays towards a fixed sun direction for paths that left the scene
	float4 *__restrict__ conn_rays;          // nullptr when the stage is off
	float4 *__restrict__ conn_color;         // what the path adds to ret if its shadow ray is unoccluded
	unsigned long long *conn_count;
	float sun_dir[3];
};

__device__ __forceinline__ V3 bary(const float *a, const float *b, const float *c, float u, float v) // :77-85
{
	const float w = 1.0f - u - v;
	return v3(a[0] * u + b[0] * v + c[0] * w, a[1] * u + b[1] * v + c[1] * w, a[2] * u + b[2] * v + c[2] * w);
}

// texture(sampler2D, st).rgb with GL_REPEAT, GL_LINEAR, one level (OglScene.cpp:33-38): the OpenGL 4.5 bilinear
// rule (spec 8.14.3) evaluated in fp32, left to right; texel = byte / 255
__device__ __forceinline__ float wrap_index(float i, float n)
{
	float m = i - n * floorf(i / n);
	return (m >= n || !(m >= 0.0f)) ? 0.0f : m;
}
__device__ __forceinline__ V3 texel_rgb(const uchar4 *__restrict__ base, int w, float i, float j)
{
	const uchar4 c = base[(int)j * w + (int)i];
	return v3((float)c.x / 255.0f, (float)c.y / 255.0f, (float)c.z / 255.0f);
}
__device__ __forceinline__ V3 sample_texture(const uchar4 *__restrict__ texels, const int4 t, float s, float tt)
{
	const float fw = (float)t.y, fh = (float)t.z;
	const float u = s * fw - 0.5f, v = tt * fh - 0.5f;
	const float fu = floorf(u), fv = floorf(v);
	const float a = u - fu, b = v - fv;
	const float i0 = wrap_index(fu, fw), i1 = wrap_index(fu + 1.0f, fw), j0 = wrap_index(fv, fh), j1 = wrap_index(fv + 1.0f, fh);
	const uchar4 *base = texels + t.x;
	const V3 c00 = texel_rgb(base, t.y, i0, j0), c10 = texel_rgb(base, t.y, i1, j0), c01 = texel_rgb(base, t.y, i0, j1), c11 = texel_rgb(base, t.y, i1, j1);
	const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
	return c00 * w00 + c10 * w10 + c01 * w01 + c11 * w11;
}

__device__ __forceinline__ V3 sample_hemisphere(float rx, float ry, float e) // :52-64
{
	rx *= 6.28318530718f;
	float cos_phi, sin_phi;
	detmath::sincos(rx, &sin_phi, &cos_phi);
	const float cos_theta = detmath::pow(1.0f - ry, 1.0f / (e + 1.0f));
	const float sin_theta = __fsqrt_rn(1.0f - cos_theta * cos_theta);
	return normalize(v3(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta));
}

__device__ __forceinline__ V3 align_direction(V3 dir, V3 target) // :66-71
{
	const V3 a = fabsf(target.x) > .01f ? v3(0.f, 1.f, 0.f) : v3(1.f, 0.f, 0.f);
	const V3 u = normalize(cross(a, target));
	const V3 v = cross(target, u);
	return u * dir.x + v * dir.y + target * dir.z;
}

// What FetchInfo (pathtracer.glsl:73-100) returns for a hit, plus the material fields the switch of Render needs
struct Surface {
	V3 normal, origin, emissive, diffuse, specular;
	int32_t illum;
	float shininess, ior;
};

// FetchInfo for scene triangle tri_idx at (u, v). The Triangle record is read from its 128-byte line: floats 0..24 are the
// reference's Triangle (Shape.hpp:70-88: p1 p2 p3, n1 n2 n3, tc1 tc2 tc3, matid), so the arithmetic is unchanged.
// matid >= 0: the caller already knows the triangle's material (from the class table), so the material is requested alongside the
// record; otherwise it is taken from the record (float 24), one dependent load later.
__device__ __forceinline__ void fetch_surface(const ShadeBuffers &B, int32_t tri_idx, float u, float v, Surface &s, int32_t matid = -1)
{
	const float4 *rec = B.shade + (size_t)tri_idx * 8u;
	const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2), r3 = __ldg(rec + 3), r4 = __ldg(rec + 4);
	const float t[18] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w, r4.x, r4.y};
	if (matid < 0) matid = __float_as_int(__ldg(rec + 6).x);
	const Material m = B.mats[matid];
	s.normal = normalize(bary(t + 9, t + 12, t + 15, u, v));
	s.origin = bary(t, t + 3, t + 6, u, v); // :138
	s.emissive = v3(m.er, m.eg, m.eb);
	s.specular = v3(m.sr, m.sg, m.sb);
	if (B.texels != nullptr && m.dtex != -1) { // :87-98 / primaryray.glsl:59-71
		const float4 r5 = __ldg(rec + 5);
		const float w = 1.0f - u - v;
		const float ss = r4.z * u + r5.x * v + r5.z * w, tt = r4.w * u + r5.y * v + r5.w * w;
		s.diffuse = sample_texture(B.texels, B.tex_table[m.dtex], ss, tt);
	} else
		s.diffuse = v3(m.dr, m.dg, m.db);
	s.illum = m.illum;
	s.shininess = m.shininess;
	s.ior = m.ior;
}

// The material switch of Render (pathtracer.glsl:141-201) for a segment that hit surface s coming along `dir` with
// throughput `color`: true = the path goes on along the new dir with the new color, false = it ends here.
__device__ __forceinline__ bool scatter(const PTArgs &A, int b, const Surface &s, float rx, float ry, float xi, V3 &dir, V3 &color)
{
	V3 normal = s.normal;
	if (s.illum < 6 && dot(dir, normal) > 0.0f) normal = -normal; // :141-142

	bool do_diffuse = false;
	switch (s.illum) { // :144-201
	case 2: {
		const float e = s.shininess * 0.01f;
		if (e > 0.3f) {
			const V3 r = reflect(dir, normal), h = sample_hemisphere(rx, ry, e);
			dir = align_direction(h, r);
			if (dot(dir, normal) < 0.0f) return false;
			color = color * (s.diffuse + s.specular * detmath::pow(dot(dir, r), e));
		} else
			do_diffuse = true;
		break;
	}
	case 1:
		do_diffuse = true;
		break;
	case 3: case 4: case 5:
		color = color * s.specular;
		dir = reflect(dir, normal);
		break;
	case 6: case 7: {
		float eta = s.ior;
		float cosi = dot(dir, normal);
		float fresnel, etai, etat;
		if (cosi > 0.0f) { etai = eta; etat = 1.0f; }
		else { etai = 1.0f; etat = eta; normal = -normal; cosi = -cosi; }
		eta = etai / etat;
		const float sint = etai / etat * __fsqrt_rn(glsl_max(0.f, 1.0f - cosi * cosi));
		if (sint >= 1.0f) fresnel = 1.0f;
		else {
			const float cost = __fsqrt_rn(glsl_max(0.f, 1.0f - sint * sint));
			const float Rs = ((etat * cosi) - (etai * cost)) / ((etat * cosi) + (etai * cost));
			const float Rp = ((etai * cosi) - (etat * cost)) / ((etai * cosi) + (etat * cost));
			fresnel = (Rs * Rs + Rp * Rp) * 0.5f;
		}
		const float cos2 = 1.0f - eta * eta * (1.0f - cosi * cosi);
		if (cos2 > 0.0f && rx >= fresnel) {
			dir = normalize(dir * eta + normal * (eta * cosi + __fsqrt_rn(cos2)));
			normal = -normal;
		} else
			dir = reflect(dir, normal);
		break;
	}
	default: break; // any other illum passes straight through
	}
	if (do_diffuse) { // :156-159
		dir = align_direction(sample_hemisphere(rx, ry, 0.0f), normal);
		color = color * s.diffuse;
	}
	// Russian roulette (BASELINE.json configs[2]; not in the reference, off unless adypt_tracer_set_russian_roulette
	// asked for it): survive with probability p = min(1, max(color)), then weight by 1/p. xi = this bounce's draw.
	if (A.rr_start >= 0 && b >= A.rr_start) {
		const float p = glsl_min(1.0f, glsl_max(color.x, glsl_max(color.y, color.z)));
		if (!(xi < p)) return false;
		color = v3(__fdiv_rn(color.x, p), __fdiv_rn(color.y, p), __fdiv_rn(color.z, p));
	}
	return true;
}

// x + a == x bit for bit when every component of a is +-0 (x is never -0 here: radiance sums start at +0 and (+0) + (-0) = +0),
// so such an addition -- a non-emissive hit, the common case -- needs no read-modify-write of the path's radiance. NaN compares unequal.
__device__ __forceinline__ bool adds_nothing(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }

// atomicAdd by ONE lane of a warp, as written. Left alone, ptxas wraps an atomic whose address is warp-uniform in its own
// aggregation (VOTEU / UPOPC / FLO, then a pair of SHFL that consume the result right behind the atomic) -- for inline PTX as
// well. That duplicates the aggregation done by hand here and, worse, ends the overlap of the atomic's latency with the work
// that follows it. `lane_zero` is (lane & a run-time zero): it makes the address formally lane-dependent, so the atomic stays one
// plain ATOMG.
__device__ __forceinline__ unsigned long long atom_add_u64(unsigned long long *p, unsigned long long v, unsigned lane_zero)
{
	return atomicAdd(p + lane_zero, v);
}

// warp-aggregated append: returns this lane's slot in the output queue (valid when `keep`)
__device__ __forceinline__ unsigned long long queue_append(bool keep, unsigned long long *counter, unsigned zero)
{
	const unsigned m = __ballot_sync(kFullMask, keep);
	if (m == 0u) return 0ull;
	const unsigned lane = threadIdx.x & 31u, leader = (unsigned)__ffs((int)m) - 1u;
	unsigned long long base = 0;
	if (lane == leader) base = atom_add_u64(counter, (unsigned long long)__popc(m), lane & zero);
	base = __shfl_sync(kFullMask, base, (int)leader);
	return base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
}

// Bounce 0 of the batch's S samples. All samples of a tmpLifetime block share the primary hit and the camera ray
// (pathtracer.glsl:113-127, 206-211), so a thread takes one pixel and a GROUP of its samples: FetchInfo, the normal and the
// emission term are evaluated once per group and only the material switch -- the part that consumes the sample's Sobol pair --
// runs per sample, and the queue slots of several samples are requested with one atomic per warp. Path id = s * npix + pixel.
constexpr int kMaxPrimaryChunk = 8;

// One item of bounce 0: pixel (item % npix), samples [(item / npix) * group, ... + group) of the batch. Called by all 32 lanes of a warp
// together (uniform trip counts, warp-collective queue appends). known_tri != -3: the caller has already read the pixel's cached primary
// hit (and, known_mat >= 0, its material id).
__device__ __forceinline__ void shade_primary_item(const ShadeBuffers &B, const PTArgs &A, const CameraArgs &cam, int group, int chunk, unsigned npix,
                                                   unsigned long long items, unsigned long long item, float bx, float by, int dims, V3 sun, bool last,
                                                   unsigned lane, int32_t known_tri, int32_t known_mat)
{
	const bool live = item < items;
	const unsigned pix = live ? (unsigned)(item % npix) : 0u;
	const int s_first = live ? (int)(item / npix) * group : 0;
	V3 dir0 = v3(0, 0, 0), ret0 = v3(0.f, 0.f, 0.f);
	Surface sf;
	sf.illum = 0;
	sf.origin = v3(0, 0, 0);
	bool hit = false;
	float fbx = 0.f, fby = 0.f;
	unsigned bias_bits = 0u;
	if (live) {
		dir0 = camera_dir(cam, A.width, A.height, (int)(pix % (unsigned)A.width), (int)(pix / (unsigned)A.width), bx, by);
		const uchar2 bb = B.bias[pix];
		fbx = (float)bb.x / 255.0f;
		fby = (float)bb.y / 255.0f;
		bias_bits = (unsigned)bb.x | ((unsigned)bb.y << 8);
		const int32_t tri = known_tri != -3 ? known_tri : B.prim_tri[pix];
		hit = tri != -1;
		if (hit) {
			const float2 uv = B.prim_uv[pix];
			fetch_surface(B, tri, uv.x, uv.y, sf, known_mat);
			ret0 = ret0 + v3(1.f, 1.f, 1.f) * sf.emissive; // :139 with color = 1, ret = 0
		} else if (B.conn_rays == nullptr)
			ret0 = ret0 + v3(1.f, 1.f, 1.f) * sun; // :130-135
	}
	// The group's samples in chunks: every sample of a chunk is shaded first and parked (two float4 per sample in local
	// memory), then the warp asks for the chunk's queue slots with ONE atomic and writes the survivors out. One atomic per
	// warp and SAMPLE put a million same-address atomics into every launch -- about what the L2 retires in the kernel's
	// whole run time (the kernel sat at its slot requests: 47 % of the stall samples, profiles/r2l_shade_primary_*).
	for (int k0 = 0; k0 < group; k0 += chunk) {
		float4 c_dir[kMaxPrimaryChunk], c_col[kMaxPrimaryChunk]; // dir + path id; colour + slot offset inside the warp's claim
		unsigned kept = 0u, total = 0u;
		for (int j = 0; j < chunk; ++j) {
			const int sidx = s_first + k0 + j;
			const bool have = live && k0 + j < group && sidx < A.n_samples;
			bool keep = false, conn = false;
			V3 dir = dir0, color = v3(1.f, 1.f, 1.f);
			const unsigned id = (unsigned)sidx * npix + pix;
			if (have) {
				B.ret[id] = make_float4(ret0.x, ret0.y, ret0.z, 0.0f);
				if (!hit) {
					conn = B.conn_rays != nullptr; // the shadow ray starts at the camera
					color = color * sun;
				} else if (!last) {
					const float rx = fract(B.sobol[sidx * dims + 0] + fbx);
					const float ry = fract(B.sobol[sidx * dims + 1] + fby);
					const float xi = A.rr_start >= 0 ? fract(B.sobol[sidx * dims + 2 * A.max_bounce] + fbx) : 0.0f;
					keep = scatter(A, 0, sf, rx, ry, xi, dir, color);
				}
			}
			if (B.conn_rays != nullptr) {
				const unsigned long long cs = queue_append(conn, B.conn_count, A.zero);
				if (conn) {
					B.conn_rays[2 * cs] = make_float4(cam.origin[0], cam.origin[1], cam.origin[2], cam.tmin);
					B.conn_rays[2 * cs + 1] = make_float4(B.sun_dir[0], B.sun_dir[1], B.sun_dir[2], __uint_as_float(id));
					B.conn_color[cs] = make_float4(color.x, color.y, color.z, 0.0f);
				}
			}
			const unsigned m = __ballot_sync(kFullMask, keep);
			if (keep) {
				kept |= 1u << j;
				c_dir[j] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(id));
				c_col[j] = make_float4(color.x, color.y, color.z, __uint_as_float(total + (unsigned)__popc(m & ((1u << lane) - 1u))));
			}
			total += (unsigned)__popc(m);
		}
		if (total != 0u) { // warp-uniform
			unsigned long long base = 0;
			if (lane == 0u) base = atom_add_u64(B.out_count, (unsigned long long)total, lane & A.zero);
			base = __shfl_sync(kFullMask, base, 0);
			for (int j = 0; j < chunk; ++j)
				if ((kept >> j) & 1u) {
					const float4 d4 = c_dir[j], c4 = c_col[j];
					const unsigned long long slot = base + (unsigned long long)__float_as_uint(c4.w);
					B.out_org[slot] = make_float4(sf.origin.x, sf.origin.y, sf.origin.z, cam.tmin);
					B.out_dir[slot] = d4;
					B.out_state[slot] = make_float4(c4.x, c4.y, c4.z, __uint_as_float(bias_bits | ((unsigned)(s_first + k0 + j) << 16)));
				}
		}
	}
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS) k_shade_primary(ShadeBuffers B, PTArgs A, CameraArgs cam, int group, int chunk)
{
	const unsigned npix = (unsigned)A.width * (unsigned)A.height;
	const unsigned n_groups = ((unsigned)A.n_samples + (unsigned)group - 1u) / (unsigned)group;
	const unsigned long long items = (unsigned long long)npix * n_groups;
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
	const unsigned long long rounds = (items + stride - 1u) / stride;
	const unsigned lane = threadIdx.x & 31u;
	float bx, by;
	subpixel_bias(A.subpixel, A.tmp_lifetime, A.first_spp, &bx, &by);
	const int dims = A.dims;
	const V3 sun = v3(A.sun[0], A.sun[1], A.sun[2]);
	const bool last = A.max_bounce == 1; // the loop of Render ends after this segment
	for (unsigned long long r = 0; r < rounds; ++r) {
		const unsigned long long item = r * stride + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
		shade_primary_item(B, A, cam, group, chunk, npix, items, item, bx, by, dims, sun, last, lane, -3, -1);
	}
}

enum { kClassMiss = 0, kClassDiffuse = 1, kClassGlossy = 2, kClassMirror = 3, kClassGlass = 4, kClassOther = 5, kClassNone = 7 };

// One queue entry of bounce b: everything of Render's loop body after the intersection (pathtracer.glsl:130-201) and the hand-over
// to the next queue. Called by all 32 lanes of a warp together (the queue appends are warp-collective); tri_idx == -2 = no entry.
// sure: the caller knows from the entry's class that the path goes on (diffuse, mirror, dielectric, pass-through; not the
// last bounce, no Russian roulette) and its block has reserved the slots of all such entries of the round with ONE atomic: the entry's slot
// is *sure_base + sure_idx (*sure_base is kNoBase until the reservation has come back).
constexpr unsigned kNoBase = 0xffffffffu;
__device__ __forceinline__ void shade_queue_entry(const ShadeBuffers &B, const PTArgs &A, int b, float tmin, bool last, int dims, unsigned q,
                                                  int32_t tri_idx, int32_t mat_idx, bool sure = false, unsigned *sure_base = nullptr, unsigned sure_idx = 0u)
{
	bool keep = false, conn = false;
	V3 origin = v3(0, 0, 0), dir = v3(0, 0, 0), color = v3(0, 0, 0);
	unsigned id = 0, bias_bits = 0;
	if (tri_idx != -2) {
		const float4 r1 = B.in_dir[q];
		const float4 st = B.in_state[q];
		dir = v3(r1.x, r1.y, r1.z);
		id = __float_as_uint(r1.w);
		color = v3(st.x, st.y, st.z);
		bias_bits = __float_as_uint(st.w);
		V3 add;
		if (tri_idx == -1) { // :130-135
			add = color * v3(A.sun[0], A.sun[1], A.sun[2]);
			if (B.conn_rays != nullptr) {
				conn = true; // the shadow ray starts where this segment started: the previous hit point
				color = add;
				const float4 r0 = B.in_org[q];
				origin = v3(r0.x, r0.y, r0.z);
				add = v3(0.f, 0.f, 0.f);
			}
		} else {
			const float2 uv = B.in_uv[q];
			Surface sf;
			fetch_surface(B, tri_idx, uv.x, uv.y, sf, mat_idx);
			origin = sf.origin;
			add = color * sf.emissive; // :139
			if (!last) {
				const unsigned s = bias_bits >> 16; // the path's sample within the batch travels with its bias bytes (= id / npix)
				const float fbx = __ldg(B.unorm8 + (bias_bits & 0xffu)), fby = __ldg(B.unorm8 + ((bias_bits >> 8) & 0xffu)); // byte / 255.0f
				const float rx = fract(B.sobol[s * dims + 2 * b] + fbx);
				const float ry = fract(B.sobol[s * dims + 2 * b + 1] + fby);
				const float xi = A.rr_start >= 0 ? fract(B.sobol[s * dims + 2 * A.max_bounce + b] + fbx) : 0.0f;
				keep = scatter(A, b, sf, rx, ry, xi, dir, color);
			}
		}
		if (!adds_nothing(add)) {
			float4 r4 = B.ret[id];
			r4.x = r4.x + add.x;
			r4.y = r4.y + add.y;
			r4.z = r4.z + add.z;
			B.ret[id] = r4;
		}
	}
	unsigned long long slot = queue_append(keep && !sure, B.out_count, A.zero);
	const unsigned sure_m = __ballot_sync(kFullMask, sure);
	if (sure_m != 0u) { // warp-uniform. One lane waits for the block's reservation (an atomic read: the flag is written with an atomic too)
		const unsigned leader = (unsigned)__ffs((int)sure_m) - 1u;
		unsigned base = 0u;
		if ((threadIdx.x & 31u) == leader)
			while ((base = atomicOr(sure_base, 0u)) == kNoBase) {}
		base = __shfl_sync(kFullMask, base, (int)leader);
		if (sure) slot = base + sure_idx; // queue positions fit 32 bits (alloc_wavefront)
	}
	if (keep) {
		B.out_org[slot] = make_float4(origin.x, origin.y, origin.z, tmin);
		B.out_dir[slot] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(id));
		B.out_state[slot] = make_float4(color.x, color.y, color.z, __uint_as_float(bias_bits));
	}
	if (B.conn_rays != nullptr) {
		const unsigned long long cs = queue_append(conn, B.conn_count, A.zero);
		if (conn) {
			B.conn_rays[2 * cs] = make_float4(origin.x, origin.y, origin.z, tmin);
			B.conn_rays[2 * cs + 1] = make_float4(B.sun_dir[0], B.sun_dir[1], B.sun_dir[2], __uint_as_float(id));
			B.conn_color[cs] = make_float4(color.x, color.y, color.z, 0.0f);
		}
	}
}

// bounce b >= 1 over the current queue
template <int MIN_CTAS, int BLOCK, bool REGROUP = true>
__global__ void __launch_bounds__(BLOCK, MIN_CTAS) k_shade_bounce(ShadeBuffers B, PTArgs A, int b, float tmin)
{
	// queue positions fit 32 bits (a batch holds fewer than 2^31 + a grid's worth of paths, alloc_wavefront)
	const unsigned total = (unsigned)*B.in_count;
	const unsigned stride = gridDim.x * blockDim.x;
	const unsigned rounds = total / stride + (total % stride != 0u ? 1u : 0u);
	const int dims = A.dims;
	const bool last = b == A.max_bounce - 1; // the loop of Render ends after this segment: only the emission / sun terms are left
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(B.segments, (unsigned long long)total);
	// Each block regroups its BLOCK (256) queue entries by shading branch before shading them, so that a warp mostly runs ONE of
	// miss / diffuse / glossy / mirror / glass instead of all of them one after the other (13.4 of 32 lanes active without
	// this). Which lane shades which entry never reaches the image: every path owns its slots and the next queue's order is
	// free. The branch comes from a one-byte-per-triangle table, the hit index of the NEXT round is already in flight while
	// this one is shaded, and ranks come from one match.any per warp: two barriers per round, no shared-memory atomics.
	constexpr unsigned kWarps = BLOCK / 32;
	static_assert(BLOCK == 256 || BLOCK == 128 || BLOCK == 64, "eight, four or two warps");
	__shared__ __align__(16) unsigned s_count[2][8][kWarps]; // [round parity][class][warp]; the other parity is zeroed for the next round
	__shared__ unsigned short s_order[BLOCK];             // regrouped position -> entry of the round
	__shared__ int32_t s_tri[BLOCK];                    // regrouped position -> hit triangle
	__shared__ int32_t s_mat[BLOCK];                    // regrouped position -> its material id
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const unsigned lt_mask = (1u << lane) - 1u;
	unsigned q0 = blockIdx.x * blockDim.x;
	// two rounds of look-ahead: the hit index of round r+2 and the class byte of round r+1 are in flight while round r is shaded
	int32_t tri_cur = -2, tri_n1 = -2;       // -2 = past the end of the queue
	unsigned cls_cur = (unsigned)kClassNone << 24; // class << 24 | material id, as in the table
	if (REGROUP) {
	if (rounds > 0) {
		if (q0 + threadIdx.x < total) tri_cur = B.in_tri[q0 + threadIdx.x];
		if (rounds > 1 && q0 + stride + threadIdx.x < total) tri_n1 = B.in_tri[q0 + stride + threadIdx.x];
		cls_cur = tri_cur == -2 ? (unsigned)kClassNone << 24 : tri_cur == -1 ? (unsigned)kClassMiss << 24 : B.tri_class[tri_cur];
	}
	if (threadIdx.x < 16u * kWarps) (&s_count[0][0][0])[threadIdx.x] = 0u;
	__syncthreads();
	}
	for (unsigned r = 0; r < rounds; ++r, q0 += stride) {
		unsigned q;
		int32_t tri_idx, mat_idx = -1;
		if (REGROUP) {
		const int32_t tri_mine = tri_cur;
		const unsigned cls = cls_cur >> 24;
		const int32_t mat_mine = (cls_cur & 0x00ffffffu) != 0x00ffffffu ? (int32_t)(cls_cur & 0x00ffffffu) : -1;
		{
			const unsigned q2 = q0 + 2u * stride + threadIdx.x;
			const int32_t tri_n2 = (r + 2 < rounds && q2 < total) ? B.in_tri[q2] : -2;
			cls_cur = tri_n1 == -2 ? (unsigned)kClassNone << 24 : tri_n1 == -1 ? (unsigned)kClassMiss << 24 : B.tri_class[tri_n1];
			tri_cur = tri_n1;
			tri_n1 = tri_n2;
		}
		const unsigned par = (unsigned)(r & 1u);
		// rank among the warp's lanes of the same class, and the warp's count per class
		const unsigned same = __match_any_sync(kFullMask, cls);
		const unsigned rank = (unsigned)__popc(same & lt_mask);
		if (rank == 0u) s_count[par][cls][warp] = (unsigned)__popc(same);
		if (threadIdx.x < 8u * kWarps) (&s_count[par ^ 1u][0][0])[threadIdx.x] = 0u; // next round's counters
		__syncthreads();
		{
			// position = (entries of lower classes) + (entries of my class in lower warps) + rank. Lane c < 8 sums class c's eight
			// warp counts (two 16-byte loads), an 8-lane exclusive scan turns the totals into class offsets, and every lane
			// picks its class's number with one shuffle.
			unsigned total_c = 0u, before_c = 0u;
			if (lane < 8u) {
				unsigned v[kWarps < 4 ? 4 : kWarps];
				if (kWarps == 2) {
					const uint2 a = *reinterpret_cast<const uint2 *>(&s_count[par][lane][0]);
					v[0] = a.x; v[1] = a.y;
				} else {
					const uint4 a = *reinterpret_cast<const uint4 *>(&s_count[par][lane][0]);
					v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
				}
				if (kWarps == 8) {
					const uint4 b4 = *reinterpret_cast<const uint4 *>(&s_count[par][lane][kWarps - 4]);
					v[kWarps - 4] = b4.x; v[kWarps - 3] = b4.y; v[kWarps - 2] = b4.z; v[kWarps - 1] = b4.w;
				}
#pragma unroll
				for (unsigned w = 0; w < kWarps; ++w) {
					total_c += v[w];
					if (w < warp) before_c += v[w];
				}
			}
			unsigned incl = total_c;
#pragma unroll
			for (unsigned d = 1; d < 8u; d <<= 1) {
				const unsigned t0 = __shfl_up_sync(kFullMask, incl, d);
				if (lane >= d) incl += t0;
			}
			const unsigned base_c = incl - total_c + before_c;
			const unsigned pos = __shfl_sync(kFullMask, base_c, (int)cls) + rank;
			s_order[pos] = (unsigned short)threadIdx.x;
			s_tri[pos] = tri_mine;
			s_mat[pos] = mat_mine;
		}
		__syncthreads();
		q = q0 + s_order[threadIdx.x];
		tri_idx = s_tri[threadIdx.x];
		mat_idx = s_mat[threadIdx.x];
		} else { // TUNING VARIANT: every lane shades its own entry, no regrouping, no barriers
			q = q0 + threadIdx.x;
			tri_idx = q < total ? B.in_tri[q] : -2;
		}
		shade_queue_entry(B, A, b, tmin, last, dims, q, tri_idx, mat_idx);
	}
}

// The same stage with EPT queue entries per thread and the warps of a block taking 32-entry chunks of the regrouped round
// dynamically. With one entry per thread a block's warps end up with one class each -- a warp of misses is done long before a warp
// of diffuse hits -- and a quarter of the stall samples sat at the round's first barrier (profiles/r2aa). Here a round is 128 * EPT
// entries: per class a contiguous region (class totals by shared-memory atomics, then one reservation per warp-level group), 4 * EPT
// chunks for 4 warps to share, and two barriers per 128 * EPT entries instead of per 128.
template <int EPT, int BLOCK = 128>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK) k_shade_bounce_multi(ShadeBuffers B, PTArgs A, int b, float tmin)
{
	constexpr unsigned kN = (unsigned)BLOCK * EPT;
	const unsigned total = (unsigned)*B.in_count; // queue positions fit 32 bits (alloc_wavefront)
	const unsigned stride = gridDim.x * kN;
	const unsigned rounds = total / stride + (total % stride != 0u ? 1u : 0u);
	const int dims = A.dims;
	const bool last = b == A.max_bounce - 1;
	// Entries of a class that always goes on get their slots in the next queue from ONE reservation per block and round (their number is
	// known once the round's class totals are). One request per 32-entry chunk is 0.85 M atomics on one address in a 27 M-segment launch,
	// about as many as the L2 retires in the kernel's whole run time (the same limit bounce 0 ran into, shade_primary_item).
	const bool block_slots = !last && A.rr_start < 0 && A.early_slots != 0;
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(B.segments, (unsigned long long)total);
	__shared__ unsigned s_total[2][8], s_fill[2][8], s_next[2];
	__shared__ unsigned s_base[2];
	__shared__ unsigned short s_order[2][kN];
	__shared__ int32_t s_tri[2][kN], s_mat[2][kN];
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lt_mask = (1u << lane) - 1u;
	if (threadIdx.x < 16u) (&s_total[0][0])[threadIdx.x] = 0u;
	else if (threadIdx.x < 32u) (&s_fill[0][0])[threadIdx.x - 16u] = 0u;
	else if (threadIdx.x < 34u) s_next[threadIdx.x - 32u] = 0u;
	else if (threadIdx.x < 36u) s_base[threadIdx.x - 34u] = kNoBase;
	__syncthreads();
	unsigned q0 = blockIdx.x * kN;
	for (unsigned r = 0; r < rounds; ++r, q0 += stride) {
		const unsigned buf = r & 1u;
		int32_t tri[EPT];
		unsigned cw[EPT], rank[EPT];
#pragma unroll
		for (int k = 0; k < EPT; ++k) {
			const unsigned q = q0 + (unsigned)k * (unsigned)BLOCK + threadIdx.x;
			tri[k] = q < total ? B.in_tri[q] : -2;
		}
#pragma unroll
		for (int k = 0; k < EPT; ++k)
			cw[k] = tri[k] == -2 ? (unsigned)kClassNone << 24 : tri[k] == -1 ? (unsigned)kClassMiss << 24 : B.tri_class[tri[k]];
#pragma unroll
		for (int k = 0; k < EPT; ++k) {
			const unsigned cls = cw[k] >> 24;
			const unsigned same = __match_any_sync(kFullMask, cls);
			rank[k] = (unsigned)__popc(same & lt_mask);
			if (rank[k] == 0u) atomicAdd(&s_total[buf][cls], (unsigned)__popc(same));
		}
		__syncthreads();
		unsigned reserved = kNoBase; // thread 0: the round's reservation, on its way while the entries are put in order
		{
			const uint4 t0 = *reinterpret_cast<const uint4 *>(&s_total[buf][0]), t1 = *reinterpret_cast<const uint4 *>(&s_total[buf][4]);
			const unsigned tot[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
			if (block_slots && threadIdx.x == 0u) {
				const unsigned n_sure = tot[kClassDiffuse] + tot[kClassMirror] + tot[kClassGlass] + tot[kClassOther];
				reserved = n_sure != 0u ? (unsigned)atom_add_u64(B.out_count, (unsigned long long)n_sure, lane & A.zero) : 0u;
			}
#pragma unroll
			for (int k = 0; k < EPT; ++k) {
				const unsigned cls = cw[k] >> 24;
				unsigned base = 0u;
#pragma unroll
				for (unsigned c = 0; c < 8u; ++c)
					if (c < cls) base += tot[c];
				const unsigned same = __match_any_sync(kFullMask, cls);
				const unsigned leader = (unsigned)__ffs((int)same) - 1u;
				unsigned off = 0u;
				if (lane == leader) off = atomicAdd(&s_fill[buf][cls], (unsigned)__popc(same));
				off = __shfl_sync(kFullMask, off, (int)leader);
				const unsigned pos = base + off + rank[k];
				s_order[buf][pos] = (unsigned short)((unsigned)k * (unsigned)BLOCK + threadIdx.x);
				s_tri[buf][pos] = tri[k];
				s_mat[buf][pos] = (int32_t)cw[k];
			}
			// the other buffer's counters for the next round (everybody has left the previous round: they are all past the barrier above)
			if (threadIdx.x < 8u) { s_total[buf ^ 1u][threadIdx.x] = 0u; s_fill[buf ^ 1u][threadIdx.x] = 0u; }
			if (threadIdx.x == 8u) s_next[buf ^ 1u] = 0u;
			if (threadIdx.x == 9u) s_base[buf ^ 1u] = kNoBase;
		}
		__syncthreads();
		if (block_slots && threadIdx.x == 0u) atomicExch(&s_base[buf], reserved); // whoever needs a slot before this waits for it (shade_queue_entry)
		const unsigned n_valid = q0 >= total ? 0u : (total - q0 < kN ? total - q0 : kN);
		const unsigned n_chunks = (n_valid + 31u) / 32u;
		const unsigned n_miss = s_total[buf][kClassMiss], n_glossy = s_total[buf][kClassGlossy]; // the regions in front of the sure ones
		for (;;) {
			unsigned c = 0u;
			if (lane == 0u) c = atomicAdd(&s_next[buf], 1u);
			c = __shfl_sync(kFullMask, c, 0);
			if (c >= n_chunks) break;
			const unsigned p = c * 32u + lane;
			const unsigned w = (unsigned)s_mat[buf][p]; // class << 24 | material id
			const unsigned cls = w >> 24;
			constexpr unsigned kSureClasses = (1u << kClassDiffuse) | (1u << kClassMirror) | (1u << kClassGlass) | (1u << kClassOther);
			const bool sure = block_slots && ((kSureClasses >> cls) & 1u) != 0u;
			const unsigned mat = w & 0x00ffffffu;
			shade_queue_entry(B, A, b, tmin, last, dims, q0 + s_order[buf][p], s_tri[buf][p], mat != 0x00ffffffu ? (int32_t)mat : -1, sure, &s_base[buf],
			                  p - n_miss - (cls > (unsigned)kClassGlossy ? n_glossy : 0u));
		}
	}
}

// Bounce 0 with the block's items regrouped by the class of the cached primary hit, like k_shade_bounce_multi: 128 threads sort 128 * EPT
// items (pixel, sample group) so that a warp's lanes run the same branch for their 16 samples (21.5 of 32 lanes active unsorted: a
// warp of neighbouring pixels mixes sky, walls and lamps), and take the sorted 32-item chunks dynamically.
template <int EPT>
__global__ void __launch_bounds__(128, 8) k_shade_primary_sorted(ShadeBuffers B, PTArgs A, CameraArgs cam, int group, int chunk)
{
	constexpr unsigned kN = 128u * EPT;
	const unsigned npix = (unsigned)A.width * (unsigned)A.height;
	const unsigned n_groups = ((unsigned)A.n_samples + (unsigned)group - 1u) / (unsigned)group;
	const unsigned long long items = (unsigned long long)npix * n_groups;
	const unsigned long long stride = (unsigned long long)gridDim.x * kN;
	const unsigned long long rounds = (items + stride - 1u) / stride;
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lt_mask = (1u << lane) - 1u;
	float bx, by;
	subpixel_bias(A.subpixel, A.tmp_lifetime, A.first_spp, &bx, &by);
	const int dims = A.dims;
	const V3 sun = v3(A.sun[0], A.sun[1], A.sun[2]);
	const bool last = A.max_bounce == 1;
	__shared__ unsigned s_total[2][8], s_fill[2][8], s_next[2];
	__shared__ unsigned short s_order[2][kN];
	__shared__ int32_t s_tri[2][kN], s_mat[2][kN];
	if (threadIdx.x < 16u) (&s_total[0][0])[threadIdx.x] = 0u;
	else if (threadIdx.x < 32u) (&s_fill[0][0])[threadIdx.x - 16u] = 0u;
	else if (threadIdx.x < 34u) s_next[threadIdx.x - 32u] = 0u;
	__syncthreads();
	unsigned long long i0 = (unsigned long long)blockIdx.x * kN;
	for (unsigned long long r = 0; r < rounds; ++r, i0 += stride) {
		const unsigned buf = (unsigned)(r & 1u);
		int32_t tri[EPT];
		unsigned cw[EPT], rank[EPT];
#pragma unroll
		for (int k = 0; k < EPT; ++k) {
			const unsigned long long item = i0 + (unsigned)k * 128u + threadIdx.x;
			tri[k] = item < items ? B.prim_tri[(unsigned)(item % npix)] : -2;
		}
#pragma unroll
		for (int k = 0; k < EPT; ++k)
			cw[k] = tri[k] == -2 ? (unsigned)kClassNone << 24 : tri[k] == -1 ? (unsigned)kClassMiss << 24 : B.tri_class[tri[k]];
#pragma unroll
		for (int k = 0; k < EPT; ++k) {
			const unsigned cls = cw[k] >> 24;
			const unsigned same = __match_any_sync(kFullMask, cls);
			rank[k] = (unsigned)__popc(same & lt_mask);
			if (rank[k] == 0u) atomicAdd(&s_total[buf][cls], (unsigned)__popc(same));
		}
		__syncthreads();
		{
			const uint4 t0 = *reinterpret_cast<const uint4 *>(&s_total[buf][0]), t1 = *reinterpret_cast<const uint4 *>(&s_total[buf][4]);
			const unsigned tot[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
			for (int k = 0; k < EPT; ++k) {
				const unsigned cls = cw[k] >> 24;
				unsigned base = 0u;
#pragma unroll
				for (unsigned c = 0; c < 8u; ++c)
					if (c < cls) base += tot[c];
				const unsigned same = __match_any_sync(kFullMask, cls);
				const unsigned leader = (unsigned)__ffs((int)same) - 1u;
				unsigned off = 0u;
				if (lane == leader) off = atomicAdd(&s_fill[buf][cls], (unsigned)__popc(same));
				off = __shfl_sync(kFullMask, off, (int)leader);
				const unsigned pos = base + off + rank[k];
				s_order[buf][pos] = (unsigned short)((unsigned)k * 128u + threadIdx.x);
				s_tri[buf][pos] = tri[k];
				s_mat[buf][pos] = (cw[k] & 0x00ffffffu) != 0x00ffffffu ? (int32_t)(cw[k] & 0x00ffffffu) : -1;
			}
			if (threadIdx.x < 8u) { s_total[buf ^ 1u][threadIdx.x] = 0u; s_fill[buf ^ 1u][threadIdx.x] = 0u; }
			if (threadIdx.x == 8u) s_next[buf ^ 1u] = 0u;
		}
		__syncthreads();
		const unsigned n_valid = i0 >= items ? 0u : (items - i0 < (unsigned long long)kN ? (unsigned)(items - i0) : kN);
		const unsigned n_chunks = (n_valid + 31u) / 32u;
		for (;;) {
			unsigned c = 0u;
			if (lane == 0u) c = atomicAdd(&s_next[buf], 1u);
			c = __shfl_sync(kFullMask, c, 0);
			if (c >= n_chunks) break;
			const unsigned p = c * 32u + lane;
			const int32_t t = s_tri[buf][p];
			// entries past the end of the item list (t == -2) sit at the end of the sorted round: an item number >= items switches the lane off
			shade_primary_item(B, A, cam, group, chunk, npix, items, t == -2 ? items : i0 + s_order[buf][p], bx, by, dims, sun, last, lane, t == -2 ? -3 : t, s_mat[buf][p]);
		}
	}
}

// connect: ret += colour * sun for every escaped path whose shadow ray reached the sun (occ == 0)
__global__ void k_connect_apply(const float4 *__restrict__ conn_rays, const uint8_t *__restrict__ occ, const unsigned long long *count,
                                const float4 *__restrict__ conn_color, float4 *__restrict__ ret)
{
	const unsigned long long total = *count;
	for (unsigned long long q = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; q < total; q += (unsigned long long)gridDim.x * blockDim.x) {
		if (occ[q]) continue;
		const unsigned id = __float_as_uint(conn_rays[2 * q + 1].w);
		const float4 c = conn_color[q];
		float4 r = ret[id];
		r.x = r.x + c.x;
		r.y = r.y + c.y;
		r.z = r.z + c.z;
		ret[id] = r;
	}
}

// main() :224-226 for the S samples of the batch, in ascending spp
__global__ void k_accumulate_mean(const float4 *__restrict__ ret, float4 *__restrict__ out, unsigned npix, int first_spp, int n_samples, float clamp)
{
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
		float4 o = out[i];
		for (int s = 0; s < n_samples; ++s) {
			const float4 r = ret[(size_t)s * npix + i];
			const float fs = (float)(first_spp + s), fs1 = (float)(first_spp + s + 1);
			o.x = (o.x * fs + glsl_min(r.x, clamp)) / fs1;
			o.y = (o.y * fs + glsl_min(r.y, clamp)) / fs1;
			o.z = (o.z * fs + glsl_min(r.z, clamp)) / fs1;
			o.w = 1.0f;
		}
		out[i] = o;
	}
}

// sharded mode: clamped radiance added to a sum accumulator in ascending spp; .w counts samples
__global__ void k_accumulate_sum(const float4 *__restrict__ ret, float4 *__restrict__ sum, unsigned npix, int n_samples, float clamp)
{
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
		float4 o = sum[i];
		for (int s = 0; s < n_samples; ++s) {
			const float4 r = ret[(size_t)s * npix + i];
			o.x += glsl_min(r.x, clamp);
			o.y += glsl_min(r.y, clamp);
			o.z += glsl_min(r.z, clamp);
			o.w += 1.0f;
		}
		sum[i] = o;
	}
}

__global__ void k_resolve_sum(const float4 *__restrict__ sum, float4 *__restrict__ out, unsigned npix)
{
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
		const float4 s = sum[i];
		out[i] = s.w > 0.0f ? make_float4(s.x / s.w, s.y / s.w, s.z / s.w, 1.0f) : make_float4(0.f, 0.f, 0.f, 1.0f);
	}
}

// diffuse colour of primaryray.glsl:59-71 from the 100-byte Triangle record
__device__ __forceinline__ V3 diffuse_of(const uchar4 *__restrict__ texels, const int4 *__restrict__ tex_table, const Material &m, const float *t, float u, float v)
{
	if (texels != nullptr && m.dtex != -1) {
		const float w = 1.0f - u - v;
		const float s = t[18] * u + t[20] * v + t[22] * w, tt = t[19] * u + t[21] * v + t[23] * w;
		return sample_texture(texels, tex_table[m.dtex], s, tt);
	}
	return v3(m.dr, m.dg, m.db);
}

// primaryray.glsl main() :46-94 after the intersection
__global__ void k_view(const uint8_t *__restrict__ tris, const Material *__restrict__ mats, const uchar4 *__restrict__ texels,
                       const int4 *__restrict__ tex_table, const int32_t *__restrict__ hit_tri, const float2 *__restrict__ hit_uv, int type,
                       unsigned npix, float4 *__restrict__ out)
{
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
		const int32_t tri = hit_tri[i];
		V3 c = v3(0.f, 0.f, 0.f);
		if (tri != -1) {
			const float *t = (const float *)(tris + (size_t)tri * 100u);
			const Material m = mats[*(const int32_t *)(t + 24)];
			const float2 uv = hit_uv[i];
			if (type == 0) c = diffuse_of(texels, tex_table, m, t, uv.x, uv.y);
			else if (type == 1) c = v3(m.sr, m.sg, m.sb);
			else if (type == 2) c = v3(m.er, m.eg, m.eb);
			else if (type == 4) c = normalize(bary(t + 9, t + 12, t + 15, uv.x, uv.y));
			else if (type == 5) c = bary(t, t + 3, t + 6, uv.x, uv.y);
		}
		out[i] = make_float4(c.x, c.y, c.z, 1.0f);
	}
}

__global__ void k_debug_math(int op, const float *__restrict__ x, const float *__restrict__ y, unsigned long long n, float *__restrict__ o, float *__restrict__ o2)
{
	for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
		if (op == 0) {
			float s, c;
			detmath::sincos(x[i], &s, &c);
			o[i] = s;
			o2[i] = c;
		} else
			o[i] = detmath::pow(x[i], y[i]);
	}
}

} // namespace adypt

using namespace adypt;

// ================================================================================================
struct adypt_tracer {
	adypt_scene *scene = nullptr;
	int device = 0; // copy of scene->device: destruction must not depend on the scene still being alive
	adypt_pt_config cfg{};
	int width = 0, height = 0;
	unsigned npix = 0;
	cudaStream_t stream = nullptr;
	CameraArgs cam{};
	int spp = 0;                 // m_pt_local_spp
	bool prim_valid = false;     // primary cache holds the hit of the current block/stratum
	int prim_block = -1;
	// images
	float4 *d_result = nullptr, *d_sum = nullptr;
	int32_t *d_prim_tri = nullptr;
	float2 *d_prim_uv = nullptr;
	uchar2 *d_bias = nullptr;
	float *d_unorm8 = nullptr;  // 256 floats, see ShadeBuffers::unorm8
	std::vector<uint8_t> h_bias;
	// Sobol
	uint32_t *d_dirs = nullptr;
	float *d_sobol = nullptr;
	// wavefront
	int batch_samples = 0;       // S
	unsigned long long capacity = 0; // paths per batch = S * npix
	uint8_t *d_slab = nullptr;               // one allocation behind all wavefront buffers below
	size_t slab_skew = 0;                    // tuning (ADYPT_WAVEFRONT_SKEW): extra bytes (multiple of 256) between the buffers
	float4 *d_rays[2] = {nullptr, nullptr};  // ray queues: [0, capacity) origins + tmin, [capacity, 2 capacity) directions + path id
	int32_t *d_hit_tri = nullptr;
	float2 *d_hit_uv = nullptr;
	float4 *d_state[2] = {nullptr, nullptr}; // per queue entry: throughput + bias bytes
	float4 *d_ret = nullptr;                 // per path: radiance
	// device counters, sized from maxBounce: [0, max_bounce] queue lengths, [conn_base + b] connect queues, [seg_slot] segments
	// statistic, [work_slot] the work counter of this tracer's traversal launches (all on t->stream, so one is enough; scene.h)
	unsigned long long *d_counts = nullptr;
	int n_slots = 0, conn_base = 0, seg_slot = 0, work_slot = 0;
	int dims_cap = 0;            // Sobol dimensions d_dirs / d_sobol are sized for
	size_t sobol_cap = 0;        // floats in d_sobol
	// connect stage (off by default: the reference's sun test is commented out, pathtracer.glsl:132)
	bool sun_visibility = false;
	float sun_dir[3] = {0.f, 0.f, 0.f};
	int rr_start = -1; // Russian roulette (opt-in extension): first bounce it applies to, -1 = off
	float4 *d_conn_rays = nullptr, *d_conn_color = nullptr;
	uint8_t *d_conn_occ = nullptr;
	unsigned long long conn_capacity = 0;
	uint64_t launches_at_create = 0;
	uint64_t host_segments = 0;  // primary segments (known on the host)
	// measurement hooks (adypt_tracer_set_profiling): CUDA events around every stage launch, and / or the instrumented
	// traversal kernel that counts the nodes and triangles the wavefront's rays touch. Off by default.
	int bounce_ctas = 0;   // tuning (ADYPT_BOUNCE_CTAS): see the switch in run_batch; 0 = default (four entries per thread, dynamic chunks)
	int primary_ctas = 0;  // tuning (ADYPT_PRIMARY_CTAS): bounce-0 kernel: 2 / 3 / 4 = unsorted at that many CTAs per SM, 22 = class-sorted 2 items per thread; 0 = default (class-sorted, 4 per thread)
	unsigned early_slots = 1; // tuning (ADYPT_EARLY_SLOTS)
	int primary_chunk = 0; // tuning (ADYPT_PRIMARY_CHUNK): samples of a group shaded per queue-slot request (1..8); 0 = default
	int primary_group = 0; // tuning (ADYPT_PRIMARY_GROUP): samples of one pixel a thread of the bounce-0 stage shades; 0 = default
	int profiling = 0;
	std::vector<cudaEvent_t> ev_pool;
	size_t ev_used = 0;
	struct Span { int stage; size_t e0, e1; };
	std::vector<Span> spans;
	double stage_ms[ADYPT_STAGE_COUNT] = {0};
	uint64_t stage_launches[ADYPT_STAGE_COUNT] = {0};
	unsigned long long *d_trace_stats = nullptr; // 2 x kStatSlots counters (bounce queues, primary rays), profiling bit 1
};

namespace {

constexpr unsigned long long kDefaultMaxPaths = 48ull << 20;

// ---- stage timing: an event pair around each stage launch while profiling bit 0 is set; resolved on request
int resolve_spans(adypt_tracer *t)
{
	if (t->spans.empty()) return ADYPT_OK;
	ADYPT_CUDA(cudaStreamSynchronize(t->stream));
	for (const adypt_tracer::Span &sp : t->spans) {
		float ms = 0.0f;
		ADYPT_CUDA(cudaEventElapsedTime(&ms, t->ev_pool[sp.e0], t->ev_pool[sp.e1]));
		t->stage_ms[sp.stage] += (double)ms;
		t->stage_launches[sp.stage] += 1;
	}
	t->spans.clear();
	t->ev_used = 0;
	return ADYPT_OK;
}

struct StageTimer { // records on construction and in end(); a no-op unless stage timing is on
	adypt_tracer *t;
	size_t e0 = 0;
	int stage;
	bool on;
	StageTimer(adypt_tracer *tr, int st) : t(tr), stage(st), on((tr->profiling & 1) != 0)
	{
		if (!on) return;
		if (t->ev_used + 2 > 8192) resolve_spans(t); // bounded pool: a long render is folded into the sums as it goes
		while (t->ev_pool.size() < t->ev_used + 2) {
			cudaEvent_t e = nullptr;
			if (cudaEventCreate(&e) != cudaSuccess) { on = false; return; }
			t->ev_pool.push_back(e);
		}
		e0 = t->ev_used;
		t->ev_used += 2;
		cudaEventRecord(t->ev_pool[e0], t->stream);
	}
	void end()
	{
		if (!on) return;
		cudaEventRecord(t->ev_pool[e0 + 1], t->stream);
		t->spans.push_back({stage, e0, e0 + 1});
		on = false;
	}
};

void free_tracer(adypt_tracer *t)
{
	DeviceGuard g(t->device);
	if (t->stream) cudaStreamSynchronize(t->stream);
	for (cudaEvent_t e : t->ev_pool) cudaEventDestroy(e);
	cudaFree(t->d_trace_stats);
	cudaFree(t->d_result); cudaFree(t->d_sum); cudaFree(t->d_prim_tri); cudaFree(t->d_prim_uv); cudaFree(t->d_bias); cudaFree(t->d_unorm8);
	cudaFree(t->d_dirs); cudaFree(t->d_sobol); cudaFree(t->d_slab);
	cudaFree(t->d_counts); cudaFree(t->d_conn_rays); cudaFree(t->d_conn_color); cudaFree(t->d_conn_occ);
	if (t->stream) cudaStreamDestroy(t->stream);
	delete t;
}

int check_config(const adypt_pt_config *c, bool roulette)
{
	if (c->max_bounce < 1) return fail(ADYPT_EINVAL, "maxBounce must be >= 1");
	// the reference's generator carries 10 005 dimensions (Sobol.hpp:9) and uses 2*maxBounce of them (OglPathTracer.cpp:139)
	if (2ll * c->max_bounce > sobol_max_dim()) return fail(ADYPT_ERANGE, "maxBounce needs more Sobol dimensions than the reference's table has (2*maxBounce <= 10005)");
	if (roulette && 3ll * c->max_bounce > sobol_max_dim()) return fail(ADYPT_ERANGE, "Russian roulette needs 3*maxBounce Sobol dimensions (<= 10005)");
	if (c->subpixel < 1 || c->tmp_lifetime < 1) return fail(ADYPT_EINVAL, "subpixel and tmpLifetime must be >= 1");
	return ADYPT_OK;
}

// buffers whose size follows the configuration: queue counters (maxBounce) and the direction numbers (2 or 3 x maxBounce dimensions)
int alloc_config_buffers(adypt_tracer *t)
{
	const int mb = t->cfg.max_bounce;
	const int n_slots = 2 * (mb + 2) + 2;
	if (n_slots != t->n_slots) {
		unsigned long long keep = 0;
		if (t->d_counts) {
			ADYPT_CUDA(cudaStreamSynchronize(t->stream));
			ADYPT_CUDA(cudaMemcpy(&keep, t->d_counts + t->seg_slot, sizeof(keep), cudaMemcpyDeviceToHost));
			cudaFree(t->d_counts);
			t->d_counts = nullptr;
		}
		ADYPT_CUDA(cudaMalloc((void **)&t->d_counts, (size_t)n_slots * sizeof(unsigned long long)));
		ADYPT_CUDA(cudaMemset(t->d_counts, 0, (size_t)n_slots * sizeof(unsigned long long)));
		t->n_slots = n_slots;
		t->conn_base = mb + 2;
		t->seg_slot = n_slots - 2;
		t->work_slot = n_slots - 1;
		ADYPT_CUDA(cudaMemcpy(t->d_counts + t->seg_slot, &keep, sizeof(keep), cudaMemcpyHostToDevice));
	}
	long long dims = 3ll * mb; // room for the opt-in roulette draws, as far as the table goes
	if (dims > sobol_max_dim()) dims = sobol_max_dim();
	if ((int)dims > t->dims_cap) {
		if (t->d_dirs) { ADYPT_CUDA(cudaStreamSynchronize(t->stream)); cudaFree(t->d_dirs); t->d_dirs = nullptr; }
		ADYPT_CUDA(cudaMalloc((void **)&t->d_dirs, (size_t)dims * 32u * 4u));
		ADYPT_CUDA(cudaMemcpy(t->d_dirs, sobol_directions(), (size_t)dims * 32u * 4u, cudaMemcpyHostToDevice));
		t->dims_cap = (int)dims;
	}
	return ADYPT_OK;
}

int grid_for(unsigned long long n, int block, int sm_count)
{
	unsigned long long g = (n + block - 1) / block;
	const unsigned long long cap = (unsigned long long)sm_count * 16ull;
	if (g > cap) g = cap;
	return g < 1 ? 1 : (int)g;
}

int alloc_wavefront(adypt_tracer *t)
{
	int S = t->cfg.tmp_lifetime;
	const unsigned long long by_mem = std::max(1ull, kDefaultMaxPaths / (unsigned long long)t->npix);
	if ((unsigned long long)S > by_mem) S = (int)by_mem;
	if (S > 65535) S = 65535; // a path's sample number shares a word with its bias bytes in the state queue
	const unsigned long long cap = (unsigned long long)S * t->npix;
	if (cap >= (1ull << 32)) return fail(ADYPT_ERANGE, "batch too large for 32-bit path ids");
	const size_t sobol_need = (size_t)t->dims_cap * (size_t)S;
	if (sobol_need > t->sobol_cap) {
		if (t->d_sobol) { ADYPT_CUDA(cudaStreamSynchronize(t->stream)); cudaFree(t->d_sobol); t->d_sobol = nullptr; t->sobol_cap = 0; }
		ADYPT_CUDA(cudaMalloc((void **)&t->d_sobol, sobol_need * 4u));
		t->sobol_cap = sobol_need;
	}
	if (cap == t->capacity && S == t->batch_samples) return ADYPT_OK;
	// One slab for all wavefront buffers, carved at fixed offsets: how the arrays lie relative to each other (DRAM channel /
	// TLB aliasing between the 16 sample planes of `ret` and the queues the bounce-0 kernel writes side by side) then no
	// longer depends on what the process allocated before -- the same kernel measured 0.96 ms in one process and 1.25 ms in
	// another with separately allocated buffers.
	if (t->d_slab) {
		ADYPT_CUDA(cudaStreamSynchronize(t->stream));
		cudaFree(t->d_slab);
	}
	t->d_slab = nullptr;
	t->d_rays[0] = t->d_rays[1] = nullptr; t->d_hit_tri = nullptr; t->d_hit_uv = nullptr; t->d_state[0] = t->d_state[1] = nullptr; t->d_ret = nullptr;
	t->capacity = 0;
	const size_t sizes[7] = {(size_t)cap * 32u, (size_t)cap * 32u, (size_t)cap * 4u, (size_t)cap * 8u, (size_t)cap * 16u, (size_t)cap * 16u, (size_t)cap * 16u};
	size_t offs[7], total = 0;
	const size_t skew = t->slab_skew;
	for (int k = 0; k < 7; ++k) {
		offs[k] = total;
		total += ((sizes[k] + 255u) & ~(size_t)255u) + skew;
	}
	ADYPT_CUDA(cudaMalloc((void **)&t->d_slab, total));
	t->d_rays[0] = (float4 *)(t->d_slab + offs[0]);
	t->d_rays[1] = (float4 *)(t->d_slab + offs[1]);
	t->d_hit_tri = (int32_t *)(t->d_slab + offs[2]);
	t->d_hit_uv = (float2 *)(t->d_slab + offs[3]);
	t->d_state[0] = (float4 *)(t->d_slab + offs[4]);
	t->d_state[1] = (float4 *)(t->d_slab + offs[5]);
	t->d_ret = (float4 *)(t->d_slab + offs[6]);
	t->capacity = cap;
	t->batch_samples = S;
	return ADYPT_OK;
}

// trace the primary rays of the stratum `spp` belongs to and store them in the primary cache
int trace_primary(adypt_tracer *t, float bx, float by)
{
	adypt_scene *s = t->scene;
	StageTimer tg(t, ADYPT_STAGE_GENERATE);
	k_generate<<<grid_for(t->npix, 256, s->sm_count), 256, 0, t->stream>>>(t->cam, t->width, t->height, bx, by, t->d_rays[0]);
	count_launch();
	ADYPT_CUDA(cudaGetLastError());
	tg.end();
	StageTimer tt(t, ADYPT_STAGE_TRACE_PRIMARY);
	ADYPT_TRY(launch_trace(s, t->d_rays[0], t->npix, t->d_prim_tri, nullptr, t->d_prim_uv, nullptr, t->stream, nullptr, t->d_counts + t->work_slot,
	                       (t->profiling & 2) ? t->d_trace_stats + kStatSlots : nullptr));
	tt.end();
	t->host_segments += t->npix;
	return ADYPT_OK;
}

// one batch: samples [first, first+n) of every pixel, all inside one tmpLifetime block
int run_batch(adypt_tracer *t, int first, int n, bool sum_mode)
{
	adypt_scene *s = t->scene;
	const adypt_pt_config &c = t->cfg;
	const int block = first / c.tmp_lifetime;
	PTArgs A;
	A.max_bounce = c.max_bounce; A.subpixel = c.subpixel; A.tmp_lifetime = c.tmp_lifetime;
	A.clamp = c.clamp; A.sun[0] = c.sun[0]; A.sun[1] = c.sun[1]; A.sun[2] = c.sun[2];
	A.width = t->width; A.height = t->height; A.first_spp = first; A.n_samples = n;
	A.rr_start = t->rr_start;
	A.zero = 0u;
	A.early_slots = t->early_slots;
	A.dims = (t->rr_start >= 0 ? 3 : 2) * c.max_bounce;
	// uSpp % uTmpLife == 0 -> trace and store the primary hit; otherwise reuse it (pathtracer.glsl:113-127)
	if (first % c.tmp_lifetime == 0 || !t->prim_valid || t->prim_block != block) {
		float bx, by;
		const int idx = (first / c.tmp_lifetime) % (c.subpixel * c.subpixel);
		const float unit = 1.0f / (float)c.subpixel;
		bx = (float)(idx / c.subpixel) * unit;
		by = (float)(idx % c.subpixel) * unit;
		ADYPT_TRY(trace_primary(t, bx, by));
		t->prim_valid = true;
		t->prim_block = block;
	}
	const int dims = A.dims;
	StageTimer ts(t, ADYPT_STAGE_OTHER);
	k_sobol<<<(dims * n + 127) / 128, 128, 0, t->stream>>>(t->d_dirs, dims, first, n, t->d_sobol);
	count_launch();
	ADYPT_CUDA(cudaMemsetAsync(t->d_counts, 0, (size_t)t->seg_slot * sizeof(unsigned long long), t->stream)); // queue lengths only
	ts.end();
	unsigned long long *trace_stats = (t->profiling & 2) ? t->d_trace_stats : nullptr;
	ShadeBuffers B;
	B.shade = s->d_shade; B.tri_class = s->d_tri_class; B.mats = s->d_mats; B.texels = s->d_texels; B.tex_table = s->d_tex_table; B.bias = t->d_bias; B.unorm8 = t->d_unorm8; B.sobol = t->d_sobol;
	B.prim_tri = t->d_prim_tri; B.prim_uv = t->d_prim_uv;
	B.in_org = B.in_dir = nullptr; B.in_tri = t->d_hit_tri; B.in_uv = t->d_hit_uv; B.in_state = nullptr; B.in_count = nullptr;
	B.out_org = t->d_rays[1]; B.out_dir = t->d_rays[1] + t->capacity; B.out_state = t->d_state[1]; B.out_count = t->d_counts + 1;
	B.ret = t->d_ret; B.segments = t->d_counts + t->seg_slot;
	const unsigned long long total = (unsigned long long)n * t->npix;
	B.conn_rays = nullptr; B.conn_color = nullptr; B.conn_count = nullptr;
	B.sun_dir[0] = t->sun_dir[0]; B.sun_dir[1] = t->sun_dir[1]; B.sun_dir[2] = t->sun_dir[2];
	if (t->sun_visibility) {
		if (t->conn_capacity < total) {
			cudaFree(t->d_conn_rays); cudaFree(t->d_conn_color); cudaFree(t->d_conn_occ);
			t->d_conn_rays = t->d_conn_color = nullptr; t->d_conn_occ = nullptr; t->conn_capacity = 0;
			ADYPT_CUDA(cudaMalloc((void **)&t->d_conn_rays, total * 32u));
			ADYPT_CUDA(cudaMalloc((void **)&t->d_conn_color, total * 16u));
			ADYPT_CUDA(cudaMalloc((void **)&t->d_conn_occ, total));
			t->conn_capacity = total;
		}
		B.conn_rays = t->d_conn_rays;
		B.conn_color = t->d_conn_color;
		B.conn_count = t->d_counts + t->conn_base;
	}
	// connect stage of bounce b: any-hit over the shadow rays queued by the shade kernel, then add the sun term
	auto connect = [&](int b) -> int {
		if (!t->sun_visibility) return ADYPT_OK;
		StageTimer tc(t, ADYPT_STAGE_CONNECT);
		ADYPT_TRY(launch_trace(s, t->d_conn_rays, total, nullptr, nullptr, nullptr, t->d_conn_occ, t->stream, t->d_counts + t->conn_base + b, t->d_counts + t->work_slot));
		k_connect_apply<<<grid_for(total, 256, s->sm_count), 256, 0, t->stream>>>(t->d_conn_rays, t->d_conn_occ, t->d_counts + t->conn_base + b, t->d_conn_color, t->d_ret);
		count_launch();
		ADYPT_CUDA(cudaGetLastError());
		tc.end();
		return ADYPT_OK;
	};
	{
		StageTimer tp(t, ADYPT_STAGE_SHADE_PRIMARY);
		const int group = t->primary_group > 0 ? t->primary_group : 16; // samples of a pixel per thread (1 .. 16 measured: profiles/r2d_primary_sweep.log)
		const unsigned long long items = (unsigned long long)t->npix * (unsigned long long)((n + group - 1) / group);
		const int g = grid_for(items, 256, s->sm_count);
		const int chunk = t->primary_chunk > 0 && t->primary_chunk <= kMaxPrimaryChunk ? t->primary_chunk : 4; // samples per queue-slot request
		switch (t->primary_ctas) { // tuning, same results: profiles/r2u_primary_chunk_sweep.log, profiles/r2ah_primary_sorted.log
		case 2: k_shade_primary<2><<<g, 256, 0, t->stream>>>(B, A, t->cam, group, chunk); break; // unsorted, 120 registers
		case 3: k_shade_primary<3><<<g, 256, 0, t->stream>>>(B, A, t->cam, group, chunk); break; // unsorted, 80 registers
		case 4: k_shade_primary<4><<<g, 256, 0, t->stream>>>(B, A, t->cam, group, chunk); break; // unsorted, 64 registers
		case 22: k_shade_primary_sorted<2><<<grid_for(items, 256, s->sm_count), 128, 0, t->stream>>>(B, A, t->cam, group, chunk); break;
		default: k_shade_primary_sorted<4><<<grid_for(items, 512, s->sm_count), 128, 0, t->stream>>>(B, A, t->cam, group, chunk); break; // items regrouped by class, 4 per thread
		}
		count_launch();
		ADYPT_CUDA(cudaGetLastError());
		tp.end();
	}
	ADYPT_TRY(connect(0));
	int cur = 1;
	for (int b = 1; b < c.max_bounce; ++b) {
		// extend: queue length is read on the device
		StageTimer te(t, ADYPT_STAGE_TRACE_BOUNCE);
		ADYPT_TRY(launch_trace(s, t->d_rays[cur], total, t->d_hit_tri, nullptr, t->d_hit_uv, nullptr, t->stream, t->d_counts + b, t->d_counts + t->work_slot, trace_stats,
		                       t->d_rays[cur] + t->capacity));
		te.end();
		B.in_org = t->d_rays[cur];
		B.in_dir = t->d_rays[cur] + t->capacity;
		B.in_state = t->d_state[cur];
		B.in_count = t->d_counts + b;
		B.out_org = t->d_rays[cur ^ 1];
		B.out_dir = t->d_rays[cur ^ 1] + t->capacity;
		B.out_state = t->d_state[cur ^ 1];
		B.out_count = t->d_counts + b + 1;
		if (t->sun_visibility) B.conn_count = t->d_counts + t->conn_base + b;
		StageTimer tb(t, ADYPT_STAGE_SHADE_BOUNCE);
		// Default: 128-thread blocks, eight per SM, FOUR queue entries per thread with the block's warps taking the regrouped round's
		// 32-entry chunks dynamically (k_shade_bounce_multi): 10.2 ms per 64 spp of C3. One entry per thread (k_shade_bounce): 12.1 ms with
		// 128-thread blocks, 14.2 with 256, 14.5 with 64, 14.7 without regrouping (profiles/r2g, r2k, r2t, r2ad logs). The alternatives stay
		// selectable for A/B runs (ADYPT_BOUNCE_CTAS).
		switch (t->bounce_ctas) {
		case 1: k_shade_bounce<8, 128, false><<<2 * grid_for(total, 256, s->sm_count), 128, 0, t->stream>>>(B, A, b, c.ray_tmin); break; // no regrouping
		case 4: k_shade_bounce<4, 256><<<grid_for(total, 256, s->sm_count), 256, 0, t->stream>>>(B, A, b, c.ray_tmin); break;
		case 8: k_shade_bounce<8, 128><<<2 * grid_for(total, 256, s->sm_count), 128, 0, t->stream>>>(B, A, b, c.ray_tmin); break;
		case 16: k_shade_bounce<16, 64><<<4 * grid_for(total, 256, s->sm_count), 64, 0, t->stream>>>(B, A, b, c.ray_tmin); break;
		case 22: k_shade_bounce_multi<2><<<grid_for(total, 256, s->sm_count), 128, 0, t->stream>>>(B, A, b, c.ray_tmin); break;
		case 26: k_shade_bounce_multi<6><<<grid_for(total, 768, s->sm_count), 128, 0, t->stream>>>(B, A, b, c.ray_tmin); break;
		case 32: k_shade_bounce_multi<2, 256><<<grid_for(total, 512, s->sm_count), 256, 0, t->stream>>>(B, A, b, c.ray_tmin); break;
		case 34: k_shade_bounce_multi<4, 256><<<grid_for(total, 1024, s->sm_count), 256, 0, t->stream>>>(B, A, b, c.ray_tmin); break;
		case 28: k_shade_bounce_multi<8><<<grid_for(total, 1024, s->sm_count), 128, 0, t->stream>>>(B, A, b, c.ray_tmin); break;
		default: k_shade_bounce_multi<4><<<grid_for(total, 512, s->sm_count), 128, 0, t->stream>>>(B, A, b, c.ray_tmin); break;
		}
		count_launch();
		ADYPT_CUDA(cudaGetLastError());
		tb.end();
		ADYPT_TRY(connect(b));
		cur ^= 1;
	}
	const int g = grid_for(t->npix, 256, s->sm_count);
	StageTimer ta(t, ADYPT_STAGE_ACCUMULATE);
	if (sum_mode) k_accumulate_sum<<<g, 256, 0, t->stream>>>(t->d_ret, t->d_sum, t->npix, n, c.clamp);
	else k_accumulate_mean<<<g, 256, 0, t->stream>>>(t->d_ret, t->d_result, t->npix, first, n, c.clamp);
	count_launch();
	ADYPT_CUDA(cudaGetLastError());
	ta.end();
	return ADYPT_OK;
}

// split [first, first+n) into batches that do not cross tmpLifetime blocks nor exceed S samples
int run_range(adypt_tracer *t, int first, int n, bool sum_mode)
{
	ADYPT_TRY(alloc_wavefront(t));
	const int L = t->cfg.tmp_lifetime;
	int s = first;
	const int end = first + n;
	while (s < end) {
		const int block_end = (s / L + 1) * L;
		const int m = std::min(std::min(end, block_end) - s, t->batch_samples);
		ADYPT_TRY(run_batch(t, s, m, sum_mode));
		s += m;
	}
	return ADYPT_OK;
}

} // namespace

extern "C" {

int adypt_tracer_create(adypt_scene *scene, const adypt_pt_config *config, int32_t width, int32_t height, uint64_t bias_seed,
                        adypt_tracer **out)
{
	return guarded([&]() -> int {
	if (!scene || !config || !out) return fail(ADYPT_EINVAL, "scene/config/out is NULL");
	*out = nullptr;
	if (width <= 0 || height <= 0 || (uint64_t)width * (uint64_t)height >= (1ull << 31)) return fail(ADYPT_EINVAL, "bad image size");
	if (!scene->d_tris || !scene->d_mats || !scene->d_shade) return fail(ADYPT_EINVAL, "scene has no triangles/materials: traversal-only scenes cannot shade");
	if (scene->bad_matid_tri >= 0)
		return fail(ADYPT_EINVAL, "triangle " + std::to_string(scene->bad_matid_tri) + " has a material id outside [0, n_materials) (an OBJ face without a usemtl, or an unknown material): the scene can be traversed but not shaded");
	ADYPT_TRY(check_config(config, false));
	DeviceGuard g(scene->device);
	adypt_tracer *t = new adypt_tracer;
	t->scene = scene;
	t->device = scene->device;
	t->cfg = *config;
	t->width = width;
	t->height = height;
	t->npix = (unsigned)width * (unsigned)height;
	cudaError_t e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
	const size_t np = t->npix;
	if (e == cudaSuccess) e = cudaMalloc((void **)&t->d_result, np * 16u);
	if (e == cudaSuccess) e = cudaMalloc((void **)&t->d_sum, np * 16u);
	if (e == cudaSuccess) e = cudaMalloc((void **)&t->d_prim_tri, np * 4u);
	if (e == cudaSuccess) e = cudaMalloc((void **)&t->d_prim_uv, np * 8u);
	if (e == cudaSuccess) e = cudaMalloc((void **)&t->d_bias, np * 2u);
	if (e == cudaSuccess) e = cudaMalloc((void **)&t->d_unorm8, 256u * 4u);
	if (e == cudaSuccess) {
		k_unorm8_table<<<1, 256>>>(t->d_unorm8);
		count_launch();
		e = cudaDeviceSynchronize();
	}
	if (e == cudaSuccess) e = cudaMemset(t->d_result, 0, np * 16u);
	if (e == cudaSuccess) e = cudaMemset(t->d_sum, 0, np * 16u);
	if (e == cudaSuccess) {
		try {
			t->h_bias.resize(np * 2u);
		} catch (const std::bad_alloc &) {
			free_tracer(t);
			return fail(ADYPT_ENOMEM, "tracer allocation: out of host memory");
		}
		fill_bias(bias_seed, np * 2u, t->h_bias.data());
		e = cudaMemcpy(t->d_bias, t->h_bias.data(), np * 2u, cudaMemcpyHostToDevice);
	}
	if (e != cudaSuccess) {
		free_tracer(t);
		return fail(e == cudaErrorMemoryAllocation ? ADYPT_ENOMEM : ADYPT_ECUDA, std::string("tracer allocation: ") + cudaGetErrorString(e));
	}
	{
		const int rc = alloc_config_buffers(t);
		if (rc != ADYPT_OK) {
			free_tracer(t);
			return rc;
		}
	}
	t->cam.tmin = config->ray_tmin;
	const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
	memcpy(t->cam.inv_proj, ident, 64);
	memcpy(t->cam.inv_view, ident, 64);
	if (const char *e = getenv("ADYPT_PRIMARY_GROUP")) t->primary_group = atoi(e);
	if (const char *e = getenv("ADYPT_PRIMARY_CTAS")) t->primary_ctas = atoi(e);
	if (const char *e = getenv("ADYPT_PRIMARY_CHUNK")) t->primary_chunk = atoi(e);
	if (const char *e = getenv("ADYPT_EARLY_SLOTS")) t->early_slots = atoi(e) != 0 ? 1u : 0u;
	if (const char *e = getenv("ADYPT_BOUNCE_CTAS")) t->bounce_ctas = atoi(e);
	if (const char *e = getenv("ADYPT_WAVEFRONT_SKEW")) t->slab_skew = ((size_t)atol(e) + 255u) & ~(size_t)255u;
	t->launches_at_create = g_launches.load();
	*out = t;
	return ADYPT_OK;
	});
}

int adypt_tracer_destroy(adypt_tracer *t)
{
	return guarded([&]() -> int {
	if (t) free_tracer(t);
	return ADYPT_OK;
	});
}

int adypt_tracer_set_config(adypt_tracer *t, const adypt_pt_config *config)
{
	return guarded([&]() -> int {
	if (!t || !config) return fail(ADYPT_EINVAL, "tracer/config is NULL");
	ADYPT_TRY(check_config(config, t->rr_start >= 0));
	DeviceGuard g(t->device);
	t->cfg = *config;
	t->cam.tmin = config->ray_tmin; // update_config_args, OglPathTracer.cpp:216
	t->prim_valid = false;
	return alloc_config_buffers(t);
	});
}

int adypt_tracer_set_russian_roulette(adypt_tracer *t, int32_t start_bounce)
{
	return guarded([&]() -> int {
	if (!t) return fail(ADYPT_EINVAL, "tracer is NULL");
	if (start_bounce >= 0) ADYPT_TRY(check_config(&t->cfg, true));
	t->rr_start = start_bounce < 0 ? -1 : start_bounce;
	return ADYPT_OK;
	});
}

int adypt_tracer_set_sun_visibility(adypt_tracer *t, int32_t enabled, const float direction[3])
{
	return guarded([&]() -> int {
	if (!t || (enabled && !direction)) return fail(ADYPT_EINVAL, "NULL argument");
	t->sun_visibility = enabled != 0;
	if (enabled) {
		// normalize(vec3(...)) as the shader would (un-fused fp32, IEEE sqrt / divide)
		const float x = direction[0], y = direction[1], z = direction[2];
		const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
		t->sun_dir[0] = x * inv;
		t->sun_dir[1] = y * inv;
		t->sun_dir[2] = z * inv;
	}
	return ADYPT_OK;
	});
}

int adypt_tracer_set_bias(adypt_tracer *t, const uint8_t *rg8)
{
	return guarded([&]() -> int {
	if (!t || !rg8) return fail(ADYPT_EINVAL, "tracer/rg8 is NULL");
	DeviceGuard g(t->scene->device);
	memcpy(t->h_bias.data(), rg8, t->h_bias.size());
	ADYPT_CUDA(cudaStreamSynchronize(t->stream));
	ADYPT_CUDA(cudaMemcpy(t->d_bias, rg8, t->h_bias.size(), cudaMemcpyHostToDevice));
	return ADYPT_OK;
	});
}

int adypt_tracer_get_bias(adypt_tracer *t, uint8_t *rg8)
{
	return guarded([&]() -> int {
	if (!t || !rg8) return fail(ADYPT_EINVAL, "tracer/rg8 is NULL");
	memcpy(rg8, t->h_bias.data(), t->h_bias.size());
	return ADYPT_OK;
	});
}

int adypt_tracer_set_camera(adypt_tracer *t, const float projection[16], const float view[16], const float position[3])
{
	return guarded([&]() -> int {
	if (!t || !projection || !view || !position) return fail(ADYPT_EINVAL, "NULL argument");
	t->cam.origin[0] = position[0]; t->cam.origin[1] = position[1]; t->cam.origin[2] = position[2];
	mat4_inverse(projection, t->cam.inv_proj); // OglPathTracer.cpp:30-31
	mat4_inverse(view, t->cam.inv_view);
	t->prim_valid = false;
	return ADYPT_OK;
	});
}

int adypt_tracer_primary(adypt_tracer *t, int32_t viewer_type)
{
	return guarded([&]() -> int {
	if (!t) return fail(ADYPT_EINVAL, "tracer is NULL");
	DeviceGuard g(t->scene->device);
	ADYPT_TRY(alloc_wavefront(t));
	t->spp = 0; // OglPathTracer.cpp:55
	t->cam.tmin = t->cfg.ray_tmin;
	adypt_scene *s = t->scene;
	k_generate<<<grid_for(t->npix, 256, s->sm_count), 256, 0, t->stream>>>(t->cam, t->width, t->height, 0.0f, 0.0f, t->d_rays[0]);
	count_launch();
	ADYPT_TRY(launch_trace(s, t->d_rays[0], t->npix, t->d_hit_tri, nullptr, t->d_hit_uv, nullptr, t->stream, nullptr, t->d_counts + t->work_slot));
	k_view<<<grid_for(t->npix, 256, s->sm_count), 256, 0, t->stream>>>(s->d_tris, s->d_mats, s->d_texels, s->d_tex_table, t->d_hit_tri, t->d_hit_uv, viewer_type, t->npix, t->d_result);
	count_launch();
	ADYPT_CUDA(cudaGetLastError());
	t->host_segments += t->npix;
	t->prim_valid = false;
	return ADYPT_OK;
	});
}

int adypt_tracer_sample(adypt_tracer *t, int32_t n_spp)
{
	return guarded([&]() -> int {
	if (!t) return fail(ADYPT_EINVAL, "tracer is NULL");
	if (n_spp <= 0) return ADYPT_OK;
	DeviceGuard g(t->scene->device);
	if (t->spp == 0) { // OglPathTracer.cpp:39-46
		t->cam.tmin = t->cfg.ray_tmin;
		ADYPT_CUDA(cudaMemsetAsync(t->d_result, 0, (size_t)t->npix * 16u, t->stream));
		t->prim_valid = false;
	}
	ADYPT_TRY(run_range(t, t->spp, n_spp, false));
	t->spp += n_spp;
	return ADYPT_OK;
	});
}

int adypt_tracer_accumulate(adypt_tracer *t, int32_t first_spp, int32_t n_spp)
{
	return guarded([&]() -> int {
	if (!t) return fail(ADYPT_EINVAL, "tracer is NULL");
	if (n_spp <= 0) return ADYPT_OK;
	if (first_spp < 0 || first_spp % t->cfg.tmp_lifetime != 0) return fail(ADYPT_EINVAL, "first_spp must be a non-negative multiple of tmpLifetime");
	DeviceGuard g(t->scene->device);
	t->cam.tmin = t->cfg.ray_tmin;
	t->prim_valid = false;
	return run_range(t, first_spp, n_spp, true);
	});
}

int adypt_tracer_sum_buffer(adypt_tracer *t, float **device_ptr, uint64_t *n_floats)
{
	return guarded([&]() -> int {
	if (!t || !device_ptr) return fail(ADYPT_EINVAL, "NULL argument");
	*device_ptr = (float *)t->d_sum;
	if (n_floats) *n_floats = (uint64_t)t->npix * 4u;
	return ADYPT_OK;
	});
}

int adypt_tracer_result_buffer(adypt_tracer *t, float **device_ptr, uint64_t *n_floats)
{
	return guarded([&]() -> int {
	if (!t || !device_ptr) return fail(ADYPT_EINVAL, "NULL argument");
	*device_ptr = (float *)t->d_result;
	if (n_floats) *n_floats = (uint64_t)t->npix * 4u;
	return ADYPT_OK;
	});
}

int adypt_tracer_clear_sum(adypt_tracer *t)
{
	return guarded([&]() -> int {
	if (!t) return fail(ADYPT_EINVAL, "tracer is NULL");
	DeviceGuard g(t->scene->device);
	ADYPT_CUDA(cudaMemsetAsync(t->d_sum, 0, (size_t)t->npix * 16u, t->stream));
	return ADYPT_OK;
	});
}

int adypt_tracer_resolve_sum(adypt_tracer *t)
{
	return guarded([&]() -> int {
	if (!t) return fail(ADYPT_EINVAL, "tracer is NULL");
	DeviceGuard g(t->scene->device);
	k_resolve_sum<<<grid_for(t->npix, 256, t->scene->sm_count), 256, 0, t->stream>>>(t->d_sum, t->d_result, t->npix);
	count_launch();
	ADYPT_CUDA(cudaGetLastError());
	return ADYPT_OK;
	});
}

int adypt_tracer_spp(adypt_tracer *t, int32_t *spp)
{
	return guarded([&]() -> int {
	if (!t || !spp) return fail(ADYPT_EINVAL, "NULL argument");
	*spp = t->spp;
	return ADYPT_OK;
	});
}

int adypt_tracer_stream(adypt_tracer *t, void **stream)
{
	return guarded([&]() -> int {
	if (!t || !stream) return fail(ADYPT_EINVAL, "NULL argument");
	*stream = (void *)t->stream;
	return ADYPT_OK;
	});
}

int adypt_tracer_sync(adypt_tracer *t)
{
	return guarded([&]() -> int {
	if (!t) return fail(ADYPT_EINVAL, "tracer is NULL");
	DeviceGuard g(t->scene->device);
	ADYPT_CUDA(cudaStreamSynchronize(t->stream));
	return ADYPT_OK;
	});
}

int adypt_tracer_read(adypt_tracer *t, float *out, int32_t channels)
{
	return guarded([&]() -> int {
	if (!t || !out) return fail(ADYPT_EINVAL, "NULL argument");
	if (channels != 3 && channels != 4) return fail(ADYPT_EINVAL, "channels must be 3 or 4");
	DeviceGuard g(t->scene->device);
	if (channels == 4) {
		ADYPT_CUDA(cudaMemcpyAsync(out, t->d_result, (size_t)t->npix * 16u, cudaMemcpyDeviceToHost, t->stream));
		ADYPT_CUDA(cudaStreamSynchronize(t->stream));
		return ADYPT_OK;
	}
	std::vector<float> tmp((size_t)t->npix * 4u);
	ADYPT_CUDA(cudaMemcpyAsync(tmp.data(), t->d_result, (size_t)t->npix * 16u, cudaMemcpyDeviceToHost, t->stream));
	ADYPT_CUDA(cudaStreamSynchronize(t->stream));
	for (size_t i = 0; i < t->npix; ++i) {
		out[3 * i] = tmp[4 * i];
		out[3 * i + 1] = tmp[4 * i + 1];
		out[3 * i + 2] = tmp[4 * i + 2];
	}
	return ADYPT_OK;
	});
}

int adypt_tracer_save_exr(adypt_tracer *t, const char *filename, int32_t save_as_fp16)
{
	return guarded([&]() -> int {
	if (!t || !filename) return fail(ADYPT_EINVAL, "NULL argument");
	std::vector<float> rgb((size_t)t->npix * 3u);
	ADYPT_TRY(adypt_tracer_read(t, rgb.data(), 3));
	const int rc = adypt_write_exr(filename, rgb.data(), t->width, t->height, save_as_fp16);
	if (rc != ADYPT_OK) return fail(rc, std::string("cannot write ") + filename);
	return ADYPT_OK;
	});
}

int adypt_tracer_primary_rays(adypt_tracer *t, float *rays, int memspace)
{
	return guarded([&]() -> int {
	if (!t || !rays) return fail(ADYPT_EINVAL, "NULL argument");
	DeviceGuard g(t->scene->device);
	t->cam.tmin = t->cfg.ray_tmin;
	float4 *dst = (float4 *)rays;
	if (memspace == ADYPT_MEM_HOST) {
		ADYPT_TRY(alloc_wavefront(t));
		dst = t->d_rays[0];
	} else if (memspace != ADYPT_MEM_DEVICE)
		return fail(ADYPT_EINVAL, "bad memspace");
	k_generate<<<grid_for(t->npix, 256, t->scene->sm_count), 256, 0, t->stream>>>(t->cam, t->width, t->height, 0.0f, 0.0f, dst);
	count_launch();
	ADYPT_CUDA(cudaGetLastError());
	if (memspace == ADYPT_MEM_HOST) ADYPT_CUDA(cudaMemcpyAsync(rays, dst, (size_t)t->npix * 32u, cudaMemcpyDeviceToHost, t->stream));
	ADYPT_CUDA(cudaStreamSynchronize(t->stream));
	return ADYPT_OK;
	});
}

int adypt_debug_math(int32_t device, int32_t op, const float *x, const float *y, uint64_t n, float *out, float *out2)
{
	return guarded([&]() -> int {
	if (!x || !out || (op == 0 && !out2) || (op == 1 && !y) || (op != 0 && op != 1)) return fail(ADYPT_EINVAL, "bad argument");
	if (n == 0) return ADYPT_OK;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(ADYPT_ENODEV, "no such CUDA device (no CPU fallback)");
	DeviceGuard g(device);
	float *d = nullptr;
	ADYPT_CUDA(cudaMalloc((void **)&d, (size_t)n * 16u));
	float *dx = d, *dy = d + n, *d0 = d + 2 * n, *d1 = d + 3 * n;
	cudaError_t e = cudaMemcpy(dx, x, (size_t)n * 4u, cudaMemcpyHostToDevice);
	if (e == cudaSuccess && y) e = cudaMemcpy(dy, y, (size_t)n * 4u, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) {
		k_debug_math<<<256, 256>>>(op, dx, dy, n, d0, d1);
		count_launch();
		e = cudaDeviceSynchronize();
	}
	if (e == cudaSuccess) e = cudaMemcpy(out, d0, (size_t)n * 4u, cudaMemcpyDeviceToHost);
	if (e == cudaSuccess && op == 0) e = cudaMemcpy(out2, d1, (size_t)n * 4u, cudaMemcpyDeviceToHost);
	cudaFree(d);
	if (e != cudaSuccess) return fail(ADYPT_ECUDA, std::string("adypt_debug_math: ") + cudaGetErrorString(e));
	return ADYPT_OK;
	});
}

int adypt_tracer_set_profiling(adypt_tracer *t, int32_t flags)
{
	return guarded([&]() -> int {
	if (!t || flags < 0 || flags > 3) return fail(ADYPT_EINVAL, "bad argument");
	DeviceGuard g(t->device);
	ADYPT_TRY(resolve_spans(t));
	if ((flags & 2) && !t->d_trace_stats) {
		ADYPT_CUDA(cudaMalloc((void **)&t->d_trace_stats, 2 * kStatSlots * sizeof(unsigned long long)));
		ADYPT_CUDA(cudaMemset(t->d_trace_stats, 0, 2 * kStatSlots * sizeof(unsigned long long)));
	}
	t->profiling = flags;
	return ADYPT_OK;
	});
}

int adypt_tracer_get_profile(adypt_tracer *t, adypt_tracer_profile *out, int32_t reset)
{
	return guarded([&]() -> int {
	if (!t || !out) return fail(ADYPT_EINVAL, "NULL argument");
	DeviceGuard g(t->device);
	ADYPT_TRY(resolve_spans(t));
	memset(out, 0, sizeof(*out));
	for (int i = 0; i < ADYPT_STAGE_COUNT; ++i) {
		out->stage_ms[i] = t->stage_ms[i];
		out->stage_launches[i] = t->stage_launches[i];
	}
	if (t->d_trace_stats) {
		unsigned long long h[2 * kStatSlots]; // [0, kStatSlots) bounce queues, [kStatSlots, 2 kStatSlots) primary rays
		ADYPT_CUDA(cudaStreamSynchronize(t->stream));
		ADYPT_CUDA(cudaMemcpy(h, t->d_trace_stats, sizeof(h), cudaMemcpyDeviceToHost));
		out->trace_nodes = h[0]; out->trace_tris = h[1]; out->trace_hits = h[2]; out->trace_max_depth = h[3] > h[kStatSlots + 3] ? h[3] : h[kStatSlots + 3]; out->trace_rays = h[4];
		out->primary_nodes = h[kStatSlots]; out->primary_tris = h[kStatSlots + 1]; out->primary_hits = h[kStatSlots + 2]; out->primary_rays = h[kStatSlots + 4];
		if (reset) ADYPT_CUDA(cudaMemset(t->d_trace_stats, 0, sizeof(h)));
	}
	if (reset)
		for (int i = 0; i < ADYPT_STAGE_COUNT; ++i) { t->stage_ms[i] = 0.0; t->stage_launches[i] = 0; }
	return ADYPT_OK;
	});
}

int adypt_tracer_stats(adypt_tracer *t, uint64_t *segments, uint64_t *launches)
{
	return guarded([&]() -> int {
	if (!t) return fail(ADYPT_EINVAL, "tracer is NULL");
	DeviceGuard g(t->scene->device);
	unsigned long long dev = 0;
	ADYPT_CUDA(cudaStreamSynchronize(t->stream));
	ADYPT_CUDA(cudaMemcpy(&dev, t->d_counts + t->seg_slot, sizeof(dev), cudaMemcpyDeviceToHost));
	if (segments) *segments = t->host_segments + dev;
	if (launches) *launches = g_launches.load() - t->launches_at_create;
	return ADYPT_OK;
	});
}

} // extern "C"
