// Host-side arithmetic of the tracer: the matrix algebra the reference gets from glm, the Sobol
// generator and the bias-image fill. Plain IEEE fp32, no contraction (x86-64 SSE2, -ffp-contract=off),
// evaluated in glm's operation order so the results are bit-identical to the reference host code.
#include "hostmath.h"
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>
#include "../../include/adypt_b200.h"
#include "guard.h"

namespace adypt {

namespace {
struct V4 {
	float v[4];
	float operator[](int i) const { return v[i]; }
};
inline V4 mul(V4 a, V4 b) { return V4{{a[0] * b[0], a[1] * b[1], a[2] * b[2], a[3] * b[3]}}; }
inline V4 sub(V4 a, V4 b) { return V4{{a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]}}; }
inline V4 add(V4 a, V4 b) { return V4{{a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]}}; }
inline V4 scale(V4 a, float s) { return V4{{a[0] * s, a[1] * s, a[2] * s, a[3] * s}}; }
} // namespace

// glm's compute_inverse<4,4>: 18 2x2 sub-determinants, four cofactor columns, determinant from the
// first row, then one reciprocal (func_matrix.inl:294-351). m(c,r) = in[4*c + r].
void mat4_inverse(const float in[16], float out[16])
{
	auto m = [&](int c, int r) { return in[4 * c + r]; };
	const float c00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3), c02 = m(1, 2) * m(3, 3) - m(3, 2) * m(1, 3), c03 = m(1, 2) * m(2, 3) - m(2, 2) * m(1, 3);
	const float c04 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3), c06 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3), c07 = m(1, 1) * m(2, 3) - m(2, 1) * m(1, 3);
	const float c08 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2), c10 = m(1, 1) * m(3, 2) - m(3, 1) * m(1, 2), c11 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2);
	const float c12 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3), c14 = m(1, 0) * m(3, 3) - m(3, 0) * m(1, 3), c15 = m(1, 0) * m(2, 3) - m(2, 0) * m(1, 3);
	const float c16 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2), c18 = m(1, 0) * m(3, 2) - m(3, 0) * m(1, 2), c19 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
	const float c20 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1), c22 = m(1, 0) * m(3, 1) - m(3, 0) * m(1, 1), c23 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);
	const V4 f0{{c00, c00, c02, c03}}, f1{{c04, c04, c06, c07}}, f2{{c08, c08, c10, c11}};
	const V4 f3{{c12, c12, c14, c15}}, f4{{c16, c16, c18, c19}}, f5{{c20, c20, c22, c23}};
	const V4 v0{{m(1, 0), m(0, 0), m(0, 0), m(0, 0)}}, v1{{m(1, 1), m(0, 1), m(0, 1), m(0, 1)}};
	const V4 v2{{m(1, 2), m(0, 2), m(0, 2), m(0, 2)}}, v3{{m(1, 3), m(0, 3), m(0, 3), m(0, 3)}};
	const V4 i0 = add(sub(mul(v1, f0), mul(v2, f1)), mul(v3, f2));
	const V4 i1 = add(sub(mul(v0, f0), mul(v2, f3)), mul(v3, f4));
	const V4 i2 = add(sub(mul(v0, f1), mul(v1, f3)), mul(v3, f5));
	const V4 i3 = add(sub(mul(v0, f2), mul(v1, f4)), mul(v2, f5));
	const V4 sa{{+1.f, -1.f, +1.f, -1.f}}, sb{{-1.f, +1.f, -1.f, +1.f}};
	const V4 col[4] = {mul(i0, sa), mul(i1, sb), mul(i2, sa), mul(i3, sb)};
	const float d0 = m(0, 0) * col[0][0], d1 = m(0, 1) * col[1][0], d2 = m(0, 2) * col[2][0], d3 = m(0, 3) * col[3][0];
	const float det = (d0 + d1) + (d2 + d3);
	const float inv_det = 1.0f / det;
	for (int c = 0; c < 4; ++c) {
		const V4 r = scale(col[c], inv_det);
		memcpy(out + 4 * c, r.v, 16);
	}
}

// glm::rotate(m, angle, axis) (ext/matrix_transform.inl:18-46)
static void rotate(float m[16], float angle, const float axis_in[3])
{
	const float c = cosf(angle), s = sinf(angle);
	const float il = 1.0f / sqrtf(axis_in[0] * axis_in[0] + axis_in[1] * axis_in[1] + axis_in[2] * axis_in[2]);
	const float a[3] = {axis_in[0] * il, axis_in[1] * il, axis_in[2] * il};
	const float t[3] = {(1.0f - c) * a[0], (1.0f - c) * a[1], (1.0f - c) * a[2]};
	const float R[3][3] = {
		{c + t[0] * a[0], t[0] * a[1] + s * a[2], t[0] * a[2] - s * a[1]},
		{t[1] * a[0] - s * a[2], c + t[1] * a[1], t[1] * a[2] + s * a[0]},
		{t[2] * a[0] + s * a[1], t[2] * a[1] - s * a[0], c + t[2] * a[2]},
	};
	float res[16];
	for (int col = 0; col < 3; ++col)
		for (int r = 0; r < 4; ++r) res[4 * col + r] = m[r] * R[col][0] + m[4 + r] * R[col][1] + m[8 + r] * R[col][2];
	memcpy(res + 12, m + 12, 16);
	memcpy(m, res, 64);
}

void camera_matrices(float fov_deg, float yaw_deg, float pitch_deg, int width, int height, float proj[16], float view[16])
{
	const float deg = 0.01745329251994329576923690768489f; // glm::radians
	const float aspect = width / (float)height;            // Camera.hpp:30
	float v[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
	const float xaxis[3] = {1.f, 0.f, 0.f}, yaxis[3] = {0.f, 1.f, 0.f};
	rotate(v, -pitch_deg * deg, xaxis); // Camera.cpp:15
	rotate(v, -yaw_deg * deg, yaxis);   // Camera.cpp:16
	memcpy(view, v, 64);
	// glm::tweakedInfinitePerspective(fovy, aspect, 0.01f) with ep = epsilon<float>() (matrix_clip_space.inl:512-533)
	const float fovy = fov_deg * deg, z_near = 0.01f, ep = 1.1920928955078125e-07f;
	const float range = tanf(fovy / 2.0f) * z_near;
	const float left = -range * aspect, right = range * aspect, bottom = -range, top = range;
	memset(proj, 0, 64);
	proj[0] = (2.0f * z_near) / (right - left);
	proj[5] = (2.0f * z_near) / (top - bottom);
	proj[10] = ep - 1.0f;
	proj[11] = -1.0f;
	proj[14] = (ep - 2.0f) * z_near;
}

// ------------------------------------------------------------------------------------------------
namespace {
// Joe-Kuo parameters, five packed words per dimension (layout: tools/derive_sobol_params.py)
const uint32_t kJoeKuoPacked[] = {
#include "sobol_joe_kuo.inc"
};
constexpr int kTableDims = (int)(sizeof(kJoeKuoPacked) / (5 * sizeof(uint32_t)));
// The reference declares kMatrices[10005][32] and m_x[10005] (Sobol.inl:4, Sobol.hpp:9) but lists 10 000 rows; C++ zero-fills
// the other five, so dimensions 10 001 .. 10 005 generate the constant 0. Same here.
constexpr int kMaxDim = 10005;
static_assert(kTableDims <= kMaxDim, "parameter table larger than the reference's");
std::vector<uint32_t> g_dirs;
std::once_flag g_dirs_once;

inline uint32_t packed_bits(const uint32_t *w, int pos, int n) // n <= 17 bits starting at bit pos of the 160-bit record
{
	if (n == 0) return 0u;
	const uint64_t lo = w[pos >> 5], hi = (pos >> 5) + 1 < 5 ? w[(pos >> 5) + 1] : 0u;
	return (uint32_t)(((lo | (hi << 32)) >> (pos & 31)) & ((1ull << n) - 1ull));
}

// Bratley-Fox recurrence: m_k = XOR_i 2^i a_i m_{k-i}  ^  2^s m_{k-s} ^ m_{k-s};  v_k = m_k << (32 - k)
void init_dirs()
{
	g_dirs.assign((size_t)kMaxDim * 32u, 0u);
	for (int j = 0; j < kTableDims; ++j) {
		const uint32_t *w = kJoeKuoPacked + 5 * j;
		const int s = (int)packed_bits(w, 0, 5);
		const uint32_t a = packed_bits(w, 5, 16);
		uint32_t m[32];
		int pos = 21;
		for (int k = 0; k < 32; ++k) {
			if (s == 0) m[k] = 1u; // dimension 1: van der Corput
			else if (k < s) {
				m[k] = (packed_bits(w, pos, k) << 1) | 1u; // m_{k+1} is odd and below 2^(k+1)
				pos += k;
			} else {
				uint32_t x = m[k - s] ^ (m[k - s] << s);
				for (int i = 1; i < s; ++i)
					if ((a >> (s - 1 - i)) & 1u) x ^= m[k - i] << i;
				m[k] = x;
			}
			g_dirs[(size_t)j * 32u + (size_t)k] = m[k] << (31 - k);
		}
	}
}
} // namespace

int sobol_max_dim() { return kMaxDim; }

const uint32_t *sobol_directions()
{
	std::call_once(g_dirs_once, init_dirs);
	return g_dirs.data();
}

void sobol_vector(uint32_t dim, uint32_t index, float *out)
{
	const uint32_t *v = sobol_directions();
	// Gray-code generator state after index+1 calls of Next(): XOR of the columns selected by gray(index+1)
	const uint32_t n = index + 1u, gray = n ^ (n >> 1);
	for (uint32_t j = 0; j < dim; ++j) {
		uint32_t x = 0;
		for (uint32_t g = gray, k = 0; g; g >>= 1, ++k)
			if (g & 1u) x ^= v[j * 32 + k];
		out[j] = (float)(x / 4294967296.0); // Sobol.cpp:20
	}
}

void fill_bias(uint64_t seed, uint64_t n_bytes, uint8_t *out)
{
	// splitmix64 stream, 8 bytes per step
	uint64_t s = seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
	for (uint64_t i = 0; i < n_bytes; i += 8) {
		s += 0x9E3779B97F4A7C15ull;
		uint64_t z = s;
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		z ^= z >> 31;
		for (uint64_t b = 0; b < 8 && i + b < n_bytes; ++b) out[i + b] = (uint8_t)(z >> (8 * b));
	}
}

} // namespace adypt

extern "C" int adypt_camera_matrices(float fov_deg, float yaw_deg, float pitch_deg, int32_t width, int32_t height,
                                     float projection[16], float view[16])
{
	return adypt::guarded([&]() -> int {
	if (!projection || !view || width <= 0 || height <= 0) return ADYPT_EINVAL;
	adypt::camera_matrices(fov_deg, yaw_deg, pitch_deg, width, height, projection, view);
	return ADYPT_OK;
	});
}

extern "C" int adypt_sobol_vector(uint32_t dim, uint32_t index, float *out)
{
	return adypt::guarded([&]() -> int {
	if (!out && dim) return adypt::fail(ADYPT_EINVAL, "out is NULL");
	if (dim > (uint32_t)adypt::sobol_max_dim()) return adypt::fail(ADYPT_ERANGE, "the reference's Sobol table has 10005 dimensions");
	adypt::sobol_vector(dim, index, out);
	return ADYPT_OK;
	});
}
