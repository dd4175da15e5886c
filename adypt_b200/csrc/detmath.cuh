// Deterministic transcendental functions for the shading stage: sin/cos and pow evaluated with a fixed
// sequence of IEEE-754 double operations (explicit fma where fused), so the CUDA kernels and the CPU oracle
// -- which carries its own copy of the same recipe (oracle/detmath.h) -- produce BIT-IDENTICAL fp32 results.
// GLSL leaves sin/cos/pow precision to the driver (the reference's NVIDIA path uses MUFU approximations);
// these are within 1 ulp of the correctly rounded fp32 value (tests/test_detmath.py), i.e. at least as accurate
// as anything the reference could have run on, and they turn image parity from "RMSE below a bound" into
// "every pixel bit-exact".
//   sincos : Cody-Waite reduction by pi/2 (two-term), fdlibm kernel polynomials, quadrant fix-up
//   pow    : exp2(y * log2(x)); log2 by atanh series on m in (sqrt(1/2), sqrt(2)], exp2 by degree-13 Taylor
// B200 keeps full-rate FP64 (unlike sm_103), and shading evaluates these once or twice per path segment.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define ADYPT_HD __host__ __device__ __forceinline__
#else
#define ADYPT_HD inline
#endif

namespace adypt {
namespace detmath {

ADYPT_HD double u64_as_double(uint64_t u)
{
#if defined(__CUDA_ARCH__)
	return __longlong_as_double((long long)u);
#else
	double d;
	memcpy(&d, &u, 8);
	return d;
#endif
}
ADYPT_HD uint64_t double_as_u64(double d)
{
#if defined(__CUDA_ARCH__)
	return (uint64_t)__double_as_longlong(d);
#else
	uint64_t u;
	memcpy(&u, &d, 8);
	return u;
#endif
}

// Coefficient tables. On the device they live in __constant__ memory: ptxas then fetches two doubles per LDCU.128 into
// uniform registers, where the same numbers written as literals cost two UMOV each (26 of the 60 instructions of the
// sin/cos polynomials in round 1's SASS). Same values, same operations, same results.
#define ADYPT_DM_SINCOS_TABLE                                                                                          \
	{                                                                                                                  \
		0.63661977236758134308, 1.57079632679489655800e+00, 6.12323399573676603587e-17, /* 2/pi, pi/2 hi, pi/2 lo */    \
		1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,                          \
		-1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01, /* sin kernel */          \
		-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,                         \
		2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02, /* cos kernel */           \
		0.0                                                                                                            \
	}
#if defined(__CUDACC__)
static __constant__ __align__(16) double kSinCosDev[16] = ADYPT_DM_SINCOS_TABLE;
#endif
static const double kSinCosHost[16] = ADYPT_DM_SINCOS_TABLE;

// sin and cos of an fp32 angle (any finite x with |x| < 2^20; the tracer only passes [0, 2*pi])
ADYPT_HD void sincos(float xf, float *s_out, float *c_out)
{
#if defined(__CUDA_ARCH__)
	const double *K = kSinCosDev;
#else
	const double *K = kSinCosHost;
#endif
	const double x = (double)xf;
	const double k = floor(x * K[0] + 0.5); // nearest multiple of pi/2
	double r = fma(-k, K[1], x);
	r = fma(-k, K[2], r);
	const double z = r * r;
	double ps = K[3];
	ps = fma(ps, z, K[4]);
	ps = fma(ps, z, K[5]);
	ps = fma(ps, z, K[6]);
	ps = fma(ps, z, K[7]);
	ps = fma(ps, z, K[8]);
	const double sn = fma(r * z, ps, r);
	double pc = K[9];
	pc = fma(pc, z, K[10]);
	pc = fma(pc, z, K[11]);
	pc = fma(pc, z, K[12]);
	pc = fma(pc, z, K[13]);
	pc = fma(pc, z, K[14]);
	const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
	const int q = (int)((long long)k & 3);
	const double s = (q == 0) ? sn : (q == 1) ? cs : (q == 2) ? -sn : -cs;
	const double c = (q == 0) ? cs : (q == 1) ? -sn : (q == 2) ? -cs : sn;
	*s_out = (float)s;
	*c_out = (float)c;
}

#define ADYPT_DM_LOG_TABLE                                                                                                       \
	{                                                                                                                            \
		1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0, /* atanh series */ \
		1.44269504088896338700, 1.41421356237309514547                                               /* 1/ln 2, sqrt 2 */ \
	}
#define ADYPT_DM_EXP_TABLE                                                                                                       \
	{                                                                                                                            \
		0.69314718055994528623, 1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,        \
		1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0                 /* ln 2, Taylor 1/13! .. 1/3! */ \
	}
#if defined(__CUDACC__)
static __constant__ __align__(16) double kLogDev[12] = ADYPT_DM_LOG_TABLE;
static __constant__ __align__(16) double kExpDev[12] = ADYPT_DM_EXP_TABLE;
#endif
static const double kLogHost[12] = ADYPT_DM_LOG_TABLE;
static const double kExpHost[12] = ADYPT_DM_EXP_TABLE;

// log2 of a positive, finite, normal double
ADYPT_HD double log2_pos(double x)
{
#if defined(__CUDA_ARCH__)
	const double *K = kLogDev;
#else
	const double *K = kLogHost;
#endif
	const uint64_t bits = double_as_u64(x);
	int e = (int)(bits >> 52) - 1023;
	double m = u64_as_double((bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull); // [1, 2)
	if (m > K[11]) {
		m *= 0.5;
		e += 1;
	}
	const double f = m - 1.0;
	const double s = f / (2.0 + f);
	const double z = s * s;
	double p = K[0];
	p = fma(p, z, K[1]);
	p = fma(p, z, K[2]);
	p = fma(p, z, K[3]);
	p = fma(p, z, K[4]);
	p = fma(p, z, K[5]);
	p = fma(p, z, K[6]);
	p = fma(p, z, K[7]);
	p = fma(p, z, K[8]);
	p = fma(p, z, K[9]);
	p = fma(p, z, 1.0);
	const double ln_m = 2.0 * s * p;
	return fma(ln_m, K[10], (double)e);
}

ADYPT_HD double exp2_any(double t)
{
#if defined(__CUDA_ARCH__)
	const double *K = kExpDev;
#else
	const double *K = kExpHost;
#endif
	if (!(t < 1100.0)) return t != t ? t : u64_as_double(0x7ff0000000000000ull);
	if (t < -1100.0) return 0.0;
	const double n = floor(t + 0.5);
	const double u = (t - n) * K[0];
	double p = K[1];
	p = fma(p, u, K[2]);
	p = fma(p, u, K[3]);
	p = fma(p, u, K[4]);
	p = fma(p, u, K[5]);
	p = fma(p, u, K[6]);
	p = fma(p, u, K[7]);
	p = fma(p, u, K[8]);
	p = fma(p, u, K[9]);
	p = fma(p, u, K[10]);
	p = fma(p, u, K[11]);
	p = fma(p, u, 0.5);
	p = fma(p, u, 1.0);
	p = fma(p, u, 1.0);
	const long long ni = (long long)n;
	if (ni > 1023) return u64_as_double(0x7ff0000000000000ull);
	if (ni < -1022) return 0.0; // far below the smallest fp32 denormal
	return p * u64_as_double((uint64_t)(ni + 1023) << 52);
}

// pow(x, y) for the shader's uses (x in [0, 1], y > 0); defined everywhere so both sides agree:
// NaN operands or x < 0 -> NaN; x == 0 -> 0 (y > 0), 1 (y == 0), inf (y < 0); x == inf -> inf / 1 / 0
ADYPT_HD float pow(float x, float y)
{
	if (x != x || y != y || x < 0.0f) return (float)u64_as_double(0x7ff8000000000000ull);
	if (x == 0.0f) return y > 0.0f ? 0.0f : (y == 0.0f ? 1.0f : (float)u64_as_double(0x7ff0000000000000ull));
	if (x > 3.40282346638528859812e+38f) return y > 0.0f ? x : (y == 0.0f ? 1.0f : 0.0f);
	if (y == 1.0f) return x; // what the general path returns too (|error| < 2^-45 relative, x is a float): skips ~150 FP64 operations
	return (float)exp2_any((double)y * log2_pos((double)x));
}

} // namespace detmath
} // namespace adypt
