// TEST INFRASTRUCTURE ONLY (included by oracle.cpp).
// The oracle's copy of the deterministic sin/cos/pow recipe the shading stage is specified with:
// a fixed sequence of IEEE-754 double operations with explicit fma, so CPU and GPU agree bit for bit. The
// product carries its own copy (adypt_b200/csrc/detmath.cuh); tests/test_detmath.py pins this one against
// numpy's correctly rounded functions (<= 1 ulp) and tests/test_gpu_tracer.py checks the two copies agree on
// the GPU.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#define ORACLE_INL inline

namespace detmath {

ORACLE_INL double u64_as_double(uint64_t u)
{
	double d;
	memcpy(&d, &u, 8);
	return d;
}
ORACLE_INL uint64_t double_as_u64(double d)
{
	uint64_t u;
	memcpy(&u, &d, 8);
	return u;
}

// sin and cos of an fp32 angle (any finite x with |x| < 2^20; the tracer only passes [0, 2*pi])
ORACLE_INL void sincos(float xf, float *s_out, float *c_out)
{
	const double x = (double)xf;
	const double k = floor(x * 0.63661977236758134308 + 0.5); // nearest multiple of pi/2
	double r = fma(-k, 1.57079632679489655800e+00, x);
	r = fma(-k, 6.12323399573676603587e-17, r);
	const double z = r * r;
	double ps = 1.58969099521155010221e-10;
	ps = fma(ps, z, -2.50507602534068634195e-08);
	ps = fma(ps, z, 2.75573137070700676789e-06);
	ps = fma(ps, z, -1.98412698298579493134e-04);
	ps = fma(ps, z, 8.33333333332248946124e-03);
	ps = fma(ps, z, -1.66666666666666324348e-01);
	const double sn = fma(r * z, ps, r);
	double pc = -1.13596475577881948265e-11;
	pc = fma(pc, z, 2.08757232129817482790e-09);
	pc = fma(pc, z, -2.75573143513906633035e-07);
	pc = fma(pc, z, 2.48015872894767294178e-05);
	pc = fma(pc, z, -1.38888888888741095749e-03);
	pc = fma(pc, z, 4.16666666666666019037e-02);
	const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
	const int q = (int)((long long)k & 3);
	const double s = (q == 0) ? sn : (q == 1) ? cs : (q == 2) ? -sn : -cs;
	const double c = (q == 0) ? cs : (q == 1) ? -sn : (q == 2) ? -cs : sn;
	*s_out = (float)s;
	*c_out = (float)c;
}

// log2 of a positive, finite, normal double
ORACLE_INL double log2_pos(double x)
{
	const uint64_t bits = double_as_u64(x);
	int e = (int)(bits >> 52) - 1023;
	double m = u64_as_double((bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull); // [1, 2)
	if (m > 1.41421356237309514547) {
		m *= 0.5;
		e += 1;
	}
	const double f = m - 1.0;
	const double s = f / (2.0 + f);
	const double z = s * s;
	double p = 1.0 / 21.0;
	p = fma(p, z, 1.0 / 19.0);
	p = fma(p, z, 1.0 / 17.0);
	p = fma(p, z, 1.0 / 15.0);
	p = fma(p, z, 1.0 / 13.0);
	p = fma(p, z, 1.0 / 11.0);
	p = fma(p, z, 1.0 / 9.0);
	p = fma(p, z, 1.0 / 7.0);
	p = fma(p, z, 1.0 / 5.0);
	p = fma(p, z, 1.0 / 3.0);
	p = fma(p, z, 1.0);
	const double ln_m = 2.0 * s * p;
	return fma(ln_m, 1.44269504088896338700, (double)e);
}

ORACLE_INL double exp2_any(double t)
{
	if (!(t < 1100.0)) return t != t ? t : u64_as_double(0x7ff0000000000000ull);
	if (t < -1100.0) return 0.0;
	const double n = floor(t + 0.5);
	const double u = (t - n) * 0.69314718055994528623;
	double p = 1.0 / 6227020800.0;
	p = fma(p, u, 1.0 / 479001600.0);
	p = fma(p, u, 1.0 / 39916800.0);
	p = fma(p, u, 1.0 / 3628800.0);
	p = fma(p, u, 1.0 / 362880.0);
	p = fma(p, u, 1.0 / 40320.0);
	p = fma(p, u, 1.0 / 5040.0);
	p = fma(p, u, 1.0 / 720.0);
	p = fma(p, u, 1.0 / 120.0);
	p = fma(p, u, 1.0 / 24.0);
	p = fma(p, u, 1.0 / 6.0);
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
ORACLE_INL float pow(float x, float y)
{
	if (x != x || y != y || x < 0.0f) return (float)u64_as_double(0x7ff8000000000000ull);
	if (x == 0.0f) return y > 0.0f ? 0.0f : (y == 0.0f ? 1.0f : (float)u64_as_double(0x7ff0000000000000ull));
	if (x > 3.40282346638528859812e+38f) return y > 0.0f ? x : (y == 0.0f ? 1.0f : 0.0f);
	if (y == 1.0f) return x; // what the general path returns too (|error| < 2^-45 relative, x is a float): skips ~150 FP64 operations
	return (float)exp2_any((double)y * log2_pos((double)x));
}

} // namespace detmath
