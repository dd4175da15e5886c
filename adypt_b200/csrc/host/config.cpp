// The .config instance file: the stand-in for InstanceConfig::{LoadFromFile, GetJson, SaveToFile, SetDefault}
// (src/InstanceConfig.cpp:10-215). Same schema, same acceptance rules and the same text on output, so a
// .config written by either program loads in the other:
//   * every key is mandatory and type-checked the way rapidjson 1.1 does it: "Uint" = a non-negative integer
//     literal that fits 32 bits, "Float" = a literal that rapidjson stores as a DOUBLE (has '.', an exponent, or
//     overflows 64 bits) -- so  "fov": 45  is rejected and must be  45.0  (InstanceConfig.cpp:17-18,
//     dep/rapidjson/document.h:972-977);
//   * output is rapidjson's PrettyWriter layout: 4-space indent, one array element per line, doubles printed
//     with rapidjson's own digit generation (Grisu2, not always the shortest: 0.3f comes out as
//     0.30000001192092898) and fixed/exponent switch-over (dtoa.h Prettify) -- the text is byte-identical.
// A small recursive-descent JSON reader replaces rapidjson here; numbers follow its normal-precision path
// (64-bit digit accumulation, then one multiply/divide by a power of ten).
#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include "../common.h"

namespace {

struct JValue {
	enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
	bool b = false;
	bool is_double = false, is_uint = false; // rapidjson's kDoubleFlag / kUintFlag
	double d = 0.0;
	uint32_t u = 0;
	std::string s;
	std::vector<JValue> arr;
	std::vector<std::pair<std::string, JValue>> obj;
	const JValue *find(const char *key) const
	{
		for (const auto &kv : obj)
			if (kv.first == key) return &kv.second;
		return nullptr;
	}
};

struct Parser {
	const char *p, *end;
	bool ok = true;
	void ws()
	{
		while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
	}
	bool lit(const char *w)
	{
		const size_t n = strlen(w);
		if ((size_t)(end - p) >= n && memcmp(p, w, n) == 0) { p += n; return true; }
		return false;
	}
	static double pow10(int n)
	{
		double r = 1.0;
		static const double e[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
		while (n > 22) { r *= 1e22; n -= 22; }
		return r * e[n];
	}
	static double fast_path(double sig, int exp10)
	{
		if (exp10 < -308) return 0.0;
		return exp10 >= 0 ? sig * pow10(exp10) : sig / pow10(-exp10);
	}
	void number(JValue &v)
	{
		v.kind = JValue::Number;
		const bool minus = p < end && *p == '-';
		if (minus) ++p;
		if (p >= end || *p < '0' || *p > '9') { ok = false; return; }
		uint64_t i = 0;
		double d = 0.0;
		bool use_double = false;
		int digits = 0, exp_adj = 0;
		if (*p == '0') ++p;
		else
			while (p < end && *p >= '0' && *p <= '9') {
				if (!use_double) {
					if (i > 1844674407370955161ull || (i == 1844674407370955161ull && *p > '5')) { use_double = true; d = (double)i; }
				}
				if (use_double) {
					if (d >= 1.7976931348623157e307) { ok = false; return; } // rapidjson: kParseErrorNumberTooBig (reader.h:1232-1236)
					d = d * 10.0 + (*p - '0');
				} else
					i = i * 10 + (uint64_t)(*p - '0');
				++digits;
				++p;
			}
		if (p < end && *p == '.') {
			++p;
			if (p >= end || *p < '0' || *p > '9') { ok = false; return; }
			if (!use_double) {
				while (p < end && *p >= '0' && *p <= '9') {
					if (i > 0x1FFFFFFFFFFFFFull) break; // 2^53 - 1: further digits cannot change a double
					i = i * 10 + (uint64_t)(*p - '0');
					--exp_adj;
					if (i != 0) ++digits;
					++p;
				}
				d = (double)i;
				use_double = true;
			}
			while (p < end && *p >= '0' && *p <= '9') {
				if (digits < 17) {
					d = d * 10.0 + (*p - '0');
					--exp_adj;
					if (d > 0.0) ++digits;
				}
				++p;
			}
		}
		int exp = 0;
		if (p < end && (*p == 'e' || *p == 'E')) {
			if (!use_double) { d = (double)i; use_double = true; }
			++p;
			bool eneg = false;
			if (p < end && (*p == '+' || *p == '-')) eneg = *p++ == '-';
			if (p >= end || *p < '0' || *p > '9') { ok = false; return; }
			exp = *p++ - '0';
			if (eneg) {
				while (p < end && *p >= '0' && *p <= '9') {
					exp = exp * 10 + (*p++ - '0');
					if (exp >= 214748364)
						while (p < end && *p >= '0' && *p <= '9') ++p;
				}
			} else {
				const int max_exp = 308 - exp_adj; // reader.h:1313-1319: a larger exponent is kParseErrorNumberTooBig
				while (p < end && *p >= '0' && *p <= '9') {
					exp = exp * 10 + (*p++ - '0');
					if (exp > max_exp) { ok = false; return; }
				}
			}
			if (eneg) exp = -exp;
		}
		if (use_double) {
			const int e10 = exp + exp_adj;
			double r = e10 < -308 ? fast_path(fast_path(d, -308), e10 + 308) : fast_path(d, e10);
			v.d = minus ? -r : r;
			v.is_double = true;
		} else {
			v.is_double = false;
			v.d = minus ? -(double)i : (double)i;
			v.is_uint = !minus && i <= 0xFFFFFFFFull;
			v.u = (uint32_t)i;
		}
	}
	void string(std::string &out)
	{
		++p; // opening quote
		while (p < end && *p != '"') {
			if (*p == '\\') {
				++p;
				if (p >= end) { ok = false; return; }
				switch (*p) {
				case 'n': out += '\n'; break;
				case 't': out += '\t'; break;
				case 'r': out += '\r'; break;
				case 'b': out += '\b'; break;
				case 'f': out += '\f'; break;
				case 'u': {
					if (end - p < 5) { ok = false; return; }
					unsigned cp = 0;
					for (int k = 1; k <= 4; ++k) {
						const char c = p[k];
						cp = cp * 16 + (unsigned)(c >= '0' && c <= '9' ? c - '0' : (c | 0x20) - 'a' + 10);
					}
					p += 4;
					if (cp < 0x80) out += (char)cp;
					else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
					else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
					break;
				}
				default: out += *p; break; // \" \\ \/
				}
				++p;
			} else
				out += *p++;
		}
		if (p >= end) { ok = false; return; }
		++p;
	}
	void value(JValue &v, int depth = 0)
	{
		ws();
		if (p >= end || depth > 64) { ok = false; return; }
		if (*p == '{') {
			v.kind = JValue::Object;
			++p;
			ws();
			if (p < end && *p == '}') { ++p; return; }
			while (ok) {
				ws();
				if (p >= end || *p != '"') { ok = false; return; }
				std::string key;
				string(key);
				ws();
				if (!ok || p >= end || *p != ':') { ok = false; return; }
				++p;
				v.obj.emplace_back(key, JValue());
				value(v.obj.back().second, depth + 1);
				ws();
				if (p < end && *p == ',') { ++p; continue; }
				if (p < end && *p == '}') { ++p; return; }
				ok = false;
			}
		} else if (*p == '[') {
			v.kind = JValue::Array;
			++p;
			ws();
			if (p < end && *p == ']') { ++p; return; }
			while (ok) {
				v.arr.emplace_back();
				value(v.arr.back(), depth + 1);
				ws();
				if (p < end && *p == ',') { ++p; continue; }
				if (p < end && *p == ']') { ++p; return; }
				ok = false;
			}
		} else if (*p == '"') {
			v.kind = JValue::String;
			string(v.s);
		} else if (lit("true")) { v.kind = JValue::Bool; v.b = true; }
		else if (lit("false")) { v.kind = JValue::Bool; v.b = false; }
		else if (lit("null")) v.kind = JValue::Null;
		else number(v);
	}
};

// rapidjson's IsFloat(): stored as a double and within float range
bool is_float(const JValue *v) { return v && v->kind == JValue::Number && v->is_double && v->d >= -3.4028234e38 && v->d <= 3.4028234e38; }
bool is_uint(const JValue *v) { return v && v->kind == JValue::Number && !v->is_double && v->is_uint; }

std::string undefined(const char *type, const char *key, const char *where)
{
	return std::string("[PARSER]ERR: undefined ") + type + " \"" + key + "\" in " + where;
}

// ---- double -> decimal digits the way rapidjson's Writer does it (internal/dtoa.h: Grisu2, then Prettify), so that a
// saved .config has the reference's text byte for byte. Grisu2 (Loitsch, PLDI 2010) works on 64-bit "do-it-yourself"
// floats: the value and its rounding boundaries are scaled by a cached power of ten c_k ~ 10^k so that the product's
// exponent lands in a fixed window, digits are peeled off the upper boundary until they identify the interval, and a
// final "weeding" step moves the last digit towards the value. It is correct but not always shortest, which is why
// std::to_chars cannot stand in for it.

struct DiyFp {
	uint64_t f = 0;
	int e = 0;
	DiyFp() {}
	DiyFp(uint64_t f_, int e_) : f(f_), e(e_) {}
	explicit DiyFp(double d)
	{
		uint64_t u;
		memcpy(&u, &d, 8);
		const int biased = (int)((u >> 52) & 0x7ff);
		const uint64_t frac = u & 0xfffffffffffffull;
		if (biased) { f = frac | (1ull << 52); e = biased - 0x3ff - 52; }
		else { f = frac; e = 1 - 0x3ff - 52; }
	}
	DiyFp times(const DiyFp &r) const // upper 64 bits of the product, rounded on the 65th
	{
		const unsigned __int128 p = (unsigned __int128)f * r.f;
		uint64_t h = (uint64_t)(p >> 64);
		if ((uint64_t)p & (1ull << 63)) ++h;
		return DiyFp(h, e + r.e + 64);
	}
	DiyFp normalized() const
	{
		const int s = __builtin_clzll(f);
		return DiyFp(f << s, e - s);
	}
	void boundaries(DiyFp *minus, DiyFp *plus) const
	{
		DiyFp pl((f << 1) + 1, e - 1);
		while (!(pl.f & (1ull << 53))) { pl.f <<= 1; --pl.e; }
		pl.f <<= 10;
		pl.e -= 10;
		DiyFp mi = f == (1ull << 52) ? DiyFp((f << 2) - 1, e - 2) : DiyFp((f << 1) - 1, e - 1);
		mi.f <<= mi.e - pl.e;
		mi.e = pl.e;
		*minus = mi;
		*plus = pl;
	}
};

// c_k for k = -348, -340, ..., 340: the 64-bit significand of 10^k rounded to nearest and its binary exponent,
// computed exactly with a little big-integer arithmetic the first time they are needed
const DiyFp *cached_powers()
{
	static const std::vector<DiyFp> table = [] {
		typedef std::vector<uint32_t> Big; // little-endian base 2^32
		auto mul_small = [](Big &x, uint32_t m) {
			uint64_t carry = 0;
			for (uint32_t &w : x) { const uint64_t t = (uint64_t)w * m + carry; w = (uint32_t)t; carry = t >> 32; }
			if (carry) x.push_back((uint32_t)carry);
		};
		auto div_small = [](Big &x, uint32_t v) {
			uint64_t rem = 0;
			for (size_t i = x.size(); i-- > 0;) { const uint64_t t = (rem << 32) | x[i]; x[i] = (uint32_t)(t / v); rem = t % v; }
			while (x.size() > 1 && x.back() == 0) x.pop_back();
		};
		auto bit_length = [](const Big &x) { return (int)(x.size() - 1) * 32 + (32 - __builtin_clz(x.back())); };
		auto bit = [](const Big &x, int i) { return i < 0 ? 0u : (x[(size_t)i / 32] >> (i % 32)) & 1u; };
		std::vector<DiyFp> t;
		for (int k = -348; k <= 340; k += 8) {
			Big n(1, 1u);
			int scale = 0; // value = n * 2^-scale
			if (k >= 0)
				for (int i = 0; i < k; ++i) mul_small(n, 10);
			else {
				scale = 64 + 64 + (int)(-k * 3.33) + 8;
				n.assign((size_t)scale / 32 + 1, 0u);
				n[(size_t)scale / 32] = 1u << (scale % 32);
				for (int i = 0; i < -k; ++i) div_small(n, 10); // nested floor divisions = floor(2^scale / 10^-k)
			}
			const int len = bit_length(n);
			uint64_t f = 0;
			for (int i = 0; i < 64; ++i) f = (f << 1) | bit(n, len - 1 - i);
			int e = len - 64 - scale;
			if (bit(n, len - 65)) {
				if (++f == 0) { f = 1ull << 63; ++e; }
			}
			t.push_back(DiyFp(f, e));
		}
		return t;
	}();
	return table.data();
}

// digits of a positive finite double: buffer[0..*length) and *k with value = digits * 10^k (dtoa.h Grisu2 + DigitGen)
void grisu2(double value, char *buffer, int *length, int *k)
{
	static const uint32_t kPow10[] = {1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000, 1000000000};
	const DiyFp v(value);
	DiyFp w_m, w_p;
	v.boundaries(&w_m, &w_p);
	// the cached power that brings w_p's exponent into the digit-generation window
	const double dk = (-61 - w_p.e) * 0.30102999566398114 + 347;
	int kk = (int)dk;
	if (dk - kk > 0.0) ++kk;
	const unsigned index = (unsigned)((kk >> 3) + 1);
	*k = -(-348 + (int)(index << 3));
	const DiyFp c = cached_powers()[index];
	const DiyFp W = v.normalized().times(c);
	DiyFp Wp = w_p.times(c), Wm = w_m.times(c);
	++Wm.f;
	--Wp.f;
	uint64_t delta = Wp.f - Wm.f;
	const uint64_t one_f = 1ull << -Wp.e, wp_w = Wp.f - W.f;
	uint32_t p1 = (uint32_t)(Wp.f >> -Wp.e);
	uint64_t p2 = Wp.f & (one_f - 1);
	auto weed = [&](uint64_t rest, uint64_t ten_kappa, uint64_t dist) { // move the last digit towards the value
		while (rest < dist && delta - rest >= ten_kappa && (rest + ten_kappa < dist || dist - rest > rest + ten_kappa - dist)) {
			--buffer[*length - 1];
			rest += ten_kappa;
		}
	};
	int kappa = 9;
	for (int i = 1; i < 9; ++i)
		if (p1 < kPow10[i]) { kappa = i; break; }
	*length = 0;
	while (kappa > 0) {
		const uint32_t unit = kPow10[kappa - 1];
		const uint32_t d = p1 / unit;
		p1 %= unit;
		if (d || *length) buffer[(*length)++] = (char)('0' + d);
		--kappa;
		const uint64_t rest = ((uint64_t)p1 << -Wp.e) + p2;
		if (rest <= delta) {
			*k += kappa;
			weed(rest, (uint64_t)kPow10[kappa] << -Wp.e, wp_w);
			return;
		}
	}
	for (;;) {
		p2 *= 10;
		delta *= 10;
		const char d = (char)(p2 >> -Wp.e);
		if (d || *length) buffer[(*length)++] = (char)('0' + d);
		p2 &= one_f - 1;
		--kappa;
		if (p2 < delta) {
			*k += kappa;
			const int index10 = -kappa;
			weed(p2, one_f, wp_w * (index10 < 9 ? kPow10[index10] : 0));
			return;
		}
	}
}

// rapidjson Writer::WriteDouble -> internal::dtoa: Grisu2 digits, then Prettify's choice of fixed or exponent form
std::string fmt_double(double v)
{
	if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
	char buf[32];
	int length = 0, k = 0;
	grisu2(std::fabs(v), buf, &length, &k);
	const std::string digits(buf, (size_t)length);
	const int kk = length + k;
	std::string out = v < 0 ? "-" : "";
	if (0 <= k && kk <= 21) out += digits + std::string((size_t)k, '0') + ".0";
	else if (0 < kk && kk <= 21) out += digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk);
	else if (-6 < kk && kk <= 0) out += "0." + std::string((size_t)(-kk), '0') + digits;
	else {
		out += digits.substr(0, 1);
		if (length > 1) out += "." + digits.substr(1);
		out += "e";
		int e = kk - 1;
		if (e < 0) { out += "-"; e = -e; }
		out += std::to_string(e);
	}
	return out;
}

std::string fmt_string(const char *s)
{
	std::string o = "\"";
	for (const unsigned char *c = (const unsigned char *)s; *c; ++c) {
		switch (*c) {
		case '"': o += "\\\""; break;
		case '\\': o += "\\\\"; break;
		case '\b': o += "\\b"; break;
		case '\f': o += "\\f"; break;
		case '\n': o += "\\n"; break;
		case '\r': o += "\\r"; break;
		case '\t': o += "\\t"; break;
		default:
			if (*c < 0x20) {
				char u[8];
				snprintf(u, sizeof(u), "\\u%04X", *c);
				o += u;
			} else
				o += (char)*c;
		}
	}
	return o + "\"";
}

} // namespace

using adypt::fail;

extern "C" {

int adypt_config_set_default(adypt_instance_config *c)
{
	return adypt::guarded([&]() -> int {
	if (!c) return fail(ADYPT_EINVAL, "config is NULL");
	memset(c, 0, sizeof(*c));
	c->width = 1280; // InstanceConfig.hpp:37
	c->height = 720;
	c->bvh.max_spatial_depth = 48; // :17
	c->bvh.triangle_sah = 0.3f;
	c->bvh.node_sah = 1.0f;
	c->pt.invocation_size = 8; // :24-27
	c->pt.stack_size = 12;
	c->pt.max_bounce = 5;
	c->pt.subpixel = 8;
	c->pt.tmp_lifetime = 16;
	c->pt.ray_tmin = 0.0001f;
	c->pt.clamp = 4.0f; // m_sun has no default in the reference; zero here
	c->cam.speed = 1.0f; // :32-33
	c->cam.mouse_sensitive = 0.3f;
	c->cam.fov = 45.0f;
	return ADYPT_OK;
	});
}

int adypt_config_load(const char *path, adypt_instance_config *c)
{
	return adypt::guarded([&]() -> int {
	if (!path || !c) return fail(ADYPT_EINVAL, "NULL argument");
	std::ifstream in(path);
	if (!in.is_open()) return fail(ADYPT_EIO, std::string("cannot open ") + path);
	std::stringstream ss;
	ss << in.rdbuf();
	const std::string src = ss.str();
	JValue doc;
	Parser ps{src.data(), src.data() + src.size()};
	ps.value(doc);
	ps.ws();
	if (!ps.ok || ps.p != ps.end || doc.kind != JValue::Object) return fail(ADYPT_EINVAL, "[PARSER]ERR: Failed to parse json");
	adypt_instance_config out;
	adypt_config_set_default(&out);

#define NEED_UINT(obj, key, where, dst)                                          \
	{                                                                            \
		const JValue *v = (obj)->find(key);                                      \
		if (!is_uint(v)) return fail(ADYPT_EINVAL, undefined("Uint", key, where)); \
		dst = (int32_t)v->u;                                                     \
	}
#define NEED_FLOAT(obj, key, where, dst)                                            \
	{                                                                               \
		const JValue *v = (obj)->find(key);                                         \
		if (!is_float(v)) return fail(ADYPT_EINVAL, undefined("Float", key, where)); \
		dst = (float)v->d;                                                          \
	}
#define NEED_STRING(obj, key, where, dst)                                                                    \
	{                                                                                                        \
		const JValue *v = (obj)->find(key);                                                                  \
		if (!v || v->kind != JValue::String) return fail(ADYPT_EINVAL, undefined("String", key, where));    \
		if (v->s.size() >= sizeof(dst)) return fail(ADYPT_ERANGE, std::string(key) + " is too long");       \
		snprintf(dst, sizeof(dst), "%s", v->s.c_str());                                                      \
	}
#define NEED_OBJECT(key, var)                                                                          \
	const JValue *var = doc.find(key);                                                                 \
	if (!var || var->kind != JValue::Object) return fail(ADYPT_EINVAL, undefined("Object", key, "document"));
#define NEED_VEC3(obj, key, where, dst)                                                                               \
	{                                                                                                                 \
		const JValue *v = (obj)->find(key);                                                                           \
		if (!v || v->kind != JValue::Array) return fail(ADYPT_EINVAL, undefined("Array", key, where));               \
		if (v->arr.size() != 3) return fail(ADYPT_EINVAL, std::string("[PARSER]ERR: size of \"") + key + "\" array is not 3"); \
		for (int i = 0; i < 3; ++i) {                                                                                 \
			if (!is_float(&v->arr[(size_t)i])) return fail(ADYPT_EINVAL, undefined("Float", key, where));             \
			dst[i] = (float)v->arr[(size_t)i].d;                                                                      \
		}                                                                                                             \
	}

	NEED_UINT(&doc, "width", "document", out.width)
	NEED_UINT(&doc, "height", "document", out.height)
	NEED_OBJECT("scene", scn)
	NEED_STRING(scn, "filename", "scn_obj", out.obj_filename)
	NEED_OBJECT("pathTracer", pt)
	NEED_UINT(pt, "invocationSize", "pt_obj", out.pt.invocation_size)
	NEED_UINT(pt, "stackSize", "pt_obj", out.pt.stack_size)
	NEED_UINT(pt, "maxBounce", "pt_obj", out.pt.max_bounce)
	NEED_UINT(pt, "subpixel", "pt_obj", out.pt.subpixel)
	NEED_UINT(pt, "tmpLifetime", "pt_obj", out.pt.tmp_lifetime)
	NEED_FLOAT(pt, "rayTMin", "pt_obj", out.pt.ray_tmin)
	NEED_FLOAT(pt, "clamp", "pt_obj", out.pt.clamp)
	NEED_VEC3(pt, "sun", "pt_obj", out.pt.sun)
	NEED_OBJECT("bvh", bvh)
	NEED_STRING(bvh, "filename", "bvh_obj", out.bvh_filename)
	NEED_UINT(bvh, "maxSpatialDepth", "bvh_obj", out.bvh.max_spatial_depth)
	NEED_FLOAT(bvh, "triangleSAH", "bvh_obj", out.bvh.triangle_sah)
	NEED_FLOAT(bvh, "nodeSAH", "bvh_obj", out.bvh.node_sah)
	NEED_OBJECT("camera", cam)
	NEED_FLOAT(cam, "speed", "cam_obj", out.cam.speed)
	NEED_FLOAT(cam, "mouseSensitive", "cam_obj", out.cam.mouse_sensitive)
	NEED_FLOAT(cam, "fov", "cam_obj", out.cam.fov)
	NEED_FLOAT(cam, "yaw", "cam_obj", out.cam.yaw)
	NEED_FLOAT(cam, "pitch", "cam_obj", out.cam.pitch)
	NEED_VEC3(cam, "position", "cam_obj", out.cam.position)
	*c = out;
	return ADYPT_OK;
	});
}

int adypt_config_format_double(double value, char out[32])
{
	return adypt::guarded([&]() -> int {
	if (!out) return fail(ADYPT_EINVAL, "out is NULL");
	if (!std::isfinite(value)) return fail(ADYPT_EINVAL, "NaN and infinity have no JSON text");
	const std::string t = fmt_double(value);
	if (t.size() >= 32) return fail(ADYPT_ERANGE, "number text too long");
	memcpy(out, t.c_str(), t.size() + 1);
	return ADYPT_OK;
	});
}

int adypt_config_to_json(const adypt_instance_config *c, char *buf, uint64_t cap, uint64_t *needed)
{
	return adypt::guarded([&]() -> int {
	if (!c) return fail(ADYPT_EINVAL, "config is NULL");
	std::string o;
	const std::string I1 = "    ", I2 = "        ", I3 = "            ";
	auto vec3 = [&](const char *key, const float *v) {
		o += I2 + "\"" + key + "\": [\n";
		for (int i = 0; i < 3; ++i) o += I3 + fmt_double((double)v[i]) + (i < 2 ? ",\n" : "\n");
		o += I2 + "]";
	};
	o += "{\n";
	o += I1 + "\"width\": " + std::to_string(c->width) + ",\n";
	o += I1 + "\"height\": " + std::to_string(c->height) + ",\n";
	o += I1 + "\"scene\": {\n" + I2 + "\"filename\": " + fmt_string(c->obj_filename) + "\n" + I1 + "},\n";
	o += I1 + "\"pathTracer\": {\n";
	o += I2 + "\"invocationSize\": " + std::to_string(c->pt.invocation_size) + ",\n";
	o += I2 + "\"stackSize\": " + std::to_string(c->pt.stack_size) + ",\n";
	o += I2 + "\"maxBounce\": " + std::to_string(c->pt.max_bounce) + ",\n";
	o += I2 + "\"subpixel\": " + std::to_string(c->pt.subpixel) + ",\n";
	o += I2 + "\"tmpLifetime\": " + std::to_string(c->pt.tmp_lifetime) + ",\n";
	o += I2 + "\"rayTMin\": " + fmt_double((double)c->pt.ray_tmin) + ",\n";
	o += I2 + "\"clamp\": " + fmt_double((double)c->pt.clamp) + ",\n";
	vec3("sun", c->pt.sun);
	o += "\n" + I1 + "},\n";
	o += I1 + "\"bvh\": {\n";
	o += I2 + "\"filename\": " + fmt_string(c->bvh_filename) + ",\n";
	o += I2 + "\"maxSpatialDepth\": " + std::to_string(c->bvh.max_spatial_depth) + ",\n";
	o += I2 + "\"triangleSAH\": " + fmt_double((double)c->bvh.triangle_sah) + ",\n";
	o += I2 + "\"nodeSAH\": " + fmt_double((double)c->bvh.node_sah) + "\n" + I1 + "},\n";
	o += I1 + "\"camera\": {\n";
	o += I2 + "\"speed\": " + fmt_double((double)c->cam.speed) + ",\n";
	o += I2 + "\"mouseSensitive\": " + fmt_double((double)c->cam.mouse_sensitive) + ",\n";
	o += I2 + "\"fov\": " + fmt_double((double)c->cam.fov) + ",\n";
	o += I2 + "\"yaw\": " + fmt_double((double)c->cam.yaw) + ",\n";
	o += I2 + "\"pitch\": " + fmt_double((double)c->cam.pitch) + ",\n";
	vec3("position", c->cam.position);
	o += "\n" + I1 + "}\n}";
	if (needed) *needed = o.size() + 1;
	if (buf) {
		if (cap < o.size() + 1) return fail(ADYPT_ERANGE, "buffer too small for the JSON text");
		memcpy(buf, o.c_str(), o.size() + 1);
	}
	return ADYPT_OK;
	});
}

int adypt_config_save(const adypt_instance_config *c, const char *path)
{
	return adypt::guarded([&]() -> int {
	if (!c || !path) return fail(ADYPT_EINVAL, "NULL argument");
	uint64_t need = 0;
	adypt_config_to_json(c, nullptr, 0, &need);
	std::string s((size_t)need, '\0');
	adypt_config_to_json(c, &s[0], need, nullptr);
	std::ofstream out(path);
	if (!out.is_open()) return fail(ADYPT_EIO, std::string("cannot write ") + path);
	out << s.c_str();
	return out.good() ? ADYPT_OK : fail(ADYPT_EIO, std::string("cannot write ") + path);
	});
}

} // extern "C"
