// Headless replacement for the viewer's export: a minimal OpenEXR 2 scanline writer (RGB, HALF or FLOAT,
// ZIP compression in 16-line chunks). Stands in for tinyexr's SaveEXR as called by
// OglPathTracer::SaveResult (src/Tracer/OglPathTracer.cpp:199-212), which writes the same kind of file
// (3 channels, fp16 or fp32, ZIP). Written against the public OpenEXR file-layout specification; zlib
// does the deflate.
#include <zlib.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/adypt_b200.h"
#include "guard.h"

namespace {

void put_u32(std::vector<uint8_t> &b, uint32_t v)
{
	for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i)));
}
void put_str(std::vector<uint8_t> &b, const char *s)
{
	while (*s) b.push_back((uint8_t)*s++);
	b.push_back(0);
}
void put_attr(std::vector<uint8_t> &b, const char *name, const char *type, const void *data, uint32_t size)
{
	put_str(b, name);
	put_str(b, type);
	put_u32(b, size);
	const uint8_t *p = (const uint8_t *)data;
	b.insert(b.end(), p, p + size);
}

// binary32 -> binary16 with the rules of the reference's writer (tinyexr's float_to_half_full, dep/tinyexr.h:7160-7195,
// "based on ISPC reference code"), so that an fp16 file holds the same bits: the first dropped bit decides the
// rounding (halves go UP, not to even), float subnormals become zero, any NaN becomes the quiet NaN 0x7e00 (sign
// kept), results of 2^16 and above become infinity, and a rounding carry may walk into the exponent.
uint16_t float_to_half(float f)
{
	uint32_t x;
	memcpy(&x, &f, 4);
	const uint32_t sign = (x >> 16) & 0x8000u, exp8 = (x >> 23) & 0xffu, man = x & 0x7fffffu;
	uint32_t h = 0;
	if (exp8 == 0) h = 0;
	else if (exp8 == 255) h = 0x7c00u | (man ? 0x200u : 0u);
	else {
		const int e = (int)exp8 - 127 + 15;
		if (e >= 31) h = 0x7c00u;
		else if (e <= 0) {
			if (14 - e <= 24) {
				const uint32_t m = man | 0x800000u;
				h = m >> (14 - e);
				if ((m >> (13 - e)) & 1u) ++h;
			}
		} else {
			h = ((uint32_t)e << 10) | (man >> 13);
			if (man & 0x1000u) ++h;
		}
	}
	return (uint16_t)(sign | h);
}

} // namespace

extern "C" int adypt_write_exr(const char *filename, const float *rgb, int32_t width, int32_t height, int32_t save_as_fp16)
{
	return adypt::guarded([&]() -> int {
	if (!filename || !rgb || width <= 0 || height <= 0) return ADYPT_EINVAL;
	const int bpc = save_as_fp16 ? 2 : 4; // bytes per channel sample
	std::vector<uint8_t> hdr;
	put_u32(hdr, 20000630u); // magic 0x76 0x2f 0x31 0x01
	put_u32(hdr, 2u);        // version 2, single-part scanline, no flags
	{
		std::vector<uint8_t> ch;
		const char *names[3] = {"B", "G", "R"}; // channels are stored in alphabetical order
		for (int c = 0; c < 3; ++c) {
			put_str(ch, names[c]);
			put_u32(ch, save_as_fp16 ? 1u : 2u); // HALF = 1, FLOAT = 2
			ch.push_back(0);                     // pLinear
			ch.push_back(0); ch.push_back(0); ch.push_back(0);
			put_u32(ch, 1u); // xSampling
			put_u32(ch, 1u); // ySampling
		}
		ch.push_back(0);
		put_attr(hdr, "channels", "chlist", ch.data(), (uint32_t)ch.size());
	}
	const uint8_t compression = 3; // ZIP_COMPRESSION: zlib over blocks of 16 scanlines
	put_attr(hdr, "compression", "compression", &compression, 1);
	const int32_t window[4] = {0, 0, width - 1, height - 1};
	put_attr(hdr, "dataWindow", "box2i", window, 16);
	put_attr(hdr, "displayWindow", "box2i", window, 16);
	const uint8_t line_order = 0; // INCREASING_Y
	put_attr(hdr, "lineOrder", "lineOrder", &line_order, 1);
	const float one = 1.0f, zero2[2] = {0.0f, 0.0f};
	put_attr(hdr, "pixelAspectRatio", "float", &one, 4);
	put_attr(hdr, "screenWindowCenter", "v2f", zero2, 8);
	const float screen_width = (float)width; // tinyexr writes the image width here (dep/tinyexr.h:11767-11773)
	put_attr(hdr, "screenWindowWidth", "float", &screen_width, 4);
	hdr.push_back(0);

	const int lines_per_chunk = 16;
	const int n_chunks = (height + lines_per_chunk - 1) / lines_per_chunk;
	std::vector<uint64_t> offsets((size_t)n_chunks);
	std::vector<uint8_t> body;
	const size_t line_bytes = (size_t)width * 3u * bpc;
	std::vector<uint8_t> raw(line_bytes * lines_per_chunk), tmp(raw.size());
	std::vector<uint8_t> packed(compressBound((uLong)raw.size()));
	const uint64_t data_start = hdr.size() + (uint64_t)n_chunks * 8u;

	for (int c = 0; c < n_chunks; ++c) {
		const int y0 = c * lines_per_chunk, y1 = (y0 + lines_per_chunk < height) ? y0 + lines_per_chunk : height;
		const size_t raw_size = line_bytes * (size_t)(y1 - y0);
		// scanline layout: for each line, each channel (B, G, R) as a contiguous run of `width` samples
		uint8_t *w = raw.data();
		for (int y = y0; y < y1; ++y)
			for (int ch = 2; ch >= 0; --ch) { // B = rgb[2], G = rgb[1], R = rgb[0]
				const float *src = rgb + (size_t)y * width * 3 + ch;
				if (save_as_fp16)
					for (int x = 0; x < width; ++x) {
						const uint16_t hv = float_to_half(src[(size_t)x * 3]);
						*w++ = (uint8_t)hv;
						*w++ = (uint8_t)(hv >> 8);
					}
				else
					for (int x = 0; x < width; ++x) {
						memcpy(w, src + (size_t)x * 3, 4);
						w += 4;
					}
			}
		// ZIP pre-process: de-interleave even/odd bytes, then byte-wise delta predictor
		{
			uint8_t *t1 = tmp.data(), *t2 = tmp.data() + (raw_size + 1) / 2;
			for (size_t i = 0; i < raw_size; ++i) {
				if (i & 1) *t2++ = raw[i];
				else *t1++ = raw[i];
			}
			int p = tmp[0];
			for (size_t i = 1; i < raw_size; ++i) {
				const int d = (int)tmp[i] - p + (128 + 256);
				p = tmp[i];
				tmp[i] = (uint8_t)d;
			}
		}
		uLongf packed_size = (uLongf)packed.size();
		const int zr = compress2(packed.data(), &packed_size, tmp.data(), (uLong)raw_size, Z_DEFAULT_COMPRESSION);
		const bool use_raw = zr != Z_OK || packed_size >= raw_size; // spec: store uncompressed if not smaller
		offsets[(size_t)c] = data_start + body.size();
		put_u32(body, (uint32_t)y0);
		put_u32(body, (uint32_t)(use_raw ? raw_size : packed_size));
		if (use_raw) body.insert(body.end(), raw.begin(), raw.begin() + (long)raw_size);
		else body.insert(body.end(), packed.begin(), packed.begin() + (long)packed_size);
	}

	FILE *f = fopen(filename, "wb");
	if (!f) return ADYPT_EIO;
	bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
	ok = ok && fwrite(offsets.data(), 8, offsets.size(), f) == offsets.size();
	ok = ok && fwrite(body.data(), 1, body.size(), f) == body.size();
	ok = (fclose(f) == 0) && ok;
	return ok ? ADYPT_OK : ADYPT_EIO;
	});
}
