// Texture file decoding for diffuse maps (map_Kd): the stand-in for stbi_load(filename, &w, &h, &n, 3) as
// OglScene::load_texture calls it (src/Tracer/OglScene.cpp:12-43). Output: tightly packed RGB8, top row first.
// Formats: PNG (all colour types and bit depths, Adam7 interlace; zlib does the inflate), JPEG (jpeg_decode.cpp),
// BMP (uncompressed, 4/8-bit palette, 16/24/32-bit) and TGA (true-colour, 15/16-bit, grey, grey+alpha,
// colour-mapped; raw or RLE, either origin). Conversions to 3 channels follow
// stb_image's rules (grey replicated, alpha dropped, 16-bit samples truncated to their high byte, 1/2/4-bit grey
// scaled by 255/85/17). GIF, PSD, PIC, PGM / PPM and HDR live in image_decode_more.cpp. A file none of them accepts fails
// to load, which the reference handles by giving the material texture index -1 (OglScene.cpp:27-32).
#include <zlib.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "host_scene.h"

namespace adypt {
namespace host {

namespace {

bool read_file(const char *path, std::vector<uint8_t> *out)
{
	FILE *f = fopen(path, "rb");
	if (!f) return false;
	uint8_t buf[1 << 16];
	size_t n;
	while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out->insert(out->end(), buf, buf + n);
	fclose(f);
	return true;
}

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline int paeth(int a, int b, int c)
{
	const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
	if (pa <= pb && pa <= pc) return a;
	return pb <= pc ? b : c;
}

// undo PNG filtering of one (sub)image in place; returns false on a bad filter byte
bool unfilter(uint8_t *data, uint32_t rows, size_t row_bytes, int bpp)
{
	const size_t stride = row_bytes + 1;
	for (uint32_t y = 0; y < rows; ++y) {
		uint8_t *cur = data + y * stride + 1;
		const uint8_t *prev = y ? data + (y - 1) * stride + 1 : nullptr;
		const int ft = data[y * stride];
		for (size_t i = 0; i < row_bytes; ++i) {
			const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= (size_t)bpp) ? prev[i - bpp] : 0;
			int v = cur[i];
			switch (ft) {
			case 0: break;
			case 1: v += a; break;
			case 2: v += b; break;
			case 3: v += (a + b) >> 1; break;
			case 4: v += paeth(a, b, c); break;
			default: return false;
			}
			cur[i] = (uint8_t)v;
		}
	}
	return true;
}

bool decode_png(const std::vector<uint8_t> &file, DecodedImage *img)
{
	static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
	if (file.size() < 33 || memcmp(file.data(), sig, 8) != 0) return false;
	size_t pos = 8;
	uint32_t w = 0, h = 0;
	int depth = 0, ctype = 0, interlace = 0;
	std::vector<uint8_t> idat, palette;
	bool seen_ihdr = false;
	while (pos + 12 <= file.size()) {
		const uint32_t len = be32(&file[pos]);
		const uint8_t *type = &file[pos + 4], *body = &file[pos + 8];
		if (pos + 12 + (size_t)len > file.size()) return false;
		if (!memcmp(type, "IHDR", 4)) {
			if (len != 13) return false;
			w = be32(body);
			h = be32(body + 4);
			depth = body[8];
			ctype = body[9];
			interlace = body[12];
			if (body[10] != 0 || body[11] != 0 || interlace > 1) return false;
			seen_ihdr = true;
		} else if (!memcmp(type, "PLTE", 4)) palette.assign(body, body + len);
		else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
		else if (!memcmp(type, "IEND", 4)) break;
		pos += 12 + (size_t)len;
	}
	if (!seen_ihdr || w == 0 || h == 0 || w > (1u << 24) || h > (1u << 24)) return false;
	int channels;
	switch (ctype) {
	case 0: channels = 1; break;
	case 2: channels = 3; break;
	case 3: channels = 1; break;
	case 4: channels = 2; break;
	case 6: channels = 4; break;
	default: return false;
	}
	if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) return false;
	if (ctype == 3 && (depth == 16 || palette.empty())) return false;
	const int bits_pp = channels * depth, bpp = (bits_pp + 7) / 8;

	// pass geometry: one pass for plain images, seven for Adam7
	struct Pass { uint32_t x0, y0, dx, dy, w, h; size_t row_bytes, offset; };
	std::vector<Pass> passes;
	size_t total = 0;
	if (!interlace) passes.push_back(Pass{0, 0, 1, 1, w, h, 0, 0});
	else {
		static const uint32_t X0[7] = {0, 4, 0, 2, 0, 1, 0}, Y0[7] = {0, 0, 4, 0, 2, 0, 1}, DX[7] = {8, 8, 4, 4, 2, 2, 1}, DY[7] = {8, 8, 8, 4, 4, 2, 2};
		for (int p = 0; p < 7; ++p) {
			const uint32_t pw = (w - X0[p] + DX[p] - 1) / DX[p], ph = (h - Y0[p] + DY[p] - 1) / DY[p];
			if (w > X0[p] && h > Y0[p] && pw && ph) passes.push_back(Pass{X0[p], Y0[p], DX[p], DY[p], pw, ph, 0, 0});
		}
	}
	for (Pass &p : passes) {
		p.row_bytes = ((size_t)p.w * bits_pp + 7) / 8;
		p.offset = total;
		total += (p.row_bytes + 1) * p.h;
	}
	std::vector<uint8_t> raw(total);
	uLongf got = (uLongf)total;
	const int zr = uncompress(raw.data(), &got, idat.data(), (uLong)idat.size());
	if ((zr != Z_OK && zr != Z_BUF_ERROR) || got < total) return false;

	img->width = (int)w;
	img->height = (int)h;
	img->rgb.assign((size_t)w * h * 3, 0);
	for (const Pass &p : passes) {
		uint8_t *d = raw.data() + p.offset;
		if (!unfilter(d, p.h, p.row_bytes, bpp)) return false;
		for (uint32_t y = 0; y < p.h; ++y) {
			const uint8_t *row = d + y * (p.row_bytes + 1) + 1;
			for (uint32_t x = 0; x < p.w; ++x) {
				uint8_t s[4] = {0, 0, 0, 0}; // up to four 8-bit samples of this pixel
				if (depth == 8) for (int c = 0; c < channels; ++c) s[c] = row[(size_t)x * channels + c];
				else if (depth == 16) for (int c = 0; c < channels; ++c) s[c] = row[((size_t)x * channels + c) * 2]; // high byte
				else {
					const size_t bit = (size_t)x * depth;
					const int v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
					s[0] = (uint8_t)(ctype == 3 ? v : v * (depth == 1 ? 255 : depth == 2 ? 85 : 17));
				}
				uint8_t *o = &img->rgb[(((size_t)p.y0 + (size_t)y * p.dy) * w + p.x0 + (size_t)x * p.dx) * 3];
				if (ctype == 3) {
					const size_t i = (size_t)s[0] * 3;
					if (i + 2 >= palette.size()) return false;
					o[0] = palette[i]; o[1] = palette[i + 1]; o[2] = palette[i + 2];
				} else if (channels <= 2) o[0] = o[1] = o[2] = s[0];
				else { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; }
			}
		}
	}
	return true;
}

// ---- TGA (stb_image.h:5230-5560 is the behaviour to match): true-colour 24/32 bpp, 15/16-bit 5-5-5, 8-bit grey,
// 16-bit grey+alpha, colour-mapped with 8- or 16-bit indices and any of those entry formats; raw or RLE; bottom-up
// unless descriptor bit 5 is set (bit 4, right-to-left, is ignored there, so it is here). Reads past the end of the
// file yield zeros, as stb_image's byte reader does.
bool decode_tga(const std::vector<uint8_t> &f, DecodedImage *img)
{
	size_t pos = 0;
	auto u8 = [&]() -> int { return pos < f.size() ? f[pos++] : 0; };
	auto u16 = [&]() -> int { const int a = u8(); return a | (u8() << 8); };
	const int id_len = u8(), indexed = u8();
	int type = u8();
	const int pal_start = u16(), pal_len = u16(), pal_bits = u8();
	u16(); u16(); // x / y origin
	const int w = u16(), h = u16(), bpp = u8(), desc = u8();
	// the acceptance test (stbi__tga_test)
	if (indexed > 1) return false;
	if (indexed) {
		if (type != 1 && type != 9) return false;
		if (pal_bits != 8 && pal_bits != 15 && pal_bits != 16 && pal_bits != 24 && pal_bits != 32) return false;
		if (bpp != 8 && bpp != 16) return false;
	} else if (type != 2 && type != 3 && type != 10 && type != 11)
		return false;
	if (w < 1 || h < 1) return false;
	if (bpp != 8 && bpp != 15 && bpp != 16 && bpp != 24 && bpp != 32) return false;
	const bool rle = type >= 8;
	if (rle) type -= 8;
	// bytes per decoded pixel: 1 grey, 2 grey+alpha, 3 rgb (15/16-bit sources are expanded to 3), 4 rgba
	const int bits = indexed ? pal_bits : bpp;
	bool rgb16 = false;
	int comp;
	if (bits == 8) comp = 1;
	else if (bits == 16 && !indexed && type == 3) comp = 2;
	else if (bits == 15 || bits == 16) { comp = 3; rgb16 = true; }
	else comp = bits / 8;
	if ((uint64_t)w * (uint64_t)h * 4u > 0x7fffffffu) return false;
	auto read_rgb16 = [&](uint8_t *out) {
		const int px = u16();
		out[0] = (uint8_t)((((px >> 10) & 31) * 255) / 31);
		out[1] = (uint8_t)((((px >> 5) & 31) * 255) / 31);
		out[2] = (uint8_t)(((px & 31) * 255) / 31);
	};
	pos += (size_t)id_len;
	std::vector<uint8_t> palette;
	if (indexed) {
		pos += (size_t)pal_start; // sic: the first-entry index is skipped as a byte count
		palette.assign((size_t)pal_len * comp + 4, 0);
		for (int i = 0; i < pal_len; ++i) {
			if (rgb16) read_rgb16(&palette[(size_t)i * 3]);
			else for (int j = 0; j < comp; ++j) palette[(size_t)i * comp + j] = (uint8_t)u8();
		}
	}
	const size_t npix = (size_t)w * h;
	std::vector<uint8_t> data(npix * comp);
	uint8_t px[4] = {0, 0, 0, 0};
	int run = 0;
	bool repeating = false;
	for (size_t i = 0; i < npix; ++i) {
		bool read_pixel = true;
		if (rle) {
			if (run == 0) {
				const int c = u8();
				run = 1 + (c & 127);
				repeating = (c >> 7) != 0;
			} else if (repeating)
				read_pixel = false;
		}
		if (read_pixel) {
			if (indexed) {
				int idx = bpp == 8 ? u8() : u16();
				if (idx >= pal_len) idx = 0;
				for (int j = 0; j < comp; ++j) px[j] = palette[(size_t)idx * comp + j];
			} else if (rgb16)
				read_rgb16(px);
			else
				for (int j = 0; j < comp && j < 4; ++j) px[j] = (uint8_t)u8();
		}
		memcpy(&data[i * comp], px, (size_t)comp);
		--run;
	}
	img->width = w;
	img->height = h;
	img->rgb.assign(npix * 3, 0);
	const bool bottom_up = ((desc >> 5) & 1) == 0;
	for (int y = 0; y < h; ++y) {
		const uint8_t *src = &data[(size_t)(bottom_up ? h - 1 - y : y) * w * comp];
		uint8_t *dst = &img->rgb[(size_t)y * w * 3];
		for (int x = 0; x < w; ++x, src += comp, dst += 3) {
			if (comp <= 2) dst[0] = dst[1] = dst[2] = src[0];                    // grey (+alpha): replicated
			else if (rgb16) { dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; } // already R,G,B
			else { dst[0] = src[2]; dst[1] = src[1]; dst[2] = src[0]; }            // B,G,R(,A) on disk
		}
	}
	return true;
}

// ---- BMP (stb_image.h:4920-5230 is the behaviour to match: BITMAPCOREHEADER / INFOHEADER / V3 (56) / V4 / V5 headers;
// 4- and 8-bit palettes, 16-bit 5-5-5 or bit-field masks, 24-bit, 32-bit; bottom-up or top-down; 1-bit and RLE files
// are rejected there, so they are here)
struct ByteReader {
	const uint8_t *p, *end;
	int u8() { return p < end ? *p++ : 0; }
	uint32_t u16() { const uint32_t a = (uint32_t)u8(); return a | ((uint32_t)u8() << 8); }
	uint32_t u32() { const uint32_t a = u16(); return a | (u16() << 16); }
	void skip(long n) { p = (n < 0 || n > end - p) ? (n < 0 ? p : end) : p + n; }
};

inline int top_bit(uint32_t z)
{
	int n = -1;
	while (z) { ++n; z >>= 1; }
	return n;
}
// a masked field moved so that its top bit lands on bit 7, then widened to 8 bits by repeating it
inline int expand_field(int v, int shift, int bits)
{
	v = shift < 0 ? v << -shift : v >> shift;
	int result = v;
	for (int z = bits; z < 8; z += bits) result += v >> z;
	return result;
}

bool decode_bmp(const std::vector<uint8_t> &f, DecodedImage *img)
{
	ByteReader r{f.data(), f.data() + f.size()};
	if (r.u8() != 'B' || r.u8() != 'M') return false;
	r.u32(); r.u16(); r.u16();
	const long offset = (long)r.u32();
	const long hsz = (long)r.u32();
	if (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124) return false;
	int32_t w, h;
	if (hsz == 12) { w = (int32_t)r.u16(); h = (int32_t)r.u16(); }
	else { w = (int32_t)r.u32(); h = (int32_t)r.u32(); }
	if (r.u16() != 1) return false;
	const int bpp = (int)r.u16();
	if (bpp == 1) return false;
	uint32_t mr = 0, mg = 0, mb = 0, ma = 0;
	if (hsz != 12) {
		const uint32_t compress = r.u32();
		if (compress == 1 || compress == 2) return false; // RLE
		for (int i = 0; i < 5; ++i) r.u32();
		if (hsz == 40 || hsz == 56) {
			if (hsz == 56) for (int i = 0; i < 4; ++i) r.u32();
			if (bpp == 16 || bpp == 32) {
				if (compress == 0) {
					if (bpp == 32) { mr = 0xffu << 16; mg = 0xffu << 8; mb = 0xffu; ma = 0xffu << 24; }
					else { mr = 31u << 10; mg = 31u << 5; mb = 31u; }
				} else if (compress == 3) {
					mr = r.u32(); mg = r.u32(); mb = r.u32();
					if (mr == mg && mg == mb) return false;
				} else
					return false;
			}
		} else {
			mr = r.u32(); mg = r.u32(); mb = r.u32(); ma = r.u32();
			for (int i = 0; i < 13; ++i) r.u32();
			if (hsz == 124) for (int i = 0; i < 4; ++i) r.u32();
		}
	}
	const bool bottom_up = h > 0;
	if (h < 0) h = -h;
	if (w <= 0 || h <= 0 || (uint64_t)w * (uint64_t)h * 3u > 0x7fffffffu) return false;
	long psize = 0;
	if (hsz == 12) { if (bpp < 24) psize = (offset - 14 - 24) / 3; }
	else if (bpp < 16) psize = (offset - 14 - hsz) >> 2;
	img->width = w;
	img->height = h;
	img->rgb.assign((size_t)w * h * 3, 0);
	uint8_t *out = img->rgb.data();
	size_t z = 0;
	if (bpp < 16) {
		if (psize == 0 || psize > 256 || (bpp != 4 && bpp != 8)) return false;
		uint8_t pal[256][3];
		memset(pal, 0, sizeof(pal));
		for (long i = 0; i < psize; ++i) {
			pal[i][2] = (uint8_t)r.u8(); pal[i][1] = (uint8_t)r.u8(); pal[i][0] = (uint8_t)r.u8();
			if (hsz != 12) r.u8();
		}
		r.skip(offset - 14 - hsz - psize * (hsz == 12 ? 3 : 4));
		const int row_bytes = bpp == 4 ? (w + 1) >> 1 : w, pad = (-row_bytes) & 3;
		for (int j = 0; j < h; ++j) {
			for (int i = 0; i < w; i += 2) {
				int v = r.u8(), v2 = 0;
				if (bpp == 4) { v2 = v & 15; v >>= 4; }
				out[z++] = pal[v][0]; out[z++] = pal[v][1]; out[z++] = pal[v][2];
				if (i + 1 == w) break;
				v = bpp == 8 ? r.u8() : v2;
				out[z++] = pal[v][0]; out[z++] = pal[v][1]; out[z++] = pal[v][2];
			}
			r.skip(pad);
		}
	} else {
		if (bpp != 16 && bpp != 24 && bpp != 32) return false;
		r.skip(offset - 14 - hsz);
		const int row_bytes = bpp == 24 ? 3 * w : bpp == 16 ? 2 * w : 0, pad = (-row_bytes) & 3;
		const bool plain = bpp == 24 || (bpp == 32 && mb == 0xffu && mg == 0xff00u && mr == 0x00ff0000u && ma == 0xff000000u);
		int rs = 0, gs = 0, bs = 0, rc = 0, gc = 0, bc = 0;
		if (!plain) {
			if (!mr || !mg || !mb) return false;
			rs = top_bit(mr) - 7; rc = __builtin_popcount(mr);
			gs = top_bit(mg) - 7; gc = __builtin_popcount(mg);
			bs = top_bit(mb) - 7; bc = __builtin_popcount(mb);
		}
		for (int j = 0; j < h; ++j) {
			for (int i = 0; i < w; ++i) {
				if (plain) {
					out[z + 2] = (uint8_t)r.u8(); out[z + 1] = (uint8_t)r.u8(); out[z] = (uint8_t)r.u8();
					if (bpp == 32) r.u8();
				} else {
					const uint32_t v = bpp == 16 ? r.u16() : r.u32();
					out[z] = (uint8_t)(expand_field((int)(v & mr), rs, rc) & 255);
					out[z + 1] = (uint8_t)(expand_field((int)(v & mg), gs, gc) & 255);
					out[z + 2] = (uint8_t)(expand_field((int)(v & mb), bs, bc) & 255);
				}
				z += 3;
			}
			r.skip(pad);
		}
	}
	if (bottom_up)
		for (int j = 0; j < h / 2; ++j) {
			uint8_t *a = out + (size_t)j * w * 3, *b = out + (size_t)(h - 1 - j) * w * 3;
			for (int i = 0; i < w * 3; ++i) { const uint8_t t = a[i]; a[i] = b[i]; b[i] = t; }
		}
	return true;
}

} // namespace

bool decode_image_file(const char *path, DecodedImage *img)
{
	std::vector<uint8_t> file;
	if (!read_file(path, &file)) return false;
	if (decode_png(file, img)) return true;
	if (file.size() > 2 && file[0] == 'B' && file[1] == 'M') return decode_bmp(file, img);
	if (decode_jpeg(file, img)) return true;
	bool recognised = false;
	if (decode_rare_formats(file, img, &recognised)) return true; // GIF, PSD, PIC, PGM / PPM, HDR (image_decode_more.cpp)
	if (recognised) return false;                                 // one of those, but corrupt: stb_image stops there too
	return decode_tga(file, img); // TGA has no magic number: like stb_image, try it last, on the header's plausibility
}

} // namespace host
} // namespace adypt
