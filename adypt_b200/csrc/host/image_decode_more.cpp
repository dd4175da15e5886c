// The rarer texture formats stbi_load(filename, &w, &h, &n, 3) accepts (OglScene::load_texture, src/Tracer/OglScene.cpp:24):
// GIF (first frame), PSD (8/16-bit RGB, raw or RLE), Softimage PIC, binary PGM / PPM and Radiance HDR (tone-mapped to 8 bits
// the way stb_image does it). Behaviour to match: the reference's vendored dep/stb_image.h v2.16 -- :6055-6395 GIF,
// :5567-5810 PSD, :5829-6025 PIC, :6776-6885 PNM, :6405-6580 + :1587-1612 HDR -- including its quirks, because the pixels
// it hands to OpenGL are what a render depends on:
//   * GIF: pixels of a transparent colour keep the background colour; a file whose first block is the trailer has no image;
//   * PSD: channels beyond the fourth are ignored, missing ones are 0 (alpha 255), 16-bit samples keep their high byte, and
//     the "white matte" is removed from partly transparent pixels in float arithmetic;
//   * PIC: channels are selected by packet masks; alpha is dropped;
//   * HDR: a scanline that does not start with 2 2 hi lo switches to flat RGBE data for the REST of the image (and restarts
//     at pixel 1 of row 0, stb_image's "this makes no sense" goto); conversion is pow(x, 1/2.2) * 255 + 0.5, truncated.
// Files are read through a model of stb_image's stdio reader (128-byte buffer, feof()-driven end test, zeros after the
// end), so truncated files decode to the same pixels too. Where stb_image itself reads or writes out of bounds (PIC decode
// errors, a 12-bit GIF minimum code size, a zero-width GIF frame at the image corner, short PNM data) the file is
// rejected or the missing bytes are zero; a run-length HDR that ends inside a scanline makes stb_image loop for ever and
// is rejected here; zero-sized images are rejected too.
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "host_scene.h"

namespace adypt {
namespace host {

namespace {

// stbi__context over a FILE (stbi__start_file): stb_image.h:1342-1416
class StbStream {
	const std::vector<uint8_t> &file_;
	size_t fpos_ = 0;     // the FILE's position
	bool feof_ = false;   // feof(): set by a short fread, cleared by fseek
	uint8_t buf_[128];
	int pos_ = 0, end_ = 0;
	bool live_ = true;    // read_from_callbacks

	size_t file_read(uint8_t *dst, size_t n)
	{
		const size_t avail = fpos_ < file_.size() ? file_.size() - fpos_ : 0;
		const size_t c = n < avail ? n : avail;
		if (c) memcpy(dst, file_.data() + fpos_, c);
		fpos_ += c;
		if (c < n) feof_ = true;
		return c;
	}
	void refill()
	{
		const size_t n = file_read(buf_, sizeof(buf_));
		pos_ = 0;
		if (n == 0) { live_ = false; end_ = 1; buf_[0] = 0; }
		else end_ = (int)n;
	}

public:
	explicit StbStream(const std::vector<uint8_t> &file) : file_(file) { refill(); }
	int get8()
	{
		if (pos_ < end_) return buf_[pos_++];
		if (live_) { refill(); return buf_[pos_++]; }
		return 0;
	}
	int get16be() { const int z = get8(); return (z << 8) + get8(); }
	uint32_t get32be() { const uint32_t z = (uint32_t)get16be(); return (z << 16) + (uint32_t)get16be(); }
	int get16le() { const int z = get8(); return z + (get8() << 8); }
	bool at_eof() const
	{
		if (!feof_) return false;
		if (!live_) return true;
		return pos_ >= end_;
	}
	void skip(int n)
	{
		if (n < 0) { pos_ = end_; return; }
		const int blen = end_ - pos_;
		if (blen < n) {
			pos_ = end_;
			fpos_ += (size_t)(n - blen); // fseek(SEEK_CUR): may pass the end, clears the end-of-file flag
			feof_ = false;
			return;
		}
		pos_ += n;
	}
	bool getn(uint8_t *out, int n)
	{
		const int blen = end_ - pos_;
		if (blen < n) {
			memcpy(out, buf_ + pos_, (size_t)blen);
			const size_t c = file_read(out + blen, (size_t)(n - blen));
			pos_ = end_;
			return c == (size_t)(n - blen);
		}
		memcpy(out, buf_ + pos_, (size_t)n);
		pos_ += n;
		return true;
	}
};

// stbi__mad3sizes_valid / mad4: every factor non-negative and the product an int
bool sizes_ok(long long a, long long b, long long c, long long d = 1)
{
	if (a < 0 || b < 0 || c < 0 || d < 0) return false;
	long long p = a;
	for (long long f : {b, c, d}) {
		if (f != 0 && p > (long long)INT_MAX / f) return false;
		p *= f;
	}
	return true;
}

// RGBA rows -> the caller's RGB8 (stbi__convert_format 4 -> 3: alpha dropped)
void store_rgba(const std::vector<uint8_t> &rgba, int w, int h, DecodedImage *img)
{
	img->width = w;
	img->height = h;
	img->rgb.resize((size_t)w * h * 3);
	for (size_t i = 0, n = (size_t)w * h; i < n; ++i) {
		img->rgb[3 * i] = rgba[4 * i];
		img->rgb[3 * i + 1] = rgba[4 * i + 1];
		img->rgb[3 * i + 2] = rgba[4 * i + 2];
	}
}

// ------------------------------------------------------------------------------------------------ PGM / PPM (P5, P6)
bool decode_pnm(const std::vector<uint8_t> &file, DecodedImage *img)
{
	StbStream s(file);
	const char p = (char)s.get8(), t = (char)s.get8();
	if (p != 'P' || (t != '5' && t != '6')) return false;
	const int comp = t == '6' ? 3 : 1;
	char c = (char)s.get8();
	auto is_space = [](char ch) { return ch == ' ' || ch == '\t' || ch == '\n' || ch == '\v' || ch == '\f' || ch == '\r'; };
	auto skip_ws = [&]() {
		for (;;) {
			while (!s.at_eof() && is_space(c)) c = (char)s.get8();
			if (s.at_eof() || c != '#') break;
			while (!s.at_eof() && c != '\n' && c != '\r') c = (char)s.get8();
		}
	};
	bool overflow = false;
	auto integer = [&]() -> int {
		long long v = 0;
		while (!s.at_eof() && c >= '0' && c <= '9') {
			v = v * 10 + (c - '0');
			if (v > INT_MAX) { overflow = true; v = INT_MAX; } // int overflow there: undefined, rejected here
			c = (char)s.get8();
		}
		return (int)v;
	};
	skip_ws();
	const int w = integer();
	skip_ws();
	const int h = integer();
	skip_ws();
	const int maxv = integer();
	if (overflow || maxv > 255) return false;
	if (!sizes_ok(comp, w, h) || w == 0 || h == 0) return false;
	std::vector<uint8_t> data((size_t)comp * w * h, 0);
	s.getn(data.data(), (int)data.size()); // a short file leaves the tail unset there; zero here
	img->width = w;
	img->height = h;
	img->rgb.resize((size_t)w * h * 3);
	for (size_t i = 0, n = (size_t)w * h; i < n; ++i)
		for (int k = 0; k < 3; ++k) img->rgb[3 * i + k] = comp == 3 ? data[3 * i + k] : data[i];
	return true;
}

// ------------------------------------------------------------------------------------------------ GIF (first frame)
struct GifCode { int16_t prefix; uint8_t first, suffix; };

bool decode_gif(const std::vector<uint8_t> &file, DecodedImage *img)
{
	StbStream s(file);
	if (s.get8() != 'G' || s.get8() != 'I' || s.get8() != 'F' || s.get8() != '8') return false;
	const int version = s.get8();
	if (version != '7' && version != '9') return false;
	if (s.get8() != 'a') return false;
	const int gw = s.get16le(), gh = s.get16le(), flags = s.get8(), bgindex = s.get8();
	s.get8(); // aspect ratio
	int transparent = -1, eflags = 0;
	uint8_t pal[256][4], lpal[256][4]; // stored B, G, R, A like stb_image does
	memset(pal, 0, sizeof(pal));
	memset(lpal, 0, sizeof(lpal));
	auto read_table = [&](uint8_t t[256][4], int entries, int transp) {
		for (int i = 0; i < entries; ++i) {
			t[i][2] = (uint8_t)s.get8();
			t[i][1] = (uint8_t)s.get8();
			t[i][0] = (uint8_t)s.get8();
			t[i][3] = transp == i ? 0 : 255;
		}
	};
	if (flags & 0x80) read_table(pal, 2 << (flags & 7), -1);
	if (!sizes_ok(gw, gh, 4) || gw == 0 || gh == 0) return false;
	std::vector<uint8_t> out((size_t)gw * gh * 4);
	for (size_t i = 0, n = (size_t)gw * gh; i < n; ++i) { // background, alpha 0
		out[4 * i] = pal[bgindex][2];
		out[4 * i + 1] = pal[bgindex][1];
		out[4 * i + 2] = pal[bgindex][0];
		out[4 * i + 3] = 0;
	}
	for (;;) {
		const int block = s.get8();
		if (block == 0x21) { // extension
			int len;
			if (s.get8() == 0xF9) { // graphic control
				len = s.get8();
				if (len == 4) {
					eflags = s.get8();
					s.get16le(); // delay
					transparent = s.get8();
				} else {
					s.skip(len);
					continue; // sic: the sub-block loop below is not run
				}
			}
			while ((len = s.get8()) != 0) s.skip(len);
			continue;
		}
		if (block != 0x2C) return false; // 0x3B (trailer before any image) and unknown blocks: no image
		const int x = s.get16le(), y = s.get16le(), w = s.get16le(), h = s.get16le();
		if (x + w > gw || y + h > gh) return false;
		const long long line = (long long)gw * 4;
		const long long start_x = (long long)x * 4, start_y = (long long)y * line, max_x = start_x + (long long)w * 4, max_y = start_y + (long long)h * line;
		long long cur_x = start_x, cur_y = start_y, step;
		int parse;
		const int lflags = s.get8();
		if (lflags & 0x40) { step = 8 * line; parse = 3; } // interlaced
		else { step = line; parse = 0; }
		const uint8_t (*table)[4];
		if (lflags & 0x80) {
			read_table(lpal, 2 << (lflags & 7), (eflags & 1) ? transparent : -1);
			table = lpal;
		} else if (flags & 0x80) {
			if (transparent >= 0 && (eflags & 1)) pal[transparent][3] = 0;
			table = pal;
		} else
			return false;
		// LZW raster
		const int lzw_cs = s.get8();
		if (lzw_cs > 11) return false; // 12 is accepted there and then indexes past the code table
		const int clear = 1 << lzw_cs;
		std::vector<GifCode> codes(4096);
		for (int i = 0; i < 4096; ++i) codes[i] = GifCode{0, 0, 0};
		for (int i = 0; i < clear; ++i) codes[i] = GifCode{-1, (uint8_t)i, (uint8_t)i};
		bool first = true;
		int codesize = lzw_cs + 1, codemask = (1 << codesize) - 1, avail = clear + 2, oldcode = -1, valid_bits = 0, len = 0;
		int32_t bits = 0;
		std::vector<uint16_t> chain;
		chain.reserve(4096);
		auto emit = [&](int code) { // stbi__out_gif_code: the prefix chain root first
			chain.clear();
			for (int c = code; ; c = codes[c].prefix) {
				chain.push_back((uint16_t)c);
				if (codes[c].prefix < 0 || chain.size() > 4096) break;
			}
			for (size_t k = chain.size(); k-- > 0;) {
				if (cur_y >= max_y) continue;
				const uint8_t *c = table[codes[chain[k]].suffix];
				const long long at = cur_x + cur_y;
				if (c[3] >= 128 && at + 4 <= (long long)out.size()) {
					out[(size_t)at] = c[2];
					out[(size_t)at + 1] = c[1];
					out[(size_t)at + 2] = c[0];
					out[(size_t)at + 3] = c[3];
				}
				cur_x += 4;
				if (cur_x >= max_x) {
					cur_x = start_x;
					cur_y += step;
					while (cur_y >= max_y && parse > 0) {
						step = (1ll << parse) * line;
						cur_y = start_y + (step >> 1);
						--parse;
					}
				}
			}
		};
		for (;;) {
			if (valid_bits < codesize) {
				if (len == 0) {
					len = s.get8(); // next data sub-block
					if (len == 0) break;
				}
				--len;
				bits |= (int32_t)((uint32_t)s.get8() << valid_bits);
				valid_bits += 8;
				continue;
			}
			const int code = bits & codemask;
			bits >>= codesize;
			valid_bits -= codesize;
			if (code == clear) {
				codesize = lzw_cs + 1;
				codemask = (1 << codesize) - 1;
				avail = clear + 2;
				oldcode = -1;
				first = false;
			} else if (code == clear + 1) {
				break; // end of stream (what follows in the file is not needed for the first frame)
			} else if (code <= avail) {
				if (first) return false; // no clear code
				if (oldcode >= 0) {
					GifCode *p = &codes[avail++];
					if (avail > 4096) return false;
					p->prefix = (int16_t)oldcode;
					p->first = codes[oldcode].first;
					p->suffix = (code == avail) ? p->first : codes[code].first;
				} else if (code == avail)
					return false;
				emit(code);
				if ((avail & codemask) == 0 && avail <= 0x0FFF) {
					++codesize;
					codemask = (1 << codesize) - 1;
				}
				oldcode = code;
			} else
				return false;
		}
		store_rgba(out, gw, gh, img);
		return true;
	}
}

// ------------------------------------------------------------------------------------------------ PSD
bool psd_rle(StbStream &s, uint8_t *p, int pixel_count)
{
	int count = 0, nleft;
	while ((nleft = pixel_count - count) > 0) {
		int len = s.get8();
		if (len == 128) continue;
		if (len < 128) {
			++len;
			if (len > nleft) return false;
			count += len;
			for (; len; --len, p += 4) *p = (uint8_t)s.get8();
		} else {
			len = 257 - len;
			if (len > nleft) return false;
			const uint8_t val = (uint8_t)s.get8();
			count += len;
			for (; len; --len, p += 4) *p = val;
		}
	}
	return true;
}

bool decode_psd(const std::vector<uint8_t> &file, DecodedImage *img)
{
	StbStream s(file);
	if (s.get32be() != 0x38425053u) return false; // "8BPS"
	if (s.get16be() != 1) return false;
	s.skip(6);
	const int channels = s.get16be();
	if (channels < 0 || channels > 16) return false;
	const int h = (int)s.get32be(), w = (int)s.get32be();
	const int depth = s.get16be();
	if (depth != 8 && depth != 16) return false;
	if (s.get16be() != 3) return false; // RGB colour mode only
	s.skip((int)s.get32be()); // mode data
	s.skip((int)s.get32be()); // image resources
	s.skip((int)s.get32be()); // layer and mask information
	const int compression = s.get16be();
	if (compression > 1) return false;
	if (!sizes_ok(4, w, h) || w == 0 || h == 0) return false;
	const int pixel_count = w * h;
	std::vector<uint8_t> out((size_t)pixel_count * 4, 0);
	if (compression) {
		s.skip((int)(uint32_t)((uint64_t)h * (uint64_t)channels * 2u)); // per-row byte counts (wraps like stb_image's int product)
		for (int ch = 0; ch < 4; ++ch) {
			uint8_t *p = out.data() + ch;
			if (ch >= channels) {
				for (int i = 0; i < pixel_count; ++i, p += 4) *p = ch == 3 ? 255 : 0;
			} else if (!psd_rle(s, p, pixel_count))
				return false;
		}
	} else {
		for (int ch = 0; ch < 4; ++ch) {
			uint8_t *p = out.data() + ch;
			if (ch >= channels) {
				for (int i = 0; i < pixel_count; ++i, p += 4) *p = ch == 3 ? 255 : 0;
			} else if (depth == 16) {
				for (int i = 0; i < pixel_count; ++i, p += 4) *p = (uint8_t)(s.get16be() >> 8);
			} else {
				for (int i = 0; i < pixel_count; ++i, p += 4) *p = (uint8_t)s.get8();
			}
		}
	}
	if (channels >= 4) { // un-premultiply against white, in stb_image's float arithmetic and int -> byte wrap
		for (int i = 0; i < pixel_count; ++i) {
			uint8_t *px = out.data() + 4 * (size_t)i;
			if (px[3] != 0 && px[3] != 255) {
				const float a = px[3] / 255.0f;
				const float ra = 1.0f / a;
				const float inv_a = 255.0f * (1 - ra);
				for (int k = 0; k < 3; ++k) px[k] = (uint8_t)(int)(px[k] * ra + inv_a);
			}
		}
	}
	store_rgba(out, w, h, img);
	return true;
}

// ------------------------------------------------------------------------------------------------ Softimage PIC
bool decode_pic(const std::vector<uint8_t> &file, DecodedImage *img)
{
	{
		StbStream t(file); // stbi__pic_test_core
		static const uint8_t magic[4] = {0x53, 0x80, 0xF6, 0x34};
		for (int i = 0; i < 4; ++i) if (t.get8() != magic[i]) return false;
		for (int i = 0; i < 84; ++i) t.get8();
		static const char pict[4] = {'P', 'I', 'C', 'T'};
		for (int i = 0; i < 4; ++i) if (t.get8() != (uint8_t)pict[i]) return false;
	}
	StbStream s(file);
	for (int i = 0; i < 92; ++i) s.get8();
	const int w = s.get16be(), h = s.get16be();
	if (s.at_eof()) return false;
	if (!sizes_ok(w, h, 4) || w == 0 || h == 0) return false;
	s.get32be(); // ratio
	s.get16be(); // fields
	s.get16be(); // pad
	std::vector<uint8_t> out((size_t)w * h * 4, 0xff);
	struct Packet { uint8_t size, type, channel; } packets[10];
	int num_packets = 0, chained;
	do {
		if (num_packets == 10) return false;
		Packet &p = packets[num_packets++];
		chained = s.get8();
		p.size = (uint8_t)s.get8();
		p.type = (uint8_t)s.get8();
		p.channel = (uint8_t)s.get8();
		if (s.at_eof()) return false;
		if (p.size != 8) return false;
	} while (chained);
	auto readval = [&](int channel, uint8_t *dest) -> bool {
		for (int i = 0, mask = 0x80; i < 4; ++i, mask >>= 1)
			if (channel & mask) {
				if (s.at_eof()) return false;
				dest[i] = (uint8_t)s.get8();
			}
		return true;
	};
	auto copyval = [](int channel, uint8_t *dest, const uint8_t *src) {
		for (int i = 0, mask = 0x80; i < 4; ++i, mask >>= 1)
			if (channel & mask) dest[i] = src[i];
	};
	for (int y = 0; y < h; ++y)
		for (int k = 0; k < num_packets; ++k) {
			const Packet &p = packets[k];
			uint8_t *dest = out.data() + (size_t)y * w * 4;
			if (p.type == 0) { // uncompressed
				for (int x = 0; x < w; ++x, dest += 4)
					if (!readval(p.channel, dest)) return false;
			} else if (p.type == 1) { // pure run-length
				int left = w;
				while (left > 0) {
					uint8_t value[4];
					int count = s.get8();
					if (s.at_eof()) return false;
					if (count > left) count = (uint8_t)left;
					if (!readval(p.channel, value)) return false;
					for (int i = 0; i < count; ++i, dest += 4) copyval(p.channel, dest, value);
					left -= count;
				}
			} else if (p.type == 2) { // mixed run-length
				int left = w;
				while (left > 0) {
					int count = s.get8();
					if (s.at_eof()) return false;
					if (count >= 128) {
						uint8_t value[4];
						if (count == 128) count = s.get16be();
						else count -= 127;
						if (count > left) return false;
						if (!readval(p.channel, value)) return false;
						for (int i = 0; i < count; ++i, dest += 4) copyval(p.channel, dest, value);
					} else {
						++count;
						if (count > left) return false;
						for (int i = 0; i < count; ++i, dest += 4)
							if (!readval(p.channel, dest)) return false;
					}
					left -= count;
				}
			} else
				return false;
		}
	store_rgba(out, w, h, img);
	return true;
}

// ------------------------------------------------------------------------------------------------ Radiance HDR
bool decode_hdr(const std::vector<uint8_t> &file, DecodedImage *img)
{
	StbStream s(file);
	char buffer[1024];
	auto token = [&]() -> char * { // stbi__hdr_gettoken
		int len = 0;
		char c = (char)s.get8();
		while (!s.at_eof() && c != '\n') {
			buffer[len++] = c;
			if (len == 1023) {
				while (!s.at_eof() && s.get8() != '\n') {}
				break;
			}
			c = (char)s.get8();
		}
		buffer[len] = 0;
		return buffer;
	};
	const char *head = token();
	if (strcmp(head, "#?RADIANCE") != 0 && strcmp(head, "#?RGBE") != 0) return false;
	bool valid = false;
	for (;;) {
		const char *t = token();
		if (t[0] == 0) break;
		if (strcmp(t, "FORMAT=32-bit_rle_rgbe") == 0) valid = true;
	}
	if (!valid) return false;
	char *t = token();
	if (strncmp(t, "-Y ", 3)) return false;
	t += 3;
	const long lh = strtol(t, &t, 10);
	while (*t == ' ') ++t;
	if (strncmp(t, "+X ", 3)) return false;
	t += 3;
	const long lw = strtol(t, nullptr, 10);
	const int height = (int)lh, width = (int)lw;
	if (!sizes_ok(width, height, 3, 4) || width == 0 || height == 0) return false;
	std::vector<float> hdr((size_t)width * height * 3, 0.0f);
	auto convert = [](float *o, const uint8_t *in) { // stbi__hdr_convert, three components
		if (in[3] != 0) {
			const float f1 = (float)ldexp(1.0f, in[3] - (int)(128 + 8));
			o[0] = in[0] * f1;
			o[1] = in[1] * f1;
			o[2] = in[2] * f1;
		} else
			o[0] = o[1] = o[2] = 0.0f;
	};
	uint8_t rgbe[4] = {0, 0, 0, 0}; // one buffer for the whole loop: after the end of a short file the last bytes repeat
	auto flat_from = [&](int j0, int i0) { // the flat loop, entered at (row j0, pixel i0)
		for (int j = j0; j < height; ++j)
			for (int i = (j == j0 ? i0 : 0); i < width; ++i) {
				s.getn(rgbe, 4);
				convert(&hdr[((size_t)j * width + i) * 3], rgbe);
			}
	};
	if (width < 8 || width >= 32768)
		flat_from(0, 0);
	else {
		std::vector<uint8_t> scan((size_t)width * 4, 0);
		for (int j = 0; j < height; ++j) {
			const int c1 = s.get8(), c2 = s.get8();
			int len = s.get8();
			if (c1 != 2 || c2 != 2 || (len & 0x80)) {
				// not run-length encoded: these four bytes are pixel 0 of ROW 0, and everything after them is flat data
				const uint8_t first_px[4] = {(uint8_t)c1, (uint8_t)c2, (uint8_t)len, (uint8_t)s.get8()};
				convert(&hdr[0], first_px);
				flat_from(0, 1);
				break;
			}
			len = (len << 8) | s.get8();
			if (len != width) return false;
			for (int k = 0; k < 4; ++k) {
				int i = 0, nleft;
				while ((nleft = width - i) > 0) {
					uint8_t count = (uint8_t)s.get8();
					if (count == 0 && s.at_eof()) return false; // past the end every count is 0: stb_image spins here for ever
					if (count > 128) {
						const uint8_t value = (uint8_t)s.get8();
						count -= 128;
						if (count > nleft) return false;
						for (int z = 0; z < count; ++z) scan[(size_t)(i++) * 4 + k] = value;
					} else {
						if (count > nleft) return false;
						for (int z = 0; z < count; ++z) scan[(size_t)(i++) * 4 + k] = (uint8_t)s.get8();
					}
				}
			}
			for (int i = 0; i < width; ++i) convert(&hdr[((size_t)j * width + i) * 3], &scan[(size_t)i * 4]);
		}
	}
	// stbi__hdr_to_ldr with the default gamma 2.2 and scale 1
	static const float gamma_i = 1.0f / 2.2f, scale_i = 1.0f;
	img->width = width;
	img->height = height;
	img->rgb.resize(hdr.size());
	for (size_t i = 0; i < hdr.size(); ++i) {
		float z = (float)pow(hdr[i] * scale_i, gamma_i) * 255 + 0.5f;
		if (z < 0) z = 0;
		if (z > 255) z = 255;
		img->rgb[i] = (uint8_t)(int)z;
	}
	return true;
}

} // namespace

// stbi__load_main's order for these formats (stb_image.h:973-989): GIF, PSD, PIC, PNM, HDR; TGA comes after them
bool decode_rare_formats(const std::vector<uint8_t> &f, DecodedImage *img, bool *recognised)
{
	*recognised = true;
	if (f.size() >= 6 && !memcmp(f.data(), "GIF8", 4) && (f[4] == '7' || f[4] == '9') && f[5] == 'a') return decode_gif(f, img);
	if (f.size() >= 4 && !memcmp(f.data(), "8BPS", 4)) return decode_psd(f, img);
	if (f.size() >= 92 && f[0] == 0x53 && f[1] == 0x80 && f[2] == 0xF6 && f[3] == 0x34 && !memcmp(f.data() + 88, "PICT", 4)) return decode_pic(f, img);
	if (f.size() >= 2 && f[0] == 'P' && (f[1] == '5' || f[1] == '6')) return decode_pnm(f, img);
	if ((f.size() >= 11 && !memcmp(f.data(), "#?RADIANCE\n", 11)) || (f.size() >= 7 && !memcmp(f.data(), "#?RGBE\n", 7))) return decode_hdr(f, img);
	*recognised = false;
	return false;
}

} // namespace host
} // namespace adypt
