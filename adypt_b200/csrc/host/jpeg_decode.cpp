// JPEG decoding for diffuse maps (map_Kd): the stand-in for stbi_load(filename, &w, &h, &n, 3) on a JPEG file, as
// OglScene::load_texture reaches it (src/Tracer/OglScene.cpp:12-43 -> dep/stb_image.h:3503-3674). Output: RGB8.
//
// A texel that differs by one count changes rendered pixels, so this decoder is held to stb_image's RESULT byte
// for byte (tests/test_textures.py compares with the reference's own stb_image compiled in oracle/_ref). That fixes
// the arithmetic, which is restated here from the format and from stb_image's documented choices:
//   * baseline and progressive Huffman JPEG, 8-bit, 1 / 3 / 4 components, sampling factors 1..4, restart intervals,
//     8- and 16-bit quantisation tables (ITU T.81); arithmetic coding and lossless modes are rejected, as there;
//   * coefficients are kept in 16-bit integers (products wrap like a C `short`);
//   * inverse DCT: the integer "islow" algorithm of the IJG library in the variant stb_image uses -- constants
//     scaled by 2^12, the column pass keeps 2 extra bits, the row pass rounds, adds 128 and clamps
//     (stb_image.h:2115-2205);
//   * chroma upsampling: triangle filters for 2x horizontally / vertically / both, pixel replication for any other
//     ratio (stb_image.h:3124-3316);
//   * YCbCr -> RGB in 20-bit fixed point with the green/Cb product truncated to its high 16 bits
//     (stb_image.h:3319-3345); Adobe CMYK / YCCK through the rounded 8x8 multiply; RGB-tagged files are copied.
#include <cstdint>
#include <cstring>
#include <vector>
#include "host_scene.h"

namespace adypt {
namespace host {

namespace {

// zigzag position -> row-major position; 15 trailing entries catch runs that overshoot in corrupt files
const uint8_t kNatural[64 + 15] = {
	0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
	35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
	63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct Huffman {
	// canonical code: for each length L, codes [first[L], first[L] + count[L]) map to symbols[offset[L] ...]
	uint8_t symbols[256];
	int32_t maxcode[18]; // (largest code of length L) + 1, left-aligned to 16 bits; 0 for unused lengths
	int32_t delta[17];   // index of a length-L code c in symbols = c + delta[L]
	bool valid = false;

	Huffman()
	{
		memset(symbols, 0, sizeof(symbols));
		memset(maxcode, 0, sizeof(maxcode));
		memset(delta, 0, sizeof(delta));
		maxcode[17] = 0x7fffffff; // an undefined table decodes nothing
	}

	bool build(const int counts[16])
	{
		int code = 0, k = 0;
		for (int len = 1; len <= 16; ++len) {
			delta[len] = k - code;
			code += counts[len - 1];
			k += counts[len - 1];
			if (counts[len - 1] && code - 1 >= (1 << len)) return false;
			maxcode[len] = code << (16 - len);
			code <<= 1;
		}
		maxcode[17] = 0x7fffffff;
		valid = k <= 256;
		return valid;
	}
};

struct Component {
	int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0;
	int dc_pred = 0;
	int x = 0, y = 0;   // samples that carry image data
	int w2 = 0, h2 = 0; // allocated plane (whole MCUs)
	std::vector<uint8_t> plane;
	std::vector<int16_t> coeff; // progressive: all blocks, 64 coefficients each, row-major inside a block
	int blocks_w = 0;
};

class Decoder {
public:
	Decoder(const uint8_t *data, size_t size) : p_(data), end_(data + size) {}

	bool run(DecodedImage *img)
	{
		if (!decode_scans()) return false;
		return convert(img);
	}

private:
	const uint8_t *p_, *end_;
	// entropy-coded segment reader
	uint32_t buf_ = 0;
	int bits_ = 0;
	bool nomore_ = false;
	int marker_ = 0xff; // 0xff = none pending
	// tables and frame
	uint16_t quant_[4][64];
	Huffman dc_[4], ac_[4];
	Component comp_[4];
	int ncomp_ = 0, width_ = 0, height_ = 0;
	int hmax_ = 1, vmax_ = 1, mcus_x_ = 0, mcus_y_ = 0;
	bool progressive_ = false, jfif_ = false;
	int adobe_transform_ = -1, rgb_ids_ = 0;
	int restart_interval_ = 0, todo_ = 0;
	// scan
	int scan_n_ = 0, order_[4] = {0, 0, 0, 0};
	int ss_ = 0, se_ = 63, ah_ = 0, al_ = 0, eob_run_ = 0;

	// ---- byte level
	int get8() { return p_ < end_ ? *p_++ : 0; }
	int get16() { const int a = get8(); return (a << 8) | get8(); }
	bool at_eof() const { return p_ >= end_; }
	void skip(int n) { p_ = (n < 0 || n > end_ - p_) ? end_ : p_ + n; }

	int next_marker()
	{
		if (marker_ != 0xff) { const int m = marker_; marker_ = 0xff; return m; }
		int x = get8();
		if (x != 0xff) return 0xff;
		while (x == 0xff) x = get8();
		return x;
	}

	// ---- bit level (T.81 F.2.2.5: 0xFF00 is a stuffed 0xFF; a marker ends the segment and zero bits follow)
	void fill()
	{
		do {
			const uint32_t b = nomore_ ? 0u : (uint32_t)get8();
			if (b == 0xff) {
				int c = get8();
				while (c == 0xff) c = get8();
				if (c != 0) { marker_ = c; nomore_ = true; return; }
			}
			buf_ |= b << (24 - bits_);
			bits_ += 8;
		} while (bits_ <= 24);
	}
	int get_bits(int n)
	{
		if (n == 0) return 0;
		if (bits_ < n) fill();
		const uint32_t v = buf_ >> (32 - n);
		buf_ <<= n;
		bits_ -= n;
		return (int)v;
	}
	int get_bit() { return get_bits(1); }
	// n magnitude bits -> signed value (T.81 F.2.2.1 EXTEND)
	int receive_extend(int n)
	{
		if (n == 0) return 0;
		const int v = get_bits(n);
		return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
	}
	int decode_symbol(const Huffman &h)
	{
		if (bits_ < 16) fill();
		const int32_t top = (int32_t)(buf_ >> 16);
		int len = 1;
		while (top >= h.maxcode[len]) ++len;
		if (len > 16 || len > bits_) return -1;
		const int idx = (int)((buf_ >> (32 - len)) & ((1u << len) - 1u)) + h.delta[len];
		if (idx < 0 || idx > 255) return -1;
		buf_ <<= len;
		bits_ -= len;
		return h.symbols[idx];
	}

	void reset_entropy()
	{
		bits_ = 0;
		buf_ = 0;
		nomore_ = false;
		for (Component &c : comp_) c.dc_pred = 0;
		marker_ = 0xff;
		todo_ = restart_interval_ ? restart_interval_ : 0x7fffffff;
		eob_run_ = 0;
	}
	// after each MCU: true = keep going, false = the segment ended without a restart marker (stop this scan quietly)
	bool mcu_done()
	{
		if (--todo_ > 0) return true;
		if (bits_ < 24) fill();
		if (marker_ < 0xd0 || marker_ > 0xd7) return false;
		reset_entropy();
		return true;
	}

	// ---- segments
	bool read_tables(int m)
	{
		if (m == 0xff) return false; // expected a marker
		if (m == 0xdd) { // DRI
			if (get16() != 4) return false;
			restart_interval_ = get16();
			return true;
		}
		if (m == 0xdb) { // DQT
			int len = get16() - 2;
			while (len > 0) {
				const int q = get8(), wide = q >> 4, t = q & 15;
				if (wide > 1 || t > 3) return false;
				for (int i = 0; i < 64; ++i) quant_[t][kNatural[i]] = (uint16_t)(wide ? get16() : get8());
				len -= wide ? 129 : 65;
			}
			return len == 0;
		}
		if (m == 0xc4) { // DHT
			int len = get16() - 2;
			while (len > 0) {
				const int q = get8(), cls = q >> 4, t = q & 15;
				if (cls > 1 || t > 3) return false;
				int counts[16], n = 0;
				for (int i = 0; i < 16; ++i) n += counts[i] = get8();
				if (n > 256) return false;
				Huffman &h = cls ? ac_[t] : dc_[t];
				if (!h.build(counts)) return false;
				for (int i = 0; i < n; ++i) h.symbols[i] = (uint8_t)get8();
				len -= 17 + n;
			}
			return len == 0;
		}
		if ((m >= 0xe0 && m <= 0xef) || m == 0xfe) { // APPn / COM
			int len = get16();
			if (len < 2) return false;
			len -= 2;
			if (m == 0xe0 && len >= 5) {
				static const char tag[5] = {'J', 'F', 'I', 'F', 0};
				bool ok = true;
				for (int i = 0; i < 5; ++i) ok &= get8() == (uint8_t)tag[i];
				len -= 5;
				if (ok) jfif_ = true;
			} else if (m == 0xee && len >= 12) {
				static const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
				bool ok = true;
				for (int i = 0; i < 6; ++i) ok &= get8() == (uint8_t)tag[i];
				len -= 6;
				if (ok) {
					get8(); get16(); get16();
					adobe_transform_ = get8();
					len -= 6;
				}
			}
			skip(len);
			return true;
		}
		return false; // SOF3.., DAC, ...: not supported (nor by the reference's decoder)
	}

	bool read_frame()
	{
		const int len = get16();
		if (len < 11 || get8() != 8) return false;
		height_ = get16();
		width_ = get16();
		ncomp_ = get8();
		if (!height_ || !width_ || (ncomp_ != 1 && ncomp_ != 3 && ncomp_ != 4) || len != 8 + 3 * ncomp_) return false;
		if ((uint64_t)width_ * (uint64_t)height_ * 3u > 0x7fffffffu) return false;
		for (int i = 0; i < ncomp_; ++i) {
			Component &c = comp_[i];
			c.id = get8();
			if (ncomp_ == 3 && c.id == "RGB"[i]) ++rgb_ids_;
			const int q = get8();
			c.h = q >> 4;
			c.v = q & 15;
			c.tq = get8();
			if (!c.h || c.h > 4 || !c.v || c.v > 4 || c.tq > 3) return false;
			if (c.h > hmax_) hmax_ = c.h;
			if (c.v > vmax_) vmax_ = c.v;
		}
		mcus_x_ = (width_ + hmax_ * 8 - 1) / (hmax_ * 8);
		mcus_y_ = (height_ + vmax_ * 8 - 1) / (vmax_ * 8);
		for (int i = 0; i < ncomp_; ++i) {
			Component &c = comp_[i];
			c.x = (width_ * c.h + hmax_ - 1) / hmax_;
			c.y = (height_ * c.v + vmax_ - 1) / vmax_;
			c.w2 = mcus_x_ * c.h * 8;
			c.h2 = mcus_y_ * c.v * 8;
			c.plane.assign((size_t)c.w2 * c.h2, 0);
			if (progressive_) {
				c.blocks_w = c.w2 / 8;
				c.coeff.assign((size_t)c.w2 * c.h2, 0);
			}
		}
		return true;
	}

	bool read_scan_header()
	{
		const int len = get16();
		scan_n_ = get8();
		if (scan_n_ < 1 || scan_n_ > 4 || scan_n_ > ncomp_ || len != 6 + 2 * scan_n_) return false;
		for (int i = 0; i < scan_n_; ++i) {
			const int id = get8(), q = get8();
			int which = 0;
			while (which < ncomp_ && comp_[which].id != id) ++which;
			if (which == ncomp_) return false;
			comp_[which].hd = q >> 4;
			comp_[which].ha = q & 15;
			if (comp_[which].hd > 3 || comp_[which].ha > 3) return false;
			order_[i] = which;
		}
		ss_ = get8();
		se_ = get8();
		const int a = get8();
		ah_ = a >> 4;
		al_ = a & 15;
		if (progressive_) {
			if (ss_ > 63 || se_ > 63 || ss_ > se_ || ah_ > 13 || al_ > 13) return false;
		} else {
			if (ss_ != 0 || ah_ != 0 || al_ != 0) return false;
			se_ = 63;
		}
		return true;
	}

	// ---- blocks
	// sequential mode: one dequantised block (T.81 F.2.2)
	bool block_sequential(int16_t *blk, Component &c)
	{
		const Huffman &hd = dc_[c.hd], &ha = ac_[c.ha];
		const uint16_t *q = quant_[c.tq];
		const int t = decode_symbol(hd);
		if (t < 0 || t > 16) return false;
		memset(blk, 0, 64 * sizeof(int16_t));
		c.dc_pred += receive_extend(t);
		blk[0] = (int16_t)(c.dc_pred * q[0]);
		for (int k = 1; k < 64;) {
			const int rs = decode_symbol(ha);
			if (rs < 0) return false;
			const int s = rs & 15, r = rs >> 4;
			if (s == 0) {
				if (rs != 0xf0) break;
				k += 16;
			} else {
				k += r;
				const int z = kNatural[k++];
				blk[z] = (int16_t)(receive_extend(s) * q[z]);
			}
		}
		return true;
	}
	// progressive mode, DC scans (T.81 G.1.2.1)
	bool block_prog_dc(int16_t *blk, Component &c)
	{
		if (se_ != 0) return false;
		if (ah_ == 0) {
			memset(blk, 0, 64 * sizeof(int16_t));
			const int t = decode_symbol(dc_[c.hd]);
			if (t < 0 || t > 16) return false;
			c.dc_pred += receive_extend(t);
			blk[0] = (int16_t)(c.dc_pred * (1 << al_));
		} else if (get_bit())
			blk[0] = (int16_t)(blk[0] + (1 << al_));
		return true;
	}
	void refine(int16_t *coef, int16_t bit) // correction bit for an already non-zero coefficient (G.1.2.3)
	{
		if (get_bit() && (*coef & bit) == 0) *coef = (int16_t)(*coef > 0 ? *coef + bit : *coef - bit);
	}
	// progressive mode, AC scans (T.81 G.1.2.2 / G.1.2.3)
	bool block_prog_ac(int16_t *blk, Component &c)
	{
		if (ss_ == 0) return false;
		const Huffman &ha = ac_[c.ha];
		if (ah_ == 0) {
			if (eob_run_) { --eob_run_; return true; }
			int k = ss_;
			do {
				const int rs = decode_symbol(ha);
				if (rs < 0) return false;
				const int s = rs & 15, r = rs >> 4;
				if (s == 0) {
					if (r < 15) {
						eob_run_ = (1 << r) + (r ? get_bits(r) : 0) - 1;
						break;
					}
					k += 16;
				} else {
					k += r;
					blk[kNatural[k++]] = (int16_t)(receive_extend(s) * (1 << al_));
				}
			} while (k <= se_);
			return true;
		}
		const int16_t bit = (int16_t)(1 << al_);
		if (eob_run_) {
			--eob_run_;
			for (int k = ss_; k <= se_; ++k) {
				int16_t *coef = &blk[kNatural[k]];
				if (*coef != 0) refine(coef, bit);
			}
			return true;
		}
		int k = ss_;
		do {
			const int rs = decode_symbol(ha);
			if (rs < 0) return false;
			int s = rs & 15, r = rs >> 4;
			if (s == 0) {
				if (r < 15) {
					eob_run_ = (1 << r) - 1 + (r ? get_bits(r) : 0);
					r = 64; // run to the end of the band
				}
			} else {
				if (s != 1) return false;
				s = get_bit() ? bit : -bit;
			}
			while (k <= se_) {
				int16_t *coef = &blk[kNatural[k++]];
				if (*coef != 0)
					refine(coef, bit);
				else {
					if (r == 0) { *coef = (int16_t)s; break; }
					--r;
				}
			}
		} while (k <= se_);
		return true;
	}

	bool scan()
	{
		reset_entropy();
		int16_t tmp[64];
		if (scan_n_ == 1) { // non-interleaved: the component's own blocks in raster order
			Component &c = comp_[order_[0]];
			const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
			for (int by = 0; by < bh; ++by)
				for (int bx = 0; bx < bw; ++bx) {
					if (!progressive_) {
						if (!block_sequential(tmp, c)) return false;
						idct(tmp, &c.plane[(size_t)c.w2 * by * 8 + (size_t)bx * 8], c.w2);
					} else {
						int16_t *blk = &c.coeff[64 * ((size_t)bx + (size_t)by * c.blocks_w)];
						if (!(ss_ == 0 ? block_prog_dc(blk, c) : block_prog_ac(blk, c))) return false;
					}
					if (!mcu_done()) return true;
				}
			return true;
		}
		for (int my = 0; my < mcus_y_; ++my)
			for (int mx = 0; mx < mcus_x_; ++mx) {
				for (int k = 0; k < scan_n_; ++k) {
					Component &c = comp_[order_[k]];
					for (int y = 0; y < c.v; ++y)
						for (int x = 0; x < c.h; ++x) {
							const int bx = mx * c.h + x, by = my * c.v + y;
							if (!progressive_) {
								if (!block_sequential(tmp, c)) return false;
								idct(tmp, &c.plane[(size_t)c.w2 * by * 8 + (size_t)bx * 8], c.w2);
							} else if (!block_prog_dc(&c.coeff[64 * ((size_t)bx + (size_t)by * c.blocks_w)], c))
								return false; // interleaved progressive scans carry DC only
						}
				}
				if (!mcu_done()) return true;
			}
		return true;
	}

	bool decode_scans()
	{
		memset(quant_, 0, sizeof(quant_));
		if (next_marker() != 0xd8) return false; // SOI
		int m = next_marker();
		while (m != 0xc0 && m != 0xc1 && m != 0xc2) {
			if (!read_tables(m)) return false;
			m = next_marker();
			while (m == 0xff) { // padding between segments
				if (at_eof()) return false;
				m = next_marker();
			}
		}
		progressive_ = m == 0xc2;
		if (!read_frame()) return false;
		m = next_marker();
		while (m != 0xd9) { // EOI
			if (m == 0xda) {
				if (!read_scan_header() || !scan()) return false;
				if (marker_ == 0xff) { // zero padding after the entropy-coded data: look for the next marker
					while (!at_eof())
						if (get8() == 0xff) { marker_ = get8(); break; }
				}
			} else if (m == 0xdc) { // DNL
				get16(); get16();
			} else if (!read_tables(m))
				return false;
			m = next_marker();
		}
		if (progressive_)
			for (int i = 0; i < ncomp_; ++i) {
				Component &c = comp_[i];
				const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
				for (int by = 0; by < bh; ++by)
					for (int bx = 0; bx < bw; ++bx) {
						int16_t *blk = &c.coeff[64 * ((size_t)bx + (size_t)by * c.blocks_w)];
						for (int k = 0; k < 64; ++k) blk[k] = (int16_t)(blk[k] * quant_[c.tq][k]);
						idct(blk, &c.plane[(size_t)c.w2 * by * 8 + (size_t)bx * 8], c.w2);
					}
			}
		return true;
	}

	// ---- inverse DCT
	static int fix(double x) { return (int)(x * 4096 + 0.5); }
	// one 8-point pass of the IJG "islow" IDCT on s[0..7] (stride st): even part in e[0..3], odd part in o[0..3];
	// output k = e[k] + o[3-k], output 7-k = e[k] - o[3-k]
	static void idct8(const int *s, int st, int e[4], int o[4])
	{
		static const int c0541 = fix(0.5411961f), c1847 = fix(-1.847759065f), c0765 = fix(0.765366865f), c1175 = fix(1.175875602f),
		                 c0298 = fix(0.298631336f), c2053 = fix(2.053119869f), c3072 = fix(3.072711026f), c1501 = fix(1.501321110f),
		                 c0899 = fix(-0.899976223f), c2562 = fix(-2.562915447f), c1961 = fix(-1.961570560f), c0390 = fix(-0.390180644f);
		const int s0 = s[0], s1 = s[st], s2 = s[2 * st], s3 = s[3 * st], s4 = s[4 * st], s5 = s[5 * st], s6 = s[6 * st], s7 = s[7 * st];
		const int z = (s2 + s6) * c0541;
		const int a2 = z + s6 * c1847, a3 = z + s2 * c0765;
		const int a0 = (s0 + s4) * 4096, a1 = (s0 - s4) * 4096;
		e[0] = a0 + a3; e[3] = a0 - a3; e[1] = a1 + a2; e[2] = a1 - a2;
		const int p3 = s7 + s3, p4 = s5 + s1, p1 = s7 + s1, p2 = s5 + s3;
		const int p5 = (p3 + p4) * c1175;
		const int q1 = p5 + p1 * c0899, q2 = p5 + p2 * c2562, q3 = p3 * c1961, q4 = p4 * c0390;
		o[3] = s1 * c1501 + q1 + q4;
		o[2] = s3 * c3072 + q2 + q3;
		o[1] = s5 * c2053 + q2 + q4;
		o[0] = s7 * c0298 + q1 + q3;
	}
	static uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
	static void idct(const int16_t *blk, uint8_t *out, int stride)
	{
		int in[64], mid[64], e[4], o[4];
		for (int i = 0; i < 64; ++i) in[i] = blk[i];
		for (int col = 0; col < 8; ++col) { // columns, keeping 2 fractional bits
			const int *s = in + col;
			if (!(s[8] | s[16] | s[24] | s[32] | s[40] | s[48] | s[56])) {
				const int dc = s[0] * 4;
				for (int r = 0; r < 8; ++r) mid[r * 8 + col] = dc;
				continue;
			}
			idct8(s, 8, e, o);
			for (int k = 0; k < 4; ++k) {
				const int ek = e[k] + 512;
				mid[k * 8 + col] = (ek + o[3 - k]) >> 10;
				mid[(7 - k) * 8 + col] = (ek - o[3 - k]) >> 10;
			}
		}
		for (int row = 0; row < 8; ++row, out += stride) { // rows: remove 2^17, round, level-shift by 128
			idct8(mid + row * 8, 1, e, o);
			for (int k = 0; k < 4; ++k) {
				const int ek = e[k] + 65536 + (128 << 17);
				out[k] = clamp8((ek + o[3 - k]) >> 17);
				out[7 - k] = clamp8((ek - o[3 - k]) >> 17);
			}
		}
	}

	// ---- upsampling of one output row (JFIF-centred triangle filters)
	static const uint8_t *upsample(uint8_t *out, const uint8_t *near, const uint8_t *far, int w, int hs, int vs)
	{
		if (hs == 1 && vs == 1) return near;
		if (hs == 1 && vs == 2) {
			for (int i = 0; i < w; ++i) out[i] = (uint8_t)((3 * near[i] + far[i] + 2) >> 2);
			return out;
		}
		if (hs == 2 && vs == 1) {
			if (w == 1) { out[0] = out[1] = near[0]; return out; }
			out[0] = near[0];
			out[1] = (uint8_t)((near[0] * 3 + near[1] + 2) >> 2);
			int i = 1;
			for (; i < w - 1; ++i) {
				const int n = 3 * near[i] + 2;
				out[i * 2] = (uint8_t)((n + near[i - 1]) >> 2);
				out[i * 2 + 1] = (uint8_t)((n + near[i + 1]) >> 2);
			}
			out[i * 2] = (uint8_t)((near[w - 2] * 3 + near[w - 1] + 2) >> 2);
			out[i * 2 + 1] = near[w - 1];
			return out;
		}
		if (hs == 2 && vs == 2) {
			if (w == 1) { out[0] = out[1] = (uint8_t)((3 * near[0] + far[0] + 2) >> 2); return out; }
			int t1 = 3 * near[0] + far[0];
			out[0] = (uint8_t)((t1 + 2) >> 2);
			for (int i = 1; i < w; ++i) {
				const int t0 = t1;
				t1 = 3 * near[i] + far[i];
				out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
				out[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
			}
			out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
			return out;
		}
		for (int i = 0; i < w; ++i) // any other ratio: nearest neighbour horizontally, rows repeat vertically
			for (int j = 0; j < hs; ++j) out[i * hs + j] = near[i];
		return out;
	}

	static uint8_t mul8(int x, int y) // round(x*y/255)
	{
		const unsigned t = (unsigned)(x * y + 128);
		return (uint8_t)((t + (t >> 8)) >> 8);
	}
	static int fixed20(float x) { return ((int)(x * 4096.0f + 0.5f)) << 8; }
	static void ycc_to_rgb(uint8_t *out, const uint8_t *y, const uint8_t *cb, const uint8_t *cr, int n)
	{
		static const int k_r = fixed20(1.40200f), k_gr = fixed20(0.71414f), k_gb = fixed20(0.34414f), k_b = fixed20(1.77200f);
		for (int i = 0; i < n; ++i, out += 3) {
			const int yf = (y[i] << 20) + (1 << 19), r_ = cr[i] - 128, b_ = cb[i] - 128;
			const int r = (yf + r_ * k_r) >> 20;
			const int g = (yf + r_ * -k_gr + (int)((unsigned)(b_ * -k_gb) & 0xffff0000u)) >> 20;
			const int b = (yf + b_ * k_b) >> 20;
			out[0] = clamp8(r); out[1] = clamp8(g); out[2] = clamp8(b);
		}
	}

	bool convert(DecodedImage *img)
	{
		img->width = width_;
		img->height = height_;
		img->rgb.assign((size_t)width_ * height_ * 3, 0);
		const bool is_rgb = ncomp_ == 3 && (rgb_ids_ == 3 || (adobe_transform_ == 0 && !jfif_));
		struct Row { int hs, vs, ystep, ypos, w_lores; const uint8_t *line0, *line1; std::vector<uint8_t> buf; } rows[4];
		for (int k = 0; k < ncomp_; ++k) {
			Row &r = rows[k];
			r.hs = hmax_ / comp_[k].h;
			r.vs = vmax_ / comp_[k].v;
			r.ystep = r.vs >> 1;
			r.ypos = 0;
			r.w_lores = (width_ + r.hs - 1) / r.hs;
			r.line0 = r.line1 = comp_[k].plane.data();
			r.buf.assign((size_t)width_ + 8, 0);
		}
		const uint8_t *c[4] = {nullptr, nullptr, nullptr, nullptr};
		for (int j = 0; j < height_; ++j) {
			uint8_t *out = &img->rgb[(size_t)j * width_ * 3];
			for (int k = 0; k < ncomp_; ++k) {
				Row &r = rows[k];
				const bool bottom = r.ystep >= (r.vs >> 1);
				c[k] = upsample(r.buf.data(), bottom ? r.line1 : r.line0, bottom ? r.line0 : r.line1, r.w_lores, r.hs, r.vs);
				if (++r.ystep >= r.vs) {
					r.ystep = 0;
					r.line0 = r.line1;
					if (++r.ypos < comp_[k].y) r.line1 += comp_[k].w2;
				}
			}
			if (ncomp_ == 1) {
				for (int i = 0; i < width_; ++i) out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = c[0][i];
			} else if (ncomp_ == 3 && is_rgb) {
				for (int i = 0; i < width_; ++i) { out[3 * i] = c[0][i]; out[3 * i + 1] = c[1][i]; out[3 * i + 2] = c[2][i]; }
			} else if (ncomp_ == 4 && adobe_transform_ == 0) { // CMYK (stored inverted)
				for (int i = 0; i < width_; ++i)
					for (int ch = 0; ch < 3; ++ch) out[3 * i + ch] = mul8(c[ch][i], c[3][i]);
			} else {
				ycc_to_rgb(out, c[0], c[1], c[2], width_);
				if (ncomp_ == 4 && adobe_transform_ == 2) // YCCK
					for (int i = 0; i < width_; ++i)
						for (int ch = 0; ch < 3; ++ch) out[3 * i + ch] = mul8(255 - out[3 * i + ch], c[3][i]);
			}
		}
		return true;
	}
};

} // namespace

bool decode_jpeg(const std::vector<uint8_t> &file, DecodedImage *img)
{
	if (file.size() < 4 || file[0] != 0xff || file[1] != 0xd8) return false;
	Decoder d(file.data(), file.size());
	return d.run(img);
}

} // namespace host
} // namespace adypt
