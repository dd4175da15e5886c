// OBJ/MTL ingest: the stand-in for Scene::LoadFromFile (src/Util/Scene.cpp:9-136) and
// OglScene::init_materials (src/Tracer/OglScene.cpp:51-91, minus the GL texture upload).
//
// The reference delegates parsing to its vendored tinyobjloader 1.2.0 (dep/tiny_obj_loader.h). To get the SAME
// Triangle bytes the parts of that library that decide a byte are restated here, from its documented
// behaviour and arithmetic:
//   * decimal -> double conversion (tiny_obj_loader.h:525-636): digit accumulation with a 1e-k table and
//     ldexp(m * 5^e, e) for exponents -- not correctly rounded, so strtod would differ in rare last-bit cases;
//   * index fix-up, 1-based / negative-relative (tiny_obj_loader.h:742-773);
//   * polygon triangulation by ear clipping on the dominant plane (tiny_obj_loader.h:1043-1236);
//   * material defaults (tiny_obj_loader.h:964-1000) and the keys Adypt reads: Kd Ke Ks Ns Ni d Tr illum map_Kd.
// Faces come out in file order, exactly one Triangle per emitted triangle.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include "host_scene.h"

namespace adypt {
namespace host {

namespace {

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline bool is_space(char c) { return c == ' ' || c == '\t'; }

// tinyobj's tryParseDouble: [sign] digits [. digits] [e|E [sign] digits]
bool parse_double(const char *s, const char *end, double *out)
{
	if (s >= end) return false;
	double mantissa = 0.0;
	int exponent = 0;
	char sign = '+', exp_sign = '+';
	const char *c = s;
	if (*c == '+' || *c == '-') sign = *c++;
	else if (!is_digit(*c)) return false;
	int read = 0;
	while (c != end && is_digit(*c)) {
		mantissa *= 10;
		mantissa += (int)(*c - '0');
		++c;
		++read;
	}
	if (read == 0) return false;
	bool has_exp = false;
	if (c != end) {
		if (*c == '.') {
			++c;
			read = 1;
			static const double lut[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
			while (c != end && is_digit(*c)) {
				mantissa += (int)(*c - '0') * (read < 8 ? lut[read] : std::pow(10.0, -read));
				++read;
				++c;
			}
			has_exp = c != end && (*c == 'e' || *c == 'E');
		} else if (*c == 'e' || *c == 'E')
			has_exp = true;
	}
	if (has_exp) {
		++c;
		if (c != end && (*c == '+' || *c == '-')) exp_sign = *c++;
		else if (c == end || !is_digit(*c)) return false;
		read = 0;
		while (c != end && is_digit(*c)) {
			exponent = exponent * 10 + (int)(*c - '0');
			++c;
			++read;
		}
		exponent *= (exp_sign == '+' ? 1 : -1);
		if (read == 0) return false;
	}
	*out = (sign == '+' ? 1 : -1) * (exponent ? std::ldexp(mantissa * std::pow(5.0, exponent), exponent) : mantissa);
	return true;
}

float next_real(const char **tok, double dflt = 0.0)
{
	*tok += strspn(*tok, " \t");
	const char *end = *tok + strcspn(*tok, " \t\r");
	double v = dflt;
	parse_double(*tok, end, &v);
	*tok = end;
	return (float)v;
}

struct Corner {
	int v = -1, vt = -1, vn = -1;
};

// 1-based -> 0-based, negative = relative to the elements read so far, 0 = invalid
bool fix_index(int idx, int n, int *out)
{
	if (idx > 0) { *out = idx - 1; return true; }
	if (idx == 0) return false;
	*out = n + idx;
	return true;
}

bool parse_corner(const char **tok, int nv, int nvn, int nvt, Corner *c)
{
	if (!fix_index(atoi(*tok), nv, &c->v)) return false;
	*tok += strcspn(*tok, "/ \t\r");
	if ((*tok)[0] != '/') return true;
	++*tok;
	if ((*tok)[0] == '/') { // v//vn
		++*tok;
		if (!fix_index(atoi(*tok), nvn, &c->vn)) return false;
		*tok += strcspn(*tok, "/ \t\r");
		return true;
	}
	if (!fix_index(atoi(*tok), nvt, &c->vt)) return false; // v/vt[/vn]
	*tok += strcspn(*tok, "/ \t\r");
	if ((*tok)[0] != '/') return true;
	++*tok;
	if (!fix_index(atoi(*tok), nvn, &c->vn)) return false;
	*tok += strcspn(*tok, "/ \t\r");
	return true;
}

int point_in_polygon(int n, const float *vx, const float *vy, float tx, float ty)
{
	int c = 0;
	for (int i = 0, j = n - 1; i < n; j = i++)
		if (((vy[i] > ty) != (vy[j] > ty)) && (tx < (vx[j] - vx[i]) * (ty - vy[i]) / (vy[j] - vy[i]) + vx[i])) c = !c;
	return c;
}

// ear clipping of one polygon, appends corner triples to `out`
void triangulate(const std::vector<Corner> &face, const std::vector<float> &v, std::vector<Corner> *out)
{
	size_t n = face.size();
	if (n < 3) return;
	size_t axes[2] = {1, 2};
	for (size_t k = 0; k < n; ++k) {
		const size_t a = (size_t)face[k % n].v, b = (size_t)face[(k + 1) % n].v, c = (size_t)face[(k + 2) % n].v;
		if (3 * a + 2 >= v.size() || 3 * b + 2 >= v.size() || 3 * c + 2 >= v.size()) continue;
		const float e0x = v[b * 3] - v[a * 3], e0y = v[b * 3 + 1] - v[a * 3 + 1], e0z = v[b * 3 + 2] - v[a * 3 + 2];
		const float e1x = v[c * 3] - v[b * 3], e1y = v[c * 3 + 1] - v[b * 3 + 1], e1z = v[c * 3 + 2] - v[b * 3 + 2];
		const float cx = std::fabs(e0y * e1z - e0z * e1y), cy = std::fabs(e0z * e1x - e0x * e1z), cz = std::fabs(e0x * e1y - e0y * e1x);
		const float eps = std::numeric_limits<float>::epsilon();
		if (cx > eps || cy > eps || cz > eps) {
			if (!(cx > cy && cx > cz)) {
				axes[0] = 0;
				if (cz > cx && cz > cy) axes[1] = 1;
			}
			break;
		}
	}
	float area = 0;
	for (size_t k = 0; k < n; ++k) {
		const size_t a = (size_t)face[k % n].v, b = (size_t)face[(k + 1) % n].v;
		if (a * 3 + axes[0] >= v.size() || a * 3 + axes[1] >= v.size() || b * 3 + axes[0] >= v.size() || b * 3 + axes[1] >= v.size()) continue;
		area += (v[a * 3 + axes[0]] * v[b * 3 + axes[1]] - v[a * 3 + axes[1]] * v[b * 3 + axes[0]]) * 0.5f;
	}
	std::vector<Corner> rest = face;
	int rounds = 10;
	size_t guess = 0;
	Corner ind[3];
	float vx[3], vy[3];
	while (rest.size() > 3 && rounds > 0) {
		n = rest.size();
		if (guess >= n) {
			rounds -= 1;
			guess -= n;
		}
		for (size_t k = 0; k < 3; ++k) {
			ind[k] = rest[(guess + k) % n];
			const size_t vi = (size_t)ind[k].v;
			if (vi * 3 + axes[0] >= v.size() || vi * 3 + axes[1] >= v.size()) vx[k] = vy[k] = 0.0f;
			else {
				vx[k] = v[vi * 3 + axes[0]];
				vy[k] = v[vi * 3 + axes[1]];
			}
		}
		const float e0x = vx[1] - vx[0], e0y = vy[1] - vy[0], e1x = vx[2] - vx[1], e1y = vy[2] - vy[1];
		const float cross = e0x * e1y - e0y * e1x;
		if (cross * area < 0.0f) { // reflex corner
			guess += 1;
			continue;
		}
		bool overlap = false;
		for (size_t other = 3; other < n; ++other) {
			const size_t idx = (guess + other) % n;
			if (idx >= rest.size()) continue;
			const size_t ovi = (size_t)rest[idx].v;
			if (ovi * 3 + axes[0] >= v.size() || ovi * 3 + axes[1] >= v.size()) continue;
			if (point_in_polygon(3, vx, vy, v[ovi * 3 + axes[0]], v[ovi * 3 + axes[1]])) {
				overlap = true;
				break;
			}
		}
		if (overlap) {
			guess += 1;
			continue;
		}
		out->push_back(ind[0]);
		out->push_back(ind[1]);
		out->push_back(ind[2]);
		rest.erase(rest.begin() + (long)((guess + 1) % n)); // drop the ear tip
	}
	if (rest.size() == 3) {
		out->push_back(rest[0]);
		out->push_back(rest[1]);
		out->push_back(rest[2]);
	}
}

struct MtlEntry {
	std::string name, diffuse_tex;
	float kd[3] = {0, 0, 0}, ke[3] = {0, 0, 0}, ks[3] = {0, 0, 0};
	int illum = 0;
	float dissolve = 1.0f, shininess = 1.0f, ior = 1.0f;
};

// Lines as tinyobj's safeGetline cuts them (tiny_obj_loader.h:419-451): '\n', "\r\n" and a lone '\r' all end a line
struct LineReader {
	std::string data;
	size_t pos = 0;
	bool open(const std::string &path)
	{
		std::ifstream in(path.c_str(), std::ios::binary);
		if (!in) return false;
		std::ostringstream ss;
		ss << in.rdbuf();
		data = ss.str();
		return true;
	}
	bool next(std::string *line)
	{
		if (pos >= data.size()) return false;
		size_t e = pos;
		while (e < data.size() && data[e] != '\n' && data[e] != '\r') ++e;
		line->assign(data, pos, e - pos);
		if (e < data.size()) e += (data[e] == '\r' && e + 1 < data.size() && data[e + 1] == '\n') ? 2 : 1;
		pos = e;
		return true;
	}
};

// what follows a texture keyword (tiny_obj_loader.h:861-962): options are consumed token by token -- a numeric option
// swallows as many tokens as it has parameters, whatever they are -- and the first thing that is not an option is
// the file name, taken to the end of the line
bool texture_name(const char *tok, std::string *name)
{
	auto word = [&](const char *w, size_t n) { return strncmp(tok, w, n) == 0 && is_space(tok[n]); };
	auto skip_token = [&]() { tok += strspn(tok, " \t"); tok += strcspn(tok, " \t\r"); };
	while (*tok != '\0' && *tok != '\r' && *tok != '\n') {
		tok += strspn(tok, " \t");
		if (word("-blendu", 7) || word("-blendv", 7)) { tok += 8; skip_token(); }
		else if (word("-clamp", 6) || word("-boost", 6)) { tok += 7; skip_token(); }
		else if (word("-bm", 3)) { tok += 4; skip_token(); }
		else if (word("-o", 2) || word("-s", 2) || word("-t", 2)) { tok += 3; skip_token(); skip_token(); skip_token(); }
		else if (word("-type", 5)) { tok += 5; skip_token(); }
		else if (word("-imfchan", 8)) { tok += 9; skip_token(); }
		else if (word("-mm", 3)) { tok += 4; skip_token(); skip_token(); }
		else {
			*name = tok;
			return true;
		}
	}
	return false;
}

// LoadMtl + MaterialFileReader (tiny_obj_loader.h:1273-1656, 1659-1692). false = the file cannot be opened.
bool load_mtl(const std::string &path, std::map<std::string, int> *index, std::vector<MtlEntry> *mats)
{
	LineReader in;
	if (!in.open(path)) return false; // tinyobj warns and tries the next name; with none left faces get material id -1
	MtlEntry cur; // "create a default material anyway": what precedes the first newmtl lands in a nameless one
	bool has_d = false;
	auto flush = [&]() {
		index->insert(std::make_pair(cur.name, (int)mats->size())); // the first definition of a name keeps it
		mats->push_back(cur);
	};
	std::string line;
	while (in.next(&line)) {
		const size_t e = line.find_last_not_of(" \t"); // the MTL reader trims trailing blanks (the OBJ reader does not)
		line = e == std::string::npos ? std::string() : line.substr(0, e + 1);
		const char *tok = line.c_str();
		tok += strspn(tok, " \t");
		if (*tok == '\0' || *tok == '#') continue;
		if (strncmp(tok, "newmtl", 6) == 0 && is_space(tok[6])) {
			if (!cur.name.empty()) flush();
			cur = MtlEntry();
			has_d = false;
			cur.name = tok + 7; // verbatim, inner blanks included
			continue;
		}
		if (tok[0] == 'K' && tok[1] == 'd' && is_space(tok[2])) { tok += 2; for (int i = 0; i < 3; ++i) cur.kd[i] = next_real(&tok); continue; }
		if (tok[0] == 'K' && tok[1] == 'e' && is_space(tok[2])) { tok += 2; for (int i = 0; i < 3; ++i) cur.ke[i] = next_real(&tok); continue; }
		if (tok[0] == 'K' && tok[1] == 's' && is_space(tok[2])) { tok += 2; for (int i = 0; i < 3; ++i) cur.ks[i] = next_real(&tok); continue; }
		if (tok[0] == 'N' && tok[1] == 'i' && is_space(tok[2])) { tok += 2; cur.ior = next_real(&tok); continue; }
		if (tok[0] == 'N' && tok[1] == 's' && is_space(tok[2])) { tok += 2; cur.shininess = next_real(&tok); continue; }
		if (strncmp(tok, "illum", 5) == 0 && is_space(tok[5])) { tok += 6; tok += strspn(tok, " \t"); cur.illum = atoi(tok); continue; }
		if (tok[0] == 'd' && is_space(tok[1])) { tok += 1; cur.dissolve = next_real(&tok); has_d = true; continue; }
		if (tok[0] == 'T' && tok[1] == 'r' && is_space(tok[2])) { tok += 2; if (!has_d) cur.dissolve = 1.0f - next_real(&tok); continue; }
		if (strncmp(tok, "map_Kd", 6) == 0 && is_space(tok[6])) {
			texture_name(tok + 7, &cur.diffuse_tex);
			continue;
		}
	}
	flush(); // "flush last material": always, even for an empty file
	return true;
}

} // namespace

std::string load_obj(const char *path, adypt_host_scene *out)
{
	const size_t len = strlen(path);
	if (len == 0) return "[SCENE]Filename invalid";
	std::string base(path);
	{
		const size_t cut = base.find_last_of("/\\");
		base = cut == std::string::npos ? std::string() : base.substr(0, cut + 1);
	}
	LineReader in;
	if (!in.open(path)) return std::string("[SCENE]Failed to load ") + path;

	std::vector<float> v, vn, vt;
	std::vector<MtlEntry> mtls;
	std::map<std::string, int> mtl_index;
	int material = -1;
	std::vector<Corner> face, tri_corners;
	out->tris.clear();

	std::string line;
	while (in.next(&line)) {
		const char *tok = line.c_str();
		tok += strspn(tok, " \t");
		if (*tok == '\0' || *tok == '#') continue;
		if (tok[0] == 'v' && is_space(tok[1])) {
			tok += 2;
			for (int i = 0; i < 3; ++i) v.push_back(next_real(&tok));
			continue;
		}
		if (tok[0] == 'v' && tok[1] == 'n' && is_space(tok[2])) {
			tok += 3;
			for (int i = 0; i < 3; ++i) vn.push_back(next_real(&tok));
			continue;
		}
		if (tok[0] == 'v' && tok[1] == 't' && is_space(tok[2])) {
			tok += 3;
			for (int i = 0; i < 2; ++i) vt.push_back(next_real(&tok));
			continue;
		}
		if (tok[0] == 'f' && is_space(tok[1])) {
			tok += 2;
			tok += strspn(tok, " \t");
			face.clear();
			while (*tok != '\0' && *tok != '\r' && *tok != '\n') {
				Corner c;
				if (!parse_corner(&tok, (int)(v.size() / 3), (int)(vn.size() / 3), (int)(vt.size() / 2), &c))
					return std::string("[SCENE]Failed to load ") + path + " (bad face index)";
				face.push_back(c);
				tok += strspn(tok, " \t\r");
			}
			tri_corners.clear();
			triangulate(face, v, &tri_corners);
			for (size_t k = 0; k + 2 < tri_corners.size(); k += 3) {
				Triangle t;
				memset(&t, 0, sizeof(t));
				t.matid = material;
				for (int c = 0; c < 3; ++c) {
					const Corner &cn = tri_corners[k + (size_t)c];
					if (cn.v < 0 || (size_t)cn.v * 3 + 2 >= v.size()) return std::string("[SCENE]Failed to load ") + path + " (vertex index out of range)";
					memcpy(t.p[c], &v[(size_t)cn.v * 3], 12);
					if (cn.vn != -1 && (size_t)cn.vn * 3 + 2 < vn.size()) memcpy(t.n[c], &vn[(size_t)cn.vn * 3], 12);
					if (cn.vt != -1 && (size_t)cn.vt * 2 + 1 < vt.size()) {
						t.tc[c][0] = vt[(size_t)cn.vt * 2];
						t.tc[c][1] = 1.0f - vt[(size_t)cn.vt * 2 + 1]; // Scene.cpp:68-72
					}
				}
				if (tri_corners[k + 2].vn == -1) { // Scene.cpp:117-123 looks at the LAST corner only
					float n[3];
					flat_normal(t.p[0], t.p[1], t.p[2], n);
					for (int c = 0; c < 3; ++c) memcpy(t.n[c], n, 12);
				}
				out->tris.push_back(t);
			}
			continue;
		}
		if (strncmp(tok, "usemtl", 6) == 0 && is_space(tok[6])) {
			// the name is everything after "usemtl " (tiny_obj_loader.h:1877-1901): extra or trailing blanks are part of it
			const auto it = mtl_index.find(std::string(tok + 7));
			material = it == mtl_index.end() ? -1 : it->second;
			continue;
		}
		if (strncmp(tok, "mtllib", 6) == 0 && is_space(tok[6])) {
			// names are separated by single spaces; they are tried in turn until ONE file opens (tiny_obj_loader.h:1904-1943)
			std::istringstream names(std::string(tok + 7));
			std::string n;
			while (std::getline(names, n, ' '))
				if (load_mtl(base + n, &mtl_index, &mtls)) break;
			continue;
		}
		// g / o / s and everything else do not change triangle order or content
	}

	// OglScene::init_materials (OglScene.cpp:51-91)
	out->mats.clear();
	out->diffuse_textures.clear();
	for (const MtlEntry &m : mtls) {
		Material g;
		memset(&g, 0, sizeof(g));
		if (!m.diffuse_tex.empty()) {
			const std::string full = base + m.diffuse_tex;
			int idx = -1;
			for (size_t i = 0; i < out->diffuse_textures.size(); ++i)
				if (out->diffuse_textures[i] == full) idx = (int)i;
			if (idx < 0) {
				out->diffuse_textures.push_back(full);
				idx = (int)out->diffuse_textures.size() - 1;
			}
			g.dtex = idx; // m_dr/g/b stay unset in the reference; zero here
		} else {
			g.dtex = -1;
			g.dr = m.kd[0]; g.dg = m.kd[1]; g.db = m.kd[2];
		}
		g.er = m.ke[0]; g.eg = m.ke[1]; g.eb = m.ke[2];
		g.sr = m.ks[0]; g.sg = m.ks[1]; g.sb = m.ks[2];
		g.illum = m.illum;
		g.shininess = m.shininess;
		g.dissolve = m.dissolve;
		g.ior = m.ior;
		out->mats.push_back(g);
	}
	if (out->tris.empty()) return std::string("[SCENE]Failed to load ") + path + " (no triangles)";
	return std::string();
}

} // namespace host
} // namespace adypt
