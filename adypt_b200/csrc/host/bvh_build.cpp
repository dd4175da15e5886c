// SBVH (spatial-split BVH, Stich et al. 2009) + collapse to the 8-wide compressed BVH of Ylitie et al. 2017,
// reproducing the reference's builder decisions bit for bit:
//   SBVH  : src/BVH/SBVHBuilder.hpp:62-306, SBVHBuilder.cpp:8-71
//   CWBVH : src/BVH/WideBVHBuilder.cpp:8-273
// Everything that can change an output byte is kept: fp32 evaluation order of every SAH term, glm's
// min/max argument order (sign of zero), strict '<' tie rules, the in-place partition order of the spatial
// split, the float Hungarian assignment, and the quantisation grid. The code structure is this project's own
// (flat work arrays, one cost table, explicit DFS for node emission). Plain IEEE fp32, no contraction.
#include "bvh_build.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <future>
#include <string>
#include <thread>
#include <unistd.h>

namespace adypt {
namespace host {

namespace {

// glm::min(x, y) = (y < x) ? y : x ; glm::max(x, y) = (x < y) ? y : x (dep/glm/detail/func_common.inl:16-30)
inline float gmin(float x, float y) { return (y < x) ? y : x; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }

inline Box empty_box()
{
	return Box{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
}
// AABB::Expand(vec): m_min = min(vec, m_min), m_max = max(vec, m_max) (Shape.hpp:47-50)
inline void grow_point(Box &b, const float p[3])
{
	for (int a = 0; a < 3; ++a) {
		b.lo[a] = gmin(p[a], b.lo[a]);
		b.hi[a] = gmax(p[a], b.hi[a]);
	}
}
// AABB::Expand(aabb): m_min = min(aabb.m_min, m_min) (Shape.hpp:51-54)
inline void grow_box(Box &b, const Box &o)
{
	for (int a = 0; a < 3; ++a) {
		b.lo[a] = gmin(o.lo[a], b.lo[a]);
		b.hi[a] = gmax(o.hi[a], b.hi[a]);
	}
}
// AABB(a, b): min(a.m_min, b.m_min) (Shape.hpp:43-45)
inline Box join(const Box &x, const Box &y)
{
	Box r;
	for (int a = 0; a < 3; ++a) {
		r.lo[a] = gmin(x.lo[a], y.lo[a]);
		r.hi[a] = gmax(x.hi[a], y.hi[a]);
	}
	return r;
}
// AABB::IntersectAABB: m_min = max(m_min, o.m_min), m_max = min(m_max, o.m_max) (Shape.hpp:55-59)
inline void clip(Box &b, const Box &o)
{
	for (int a = 0; a < 3; ++a) {
		b.lo[a] = gmax(b.lo[a], o.lo[a]);
		b.hi[a] = gmin(b.hi[a], o.hi[a]);
	}
}
// AABB::GetArea (Shape.hpp:63-67); also evaluated on empty / inverted boxes, exactly like the reference
inline float area(const Box &b)
{
	const float ex = b.hi[0] - b.lo[0], ey = b.hi[1] - b.lo[1], ez = b.hi[2] - b.lo[2];
	return (ex * (ey + ez) + ey * ez) * 2.0f;
}
inline float center(const Box &b, int a) { return (b.lo[a] + b.hi[a]) * 0.5f; }

struct Ref {
	Box box;
	int32_t tri;
};

struct Span { // NodeSpec: the last `count` references on the stack
	Box box;
	int count;
};

class SbvhBuilder {
public:
	SbvhBuilder(const Triangle *tris, size_t n, const Box &scene, const BvhConfig &cfg, BinaryBvh *out)
	    : tris_(tris), n_(n), scene_(scene), cfg_(cfg), out_(out) {}

	// sub-builder for one subtree: its own reference stack and node array, merged by the parent afterwards
	SbvhBuilder(const SbvhBuilder &parent, std::vector<Ref> &&refs, BinaryBvh *out)
	    : tris_(parent.tris_), n_(parent.n_), scene_(parent.scene_), cfg_(parent.cfg_), out_(out), refs_(std::move(refs)),
	      min_overlap_(parent.min_overlap_), fork_depth_(parent.fork_depth_) {}

	void run()
	{
		out_->nodes.clear();
		out_->leaf_count = 0;
		// Subtrees are independent (a node only touches its own references, which sit on top of the stack), so
		// the top levels fork: each child gets a copy of its reference range and builds into a private node
		// array; the parent appends right subtree then left subtree and rebases their child indices. The
		// output is the sequential builder's, byte for byte, whatever the thread timing.
		const unsigned hw = std::thread::hardware_concurrency();
		fork_depth_ = hw >= 32 ? 6 : hw >= 8 ? 5 : hw >= 4 ? 4 : hw >= 2 ? 2 : 0;
		refs_.reserve(n_ * 2);
		refs_.resize(n_);
		for (size_t i = 0; i < n_; ++i) {
			refs_[i].tri = (int32_t)i;
			refs_[i].box = triangle_box(tris_[i]);
		}
		min_overlap_ = area(scene_) * 1e-5f; // SBVHBuilder.cpp:59
		out_->nodes.reserve(n_ * 2);
		build(Span{scene_, (int)n_}, 0);
		out_->nodes.shrink_to_fit();
	}

private:
	const Triangle *tris_;
	size_t n_;
	Box scene_;
	BvhConfig cfg_;
	BinaryBvh *out_;
	std::vector<Ref> refs_;
	std::vector<Box> suffix_; // boxes accumulated from the right
	float min_overlap_ = 0.f;
	int fork_depth_ = 0;                         // fork subtrees down to this depth ...
	static constexpr int kForkMinRefs = 16384;   // ... while they hold at least this many references
	static constexpr int kBins = 32; // kSpatialBinNum, SBVHBuilder.hpp:16

	float tri_cost(int count) const { return cfg_.triangle_sah * count; } // InstanceConfig.hpp:18
	float node_cost(int count) const { return cfg_.node_sah * count; }    // InstanceConfig.hpp:19

	struct ObjectSplit {
		Box left, right;
		int dim = 0, left_count = 0;
		float sah = FLT_MAX;
	};
	struct SpatialSplit {
		int dim = 0;
		float pos = 0.f, sah = FLT_MAX;
	};

	Ref *span_begin(const Span &s) { return refs_.data() + (refs_.size() - (size_t)s.count); }

	void sort_span(const Span &s, int dim)
	{
		// total order (centre[dim], triangle id): the result does not depend on the sort algorithm
		Ref *b = span_begin(s);
		std::sort(b, b + s.count, [dim](const Ref &l, const Ref &r) {
			const float lc = center(l.box, dim), rc = center(r.box, dim);
			return lc < rc || (lc == rc && l.tri < r.tri);
		});
	}

	// SBVHBuilder.hpp:95-134
	void object_sweep(const Span &s, int dim, float node_sah, ObjectSplit *best)
	{
		sort_span(s, dim);
		const Ref *r = span_begin(s);
		const int n = s.count;
		suffix_.resize((size_t)n);
		suffix_[n - 1] = r[n - 1].box;
		for (int i = n - 2; i >= 1; --i) suffix_[i] = join(r[i].box, suffix_[i + 1]);
		Box left = r[0].box;
		for (int i = 1; i <= n - 1; ++i) {
			const float sah = node_sah + tri_cost(i) * area(left) + tri_cost(n - i) * area(suffix_[i]);
			if (sah < best->sah) {
				best->dim = dim;
				best->left_count = i;
				best->left = left;
				best->right = suffix_[i];
				best->sah = sah;
			}
			grow_box(left, r[i].box);
		}
	}

	// SBVHBuilder.hpp:136-164: clip one reference's triangle against the plane x[dim] = pos
	void split_ref(const Ref &ref, int dim, float pos, Ref *left, Ref *right) const
	{
		left->box = right->box = empty_box();
		left->tri = right->tri = ref.tri;
		const Triangle &t = tris_[ref.tri];
		for (int i = 0; i < 3; ++i) {
			const float *v0 = t.p[i], *v1 = t.p[(i + 1) % 3];
			const float p0 = v0[dim], p1 = v1[dim];
			if (p0 <= pos) grow_point(left->box, v0);
			if (p0 >= pos) grow_point(right->box, v0);
			if ((p0 < pos && pos < p1) || (p1 < pos && pos < p0)) {
				// glm::mix(v0, v1, clamp(a, 0, 1)) = v0 + a * (v1 - v0) (func_common.inl:109)
				float a = (pos - p0) / (p1 - p0);
				a = gmin(gmax(a, 0.0f), 1.0f);
				const float x[3] = {v0[0] + a * (v1[0] - v0[0]), v0[1] + a * (v1[1] - v0[1]), v0[2] + a * (v1[2] - v0[2])};
				grow_point(left->box, x);
				grow_point(right->box, x);
			}
		}
		left->box.hi[dim] = pos;
		clip(left->box, ref.box);
		right->box.lo[dim] = pos;
		clip(right->box, ref.box);
	}

	static int clamp_bin(int v) { return std::min(std::max(v, 0), kBins - 1); }

	// SBVHBuilder.hpp:166-219
	void spatial_sweep(const Span &s, int dim, float node_sah, SpatialSplit *best)
	{
		struct Bin {
			Box box;
			int in, out;
		};
		Bin bins[kBins];
		for (int i = 0; i < kBins; ++i) bins[i] = Bin{empty_box(), 0, 0};
		const float bin_width = (s.box.hi[dim] - s.box.lo[dim]) / kBins, inv_bin_width = 1.0f / bin_width;
		const float base = s.box.lo[dim];
		const Ref *r = span_begin(s);
		Ref cur, l, rr;
		for (int i = 0; i < s.count; ++i) {
			int bin = clamp_bin((int)((r[i].box.lo[dim] - base) * inv_bin_width));
			const int last = clamp_bin((int)((r[i].box.hi[dim] - base) * inv_bin_width));
			bins[bin].in++;
			cur = r[i];
			for (; bin < last; ++bin) {
				split_ref(cur, dim, (bin + 1) * bin_width + base, &l, &rr);
				grow_box(bins[bin].box, l.box);
				cur = rr;
			}
			grow_box(bins[last].box, cur.box);
			bins[last].out++;
		}
		Box suffix[kBins];
		suffix[kBins - 1] = bins[kBins - 1].box;
		for (int i = kBins - 2; i >= 1; --i) suffix[i] = join(bins[i].box, suffix[i + 1]);
		Box left = bins[0].box;
		int left_n = 0, right_n = s.count;
		for (int i = 1; i < kBins; ++i) {
			left_n += bins[i - 1].in;
			right_n -= bins[i - 1].out;
			const float sah = node_sah + tri_cost(left_n) * area(left) + tri_cost(right_n) * area(suffix[i]);
			if (sah < best->sah) {
				best->sah = sah;
				best->dim = dim;
				best->pos = base + i * bin_width;
			}
			grow_box(left, bins[i].box);
		}
	}

	// SBVHBuilder.hpp:230-306. The in-place partition order matters: the unsplit/duplicate decisions below
	// see the straddling references in the order the swaps leave them.
	void apply_spatial(const Span &s, const SpatialSplit &ss, Span *left, Span *right)
	{
		left->box = right->box = empty_box();
		const size_t base = refs_.size() - (size_t)s.count;
		int left_end = 0, right_begin = s.count, right_end = s.count;
		for (int i = 0; i < right_begin; ++i) {
			if (refs_[base + i].box.hi[ss.dim] <= ss.pos) {
				grow_box(left->box, refs_[base + i].box);
				std::swap(refs_[base + i], refs_[base + left_end++]);
			} else if (refs_[base + i].box.lo[ss.dim] >= ss.pos) {
				grow_box(right->box, refs_[base + i].box);
				std::swap(refs_[base + i--], refs_[base + --right_begin]);
			}
		}
		Ref l, r;
		while (left_end < right_begin) {
			split_ref(refs_[base + left_end], ss.dim, ss.pos, &l, &r);
			Box lub = left->box, ldb = left->box, rub = right->box, rdb = right->box;
			grow_box(lub, refs_[base + left_end].box); // whole reference goes left
			grow_box(rub, refs_[base + left_end].box); // whole reference goes right
			grow_box(ldb, l.box);                      // duplicated, clipped halves
			grow_box(rdb, r.box);
			const float lac = tri_cost(left_end), rac = tri_cost(right_end - right_begin);
			const float lbc = tri_cost(1 + left_end), rbc = tri_cost(1 + right_end - right_begin);
			const float unsplit_left = area(lub) * lbc + area(right->box) * rac;
			const float unsplit_right = area(left->box) * lac + area(rub) * rbc;
			const float duplicate = area(ldb) * lbc + area(rdb) * rbc;
			if (unsplit_left < unsplit_right && unsplit_left < duplicate) {
				left->box = lub;
				left_end++;
			} else if (unsplit_right < duplicate) {
				right->box = rub;
				std::swap(refs_[base + left_end], refs_[base + --right_begin]);
			} else {
				refs_.emplace_back();
				left->box = ldb;
				right->box = rdb;
				refs_[base + left_end++] = l;
				refs_[base + right_end++] = r;
			}
		}
		left->count = left_end;
		right->count = right_end - right_begin;
	}

	int emit_leaf(const Span &s)
	{
		BinaryNode n;
		n.box = s.box;
		n.left = -1;
		n.tri = refs_.back().tri;
		refs_.pop_back();
		out_->nodes.push_back(n);
		out_->leaf_count++;
		return (int)out_->nodes.size() - 1;
	}

	// SBVHBuilder.cpp:8-44
	int build(const Span &s, int depth)
	{
		if (s.count == 1) return emit_leaf(s);
		const float node_sah = area(s.box) * node_cost(2);
		ObjectSplit os;
		for (int d = 0; d < 3; ++d) object_sweep(s, d, node_sah, &os);
		SpatialSplit ss;
		if (depth <= cfg_.max_spatial_depth) {
			Box overlap = os.left;
			clip(overlap, os.right);
			if (area(overlap) >= min_overlap_)
				for (int d = 0; d < 3; ++d) spatial_sweep(s, d, node_sah, &ss);
		}
		const int node = (int)out_->nodes.size();
		out_->nodes.emplace_back();
		out_->nodes[node].box = s.box;
		out_->nodes[node].tri = 0;
		Span left{empty_box(), 0}, right{empty_box(), 0};
		if (ss.sah < os.sah) apply_spatial(s, ss, &left, &right);
		if (left.count == 0 || right.count == 0) { // SBVHBuilder.hpp:308-318
			sort_span(s, os.dim);
			left.count = os.left_count;
			left.box = os.left;
			right.count = s.count - os.left_count;
			right.box = os.right;
		}
		if (depth < fork_depth_ && s.count >= kForkMinRefs) {
			// layout on the stack after either split: [left references][right references]
			const size_t base = refs_.size() - (size_t)(left.count + right.count);
			std::vector<Ref> lrefs(refs_.begin() + (long)base, refs_.begin() + (long)base + left.count);
			std::vector<Ref> rrefs(refs_.begin() + (long)base + left.count, refs_.end());
			refs_.resize(base);
			BinaryBvh lout, rout;
			SbvhBuilder lb(*this, std::move(lrefs), &lout), rb(*this, std::move(rrefs), &rout);
			auto fut = std::async(std::launch::async, [&rb, right, depth]() { rb.build(right, depth + 1); });
			lb.build(left, depth + 1);
			fut.get();
			auto append = [this](const BinaryBvh &sub) {
				const int start = (int)out_->nodes.size();
				out_->nodes.reserve(out_->nodes.size() + sub.nodes.size());
				for (BinaryNode n : sub.nodes) {
					if (n.left != -1) n.left += start;
					out_->nodes.push_back(n);
				}
				out_->leaf_count += sub.leaf_count;
				return start;
			};
			append(rout); // right child lands at node + 1
			out_->nodes[node].left = append(lout);
			return node;
		}
		build(right, depth + 1); // right child lands at node + 1
		const int l = build(left, depth + 1);
		out_->nodes[node].left = l;
		return node;
	}
};

// ------------------------------------------------------------------------------------------------
// wide collapse

enum CostKind : int32_t { kInternal = 0, kLeaf = 1, kDistribute = 2 };
struct Cost {
	float sah;
	int32_t kind;
	int32_t split[2]; // roots given to the left / right binary child
};

class WideBuilder {
public:
	WideBuilder(const BinaryBvh &b, const BvhConfig &cfg, WideBvh *out) : b_(b), cfg_(cfg), out_(out) {}

	bool run()
	{
		out_->nodes.clear();
		out_->tri_indices.clear();
		if (b_.nodes.empty() || b_.nodes[0].left == -1) return false; // reference: out-of-bounds read
		cost_.assign(b_.nodes.size() * 7, Cost{0.f, 0, {0, 0}});
		analyse(0);
		out_->nodes.emplace_back();
		memset(&out_->nodes[0], 0, sizeof(Node));
		out_->tri_indices.reserve((size_t)b_.leaf_count);
		emit(0, 0);
		out_->nodes.shrink_to_fit();
		cost_.clear();
		cost_.shrink_to_fit();
		return true;
	}

private:
	const BinaryBvh &b_;
	BvhConfig cfg_;
	WideBvh *out_;
	std::vector<Cost> cost_; // [node][i-1]: cheapest way to present the subtree as <= i roots, i = 1..7

	Cost &C(int node, int i) { return cost_[(size_t)node * 7 + (size_t)(i - 1)]; }
	float tri_cost(int count) const { return cfg_.triangle_sah * count; }
	float node_cost(int count) const { return cfg_.node_sah * count; }
	bool is_leaf(int n) const { return b_.nodes[n].left == -1; }

	// WideBVHBuilder.cpp:24-91
	int analyse(int n)
	{
		const float a = area(b_.nodes[n].box);
		if (is_leaf(n)) {
			for (int i = 1; i <= 7; ++i) {
				C(n, i).sah = tri_cost(1) * a;
				C(n, i).kind = kLeaf;
			}
			return 1;
		}
		const int l = b_.nodes[n].left, r = n + 1;
		const int rc = analyse(r), lc = analyse(l);
		const int tri_count = rc + lc;
		{
			const float c_leaf = tri_count <= 3 ? a * tri_cost(tri_count) : FLT_MAX;
			float c_internal = FLT_MAX;
			const float node_sah = a * node_cost(8);
			for (int k = 1; k < 8; ++k) {
				const float v = node_sah + C(l, k).sah + C(r, 8 - k).sah;
				if (v < c_internal) {
					c_internal = v;
					C(n, 1).split[0] = k;
					C(n, 1).split[1] = 8 - k;
				}
			}
			if (c_leaf < c_internal) {
				C(n, 1).sah = c_leaf;
				C(n, 1).kind = kLeaf;
			} else {
				C(n, 1).sah = c_internal;
				C(n, 1).kind = kInternal;
			}
		}
		for (int i = 2; i <= 7; ++i) {
			float c_dist = FLT_MAX;
			for (int k = 1; k < i; ++k) {
				const float v = C(l, k).sah + C(r, i - k).sah;
				if (v < c_dist) {
					c_dist = v;
					C(n, i).split[0] = k;
					C(n, i).split[1] = i - k;
				}
			}
			if (c_dist < C(n, i - 1).sah) {
				C(n, i).sah = c_dist;
				C(n, i).kind = kDistribute;
			} else
				C(n, i) = C(n, i - 1);
		}
		return tri_count;
	}

	// WideBVHBuilder.cpp:93-108
	void gather_children(int n, int i, int *count, int out[8])
	{
		const int child[2] = {b_.nodes[n].left, n + 1};
		const int share[2] = {C(n, i).split[0], C(n, i).split[1]};
		for (int c = 0; c < 2; ++c) {
			if (C(child[c], share[c]).kind == kDistribute) gather_children(child[c], share[c], count, out);
			else out[(*count)++] = child[c];
		}
	}

	// WideBVHBuilder.cpp:110-118: right subtree first
	int gather_leaves(int n)
	{
		if (is_leaf(n)) {
			out_->tri_indices.push_back(b_.nodes[n].tri);
			return 1;
		}
		const int r = gather_leaves(n + 1);
		return r + gather_leaves(b_.nodes[n].left);
	}

	// Min-cost assignment of n children (rows) to the 8 slots (columns), the O(n^2 m) potentials method in
	// fp32 exactly as WideBVHBuilder.cpp:120-163 runs it (same update order => same ties).
	static void assign_slots(const float cost[8][8], int n, int slot_of_child[8])
	{
		const float kInf = 1e12f;
		float u[9], v[9], minv[9];
		int p[9], way[9];
		bool used[9];
		for (int j = 0; j < 9; ++j) { u[j] = 0.f; v[j] = 0.f; p[j] = 0; way[j] = 0; }
		for (int i = 1; i <= n; ++i) {
			p[0] = i;
			int j0 = 0;
			for (int j = 0; j < 9; ++j) { minv[j] = kInf; used[j] = false; }
			do {
				used[j0] = true;
				const int i0 = p[j0];
				int j1 = 0;
				float delta = kInf;
				for (int j = 1; j <= 8; ++j)
					if (!used[j]) {
						const float cur = cost[i0 - 1][j - 1] - u[i0] - v[j];
						if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
						if (minv[j] < delta) { delta = minv[j]; j1 = j; }
					}
				for (int j = 0; j <= 8; ++j) {
					if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
					else minv[j] -= delta;
				}
				j0 = j1;
			} while (p[j0] != 0);
			do {
				const int j1 = way[j0];
				p[j0] = p[j1];
				j0 = j1;
			} while (j0);
		}
		for (int j = 1; j <= 8; ++j)
			if (p[j] != 0) slot_of_child[p[j] - 1] = j - 1;
	}

	// WideBVHBuilder.cpp:165-273
	void emit(int wide, int bin)
	{
		int child[8], n_child = 0;
		gather_children(bin, 1, &n_child, child);
		const Box &box = b_.nodes[bin].box;
		float cell[3];
		{
			Node &cur = out_->nodes[wide];
			cur.px = box.lo[0];
			cur.py = box.lo[1];
			cur.pz = box.lo[2];
			const float k = (float)(1.0 / (double)((1 << 8) - 1));
			uint32_t e[3];
			for (int a = 0; a < 3; ++a) {
				const float c = (box.hi[a] - box.lo[a]) * k;
				const int ex = c == 0 ? -128 : (int)std::ceil(std::log2(c));
				cell[a] = exp2f((float)ex);
				uint32_t bits;
				memcpy(&bits, &cell[a], 4);
				e[a] = (uint8_t)(bits >> 23);
			}
			cur.head_w = e[0] | (e[1] << 8) | (e[2] << 16); // imask (top byte) filled below
		}
		int slot_of_child[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		{
			float cost[8][8];
			for (int i = 0; i < n_child; ++i) {
				const Box &cb = b_.nodes[child[i]].box;
				const float dx = center(cb, 0) - center(box, 0), dy = center(cb, 1) - center(box, 1), dz = center(cb, 2) - center(box, 2);
				for (int j = 0; j < 8; ++j) cost[i][j] = ((j & 1) ? -dx : dx) + ((j & 2) ? -dy : dy) + ((j & 4) ? -dz : dz);
			}
			assign_slots(cost, n_child, slot_of_child);
		}
		int at_slot[8];
		for (int i = 0; i < 8; ++i) at_slot[i] = -1;
		for (int i = 0; i < n_child; ++i) at_slot[slot_of_child[i]] = child[i];

		const uint32_t child_base = (uint32_t)out_->nodes.size(), tri_base = (uint32_t)out_->tri_indices.size();
		uint8_t meta[8], q[6][8];
		memset(meta, 0, sizeof(meta));
		memset(q, 0, sizeof(q));
		uint32_t imask = 0;
		for (int i = 0; i < 8; ++i) {
			const int s = at_slot[i];
			if (s < 0) continue;
			const Box &cb = b_.nodes[s].box;
			for (int a = 0; a < 3; ++a) {
				uint32_t lo = (uint32_t)std::floor((cb.lo[a] - box.lo[a]) / cell[a]);
				uint32_t hi = (uint32_t)std::ceil((cb.hi[a] - box.lo[a]) / cell[a]);
				lo = std::min(lo, 255u);
				hi = std::min(hi, 255u);
				q[a][i] = (uint8_t)lo;
				q[3 + a][i] = (uint8_t)hi;
			}
			if (C(s, 1).kind == kLeaf) {
				const uint32_t first = (uint32_t)out_->tri_indices.size() - tri_base;
				const int cnt = gather_leaves(s);
				uint8_t m = 0;
				if (cnt == 1) m = 0x20;
				if (cnt == 2) m = 0x60;
				if (cnt == 3) m = 0xE0;
				meta[i] = (uint8_t)(m | first);
			} else {
				const uint32_t ordinal = (uint32_t)out_->nodes.size() - child_base; // compacted, NOT the slot
				out_->nodes.emplace_back();
				memset(&out_->nodes.back(), 0, sizeof(Node));
				meta[i] = (uint8_t)(0x20u | (ordinal + 24u));
				imask |= 1u << ordinal;
			}
		}
		{
			Node &cur = out_->nodes[wide]; // re-fetch: the vector may have grown
			cur.head_w |= imask << 24;
			cur.child_base = child_base;
			cur.tri_base = tri_base;
			auto pack = [](const uint8_t *b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); };
			cur.meta_lo = pack(meta); cur.meta_hi = pack(meta + 4);
			cur.lox_lo = pack(q[0]); cur.lox_hi = pack(q[0] + 4);
			cur.loy_lo = pack(q[1]); cur.loy_hi = pack(q[1] + 4);
			cur.loz_lo = pack(q[2]); cur.loz_hi = pack(q[2] + 4);
			cur.hix_lo = pack(q[3]); cur.hix_hi = pack(q[3] + 4);
			cur.hiy_lo = pack(q[4]); cur.hiy_hi = pack(q[4] + 4);
			cur.hiz_lo = pack(q[5]); cur.hiz_hi = pack(q[5] + 4);
		}
		for (int i = 0; i < n_child; ++i)
			if (C(child[i], 1).kind == kInternal)
				emit((int)(child_base + (uint32_t)(meta[slot_of_child[i]] & 0x1fu) - 24u), child[i]);
	}
};

} // namespace

Box triangle_box(const Triangle &t)
{
	// Triangle::GetAABB: min(p0, min(p1, p2)) / max(p0, max(p1, p2)) (Shape.hpp:81-87)
	Box b;
	for (int a = 0; a < 3; ++a) {
		b.lo[a] = gmin(t.p[0][a], gmin(t.p[1][a], t.p[2][a]));
		b.hi[a] = gmax(t.p[0][a], gmax(t.p[1][a], t.p[2][a]));
	}
	return b;
}

Box scene_box(const Triangle *tris, size_t n)
{
	Box b = empty_box();
	for (size_t i = 0; i < n; ++i) grow_box(b, triangle_box(tris[i])); // Scene.cpp:125
	return b;
}

void build_binary(const Triangle *tris, size_t n_tris, const Box &scene, const BvhConfig &cfg, BinaryBvh *out)
{
	SbvhBuilder(tris, n_tris, scene, cfg, out).run();
}

bool build_wide(const BinaryBvh &sbvh, const BvhConfig &cfg, WideBvh *out) { return WideBuilder(sbvh, cfg, out).run(); }

// ------------------------------------------------------------------------------------------------
static const char kMagic[] = "CWBVH_1.0"; // WideBVH.hpp:32, written with its terminating NUL

bool save_bvh_file(const char *path, const WideBvh &bvh, const BvhConfig &cfg)
{
	// written next to the target and renamed into place, so a concurrent reader (several ranks sharing one
	// cache directory) never sees a half-written file
	const std::string tmp = std::string(path) + ".tmp." + std::to_string((long)getpid());
	FILE *f = fopen(tmp.c_str(), "wb");
	if (!f) return false;
	const uint32_t n_idx = (uint32_t)bvh.tri_indices.size();
	bool ok = fwrite(kMagic, 1, sizeof(kMagic), f) == sizeof(kMagic);
	ok = ok && fwrite(&cfg, sizeof(cfg), 1, f) == 1;
	ok = ok && fwrite(&n_idx, 4, 1, f) == 1;
	ok = ok && (n_idx == 0 || fwrite(bvh.tri_indices.data(), 4, n_idx, f) == n_idx);
	ok = ok && (bvh.nodes.empty() || fwrite(bvh.nodes.data(), sizeof(Node), bvh.nodes.size(), f) == bvh.nodes.size());
	ok = (fclose(f) == 0) && ok;
	if (ok) ok = rename(tmp.c_str(), path) == 0;
	if (!ok) remove(tmp.c_str());
	return ok;
}

bool load_bvh_file(const char *path, const BvhConfig &expected, WideBvh *out)
{
	out->nodes.clear();
	out->tri_indices.clear();
	FILE *f = fopen(path, "rb");
	if (!f) return false;
	std::vector<uint8_t> buf;
	uint8_t chunk[1 << 16];
	size_t got;
	while ((got = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
	fclose(f);
	size_t pos = sizeof(kMagic);
	if (buf.size() < pos + sizeof(BvhConfig) + 4 || memcmp(buf.data(), kMagic, sizeof(kMagic)) != 0) return false;
	BvhConfig cfg;
	memcpy(&cfg, buf.data() + pos, sizeof(cfg));
	pos += sizeof(cfg);
	// reused only when the three build parameters match (WideBVH.cpp:42-45)
	if (cfg.node_sah != expected.node_sah || cfg.triangle_sah != expected.triangle_sah || cfg.max_spatial_depth != expected.max_spatial_depth)
		return false;
	uint32_t n_idx;
	memcpy(&n_idx, buf.data() + pos, 4);
	pos += 4;
	if (buf.size() < pos + (size_t)n_idx * 4) return false; // truncated (the reference reads what is there)
	out->tri_indices.resize(n_idx);
	if (n_idx) memcpy(out->tri_indices.data(), buf.data() + pos, (size_t)n_idx * 4);
	pos += (size_t)n_idx * 4;
	const size_t n_nodes = (buf.size() - pos) / sizeof(Node);
	out->nodes.resize(n_nodes);
	if (n_nodes) memcpy(out->nodes.data(), buf.data() + pos, n_nodes * sizeof(Node));
	return true;
}

} // namespace host
} // namespace adypt
