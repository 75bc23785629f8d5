// chaining.cpp -- x-drop ungapped extension of seed hits and greedy diagonal-graph chaining into approximate HSPs.
//
// This is the host bridge between the two GPU stages; its output (d_min, d_max per chain) defines the DP bands, so
// every arithmetic step, tie and container-order effect of the reference is kept:
//   dp/ungapped_align.cpp:150-214          xdrop_ungapped
//   chaining/greedy_align.cpp:58-127       DiagGraph::load / sort / prune
//   chaining/greedy_align.cpp:150-214      get_hgap_link / get_vgap_link / get_link
//   chaining/greedy_align.cpp:220-273      Aligner::get_approximate_link
//   chaining/greedy_align.cpp:275-366      Aligner::forward_pass  (std::map window, erase-while-iterating)
//   chaining/diag_graph.h:27-193           DiagonalNode, Edge, add_edge, get_edge, prefix_score
//   chaining/backtrace.cpp:36-76,78-173,269-356   disjoint, backtrace_old, top-node loop
//   chaining/greedy_align.cpp:426-497      merge_score / merge / merge_hsps / Chaining::run
// The anchor (max_diag) bookkeeping of the reference is not carried: it only feeds the opt-in anchored SWIPE.
#include "chaining.h"
#include "../../../include/dmnd_b200.h"
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <utility>

namespace dmnd {

Segment xdrop_ungapped(const Scoring& sc, const int8_t* query, const int8_t* cbs, const int8_t* subject, int qa, int sa) {
	const int xdrop = sc.raw_ungapped_xdrop;
	int score = 0, st = 0, n = 1, delta = 0, len = 0;
	int q = qa - 1, s = sa - 1, ql, sl;
	while (score - st < xdrop && (ql = query[q] & 31) != DMND_DELIMITER && (sl = subject[s] & 31) != DMND_DELIMITER) {
		st += sc.score(ql, sl);
		if (cbs) st += cbs[q];
		if (st > score) { score = st; delta = n; }
		--q; --s; ++n;
	}
	q = qa; s = sa; st = score; n = 1;
	while (score - st < xdrop && (ql = query[q] & 31) != DMND_DELIMITER && (sl = subject[s] & 31) != DMND_DELIMITER) {
		st += sc.score(ql, sl);
		if (cbs) st += cbs[q];
		if (st > score) { score = st; len = n; }
		++q; ++s; ++n;
	}
	return Segment{ qa - delta, sa - delta, len + delta, score };
}

namespace {

struct Node : Segment {
	int link_idx, prefix_score, path_max, path_min;
	explicit Node(const Segment& s) : Segment(s), link_idx(-1), prefix_score(s.score), path_max(s.score), path_min(s.score) {}
	int rel_score() const { return prefix_score == path_max ? prefix_score : prefix_score - path_min; }
};
struct Edge {
	int prefix_score, path_max, j, path_min, prefix_score_begin;
	unsigned node_in, node_out;
};
struct Link {
	int subject_pos1 = -1, query_pos1 = 0, subject_pos2 = 0, query_pos2 = 0, score1 = 0, score2 = 0;
	void reset() { subject_pos1 = -1; score1 = 0; score2 = 0; }
	void transpose() { std::swap(subject_pos1, query_pos1); std::swap(subject_pos2, query_pos2); }
};
struct Seq {
	const int8_t* p;
	int operator[](int i) const { return p[i] & 31; }
};

constexpr size_t NO_EDGE = (size_t)-1;
constexpr double SPACE_PENALTY = 0.1;
constexpr int LINK_PADDING = 10, REVERSE_LINK_MIN_OVERHANG = 10;

struct Graph {
	std::vector<Node> nodes;
	std::vector<Edge> edges;

	// diag_graph.h:123-138
	void add_edge(const Edge& e) {
		for (size_t k = e.node_in + 1; k < nodes.size(); ++k) {
			if (nodes[k].link_idx == -1) break;
			++nodes[k].link_idx;
		}
		Node& d = nodes[e.node_in];
		if (e.prefix_score > d.prefix_score) { d.prefix_score = e.prefix_score; d.path_max = e.path_max; d.path_min = e.path_min; }
		edges.insert(edges.begin() + d.link_idx++, e);
	}
	// diag_graph.h:140-161 ; returns edge index or NO_EDGE
	size_t get_edge(size_t node, int j) const {
		const Node& d = nodes[node];
		if (d.score == 0) return (size_t)((ptrdiff_t)d.link_idx - 1);
		if (edges.empty()) return NO_EDGE;
		int max_score = d.score;
		ptrdiff_t max_i = -1;
		for (ptrdiff_t i = (ptrdiff_t)d.link_idx - 1; i >= 0 && edges[(size_t)i].node_in == node; --i)
			if (edges[(size_t)i].j < j && edges[(size_t)i].prefix_score > max_score) { max_i = i; max_score = edges[(size_t)i].prefix_score; }
		return max_i >= 0 ? (size_t)max_i : NO_EDGE;
	}
	bool edge_ok(size_t e) const { return e != NO_EDGE && e < edges.size(); }
	// diag_graph.h:163-169
	int prefix_score(size_t node, int j, int& path_max, int& path_min) const {
		const size_t i = get_edge(node, j);
		const bool none = !edge_ok(i);
		path_max = none ? nodes[node].score : std::max(nodes[node].score, edges[i].path_max);
		path_min = none ? nodes[node].score : edges[i].path_min;
		return none ? nodes[node].score : std::max(nodes[node].score, edges[i].prefix_score);
	}
};

// Ordered diagonal -> node window of Aligner::forward_pass (a std::map<int, unsigned> in the reference).  A sorted
// vector reproduces its ordered iteration and erase-while-iterating behaviour with indices instead of node iterators.
struct Window {
	std::vector<std::pair<int, unsigned>> v;
	void clear() { v.clear(); }
	size_t size() const { return v.size(); }
	// index of `key`, inserting (key, val) when absent (std::map::find + insert)
	size_t find_or_insert(int key, unsigned val) {
		size_t lo = 0, hi = v.size();
		while (lo < hi) { const size_t mid = (lo + hi) / 2; if (v[mid].first < key) lo = mid + 1; else hi = mid; }
		if (lo == v.size() || v[lo].first != key) v.insert(v.begin() + (ptrdiff_t)lo, std::make_pair(key, val));
		return lo;
	}
	void erase(size_t i) { v.erase(v.begin() + (ptrdiff_t)i); }
};

struct Chainer {
	const Scoring& sc;
	Seq query, subject;
	int qlen, slen;
	Graph& g;
	Window& window;

	int score_range(Seq q, Seq s, int i, int j, int j_end) const {
		int v = 0;
		while (j < j_end) { v += sc.score(q[i], s[j]); ++i; ++j; }
		return v;
	}
	// greedy_align.cpp:150-199
	int hgap_link(const Segment& d1, const Segment& d2, Seq q, Seq s, Link& l, int padding) const {
		const int d = d1.diag() - d2.diag(),
			j2_end = std::min(std::max(d2.j, d1.subject_last() + d + 1 + padding), d2.subject_last());
		int j1;
		bool space;
		if (d1.subject_last() < d2.j - d - 1) { j1 = d1.subject_last(); space = true; }
		else { j1 = std::max(d2.j - d - 1 - padding, d1.j); space = false; }
		int j2 = j1 + d + 1, i1 = d1.i + (j1 - d1.j), i2 = i1 + 1;
		if (j2 > d2.subject_last()) { l.reset(); return INT_MIN; }
		int score1 = 0, score2 = score_range(q, s, i2, j2, d2.j) + d2.score - score_range(q, s, d2.i, d2.j, j2);
		int max_score = INT_MIN;
		for (;;) {
			if (score1 + score2 > max_score) {
				max_score = score1 + score2;
				l.query_pos1 = i1; l.subject_pos1 = j1; l.query_pos2 = i2; l.subject_pos2 = j2; l.score1 = score1; l.score2 = score2;
			}
			score2 -= sc.score(q[i2], s[j2]);
			++i1; ++i2; ++j1; ++j2;
			if (j2 > j2_end) break;
			score1 += sc.score(q[i1], s[j1]);
		}
		const int j1_end = j2_end - d;
		if (space) l.score1 += d1.score;
		else l.score1 += d1.score - score_range(q, s, d1.diag() + j1_end, j1_end, d1.subject_end())
			+ score_range(q, s, d1.query_end(), d1.subject_end(), j1_end) - score1;
		return max_score;
	}
	static Segment transposed(const Segment& s) { return Segment{ s.j, s.i, s.len, s.score }; }
	int get_link(const Segment& d1, const Segment& d2, Link& l, int padding) const {
		if (d1.diag() < d2.diag()) {
			const int s = hgap_link(transposed(d1), transposed(d2), subject, query, l, padding);
			l.transpose();
			return s;
		}
		return hgap_link(d1, d2, query, subject, l, padding);
	}

	// greedy_align.cpp:220-273
	int approximate_link(int d_idx, int e_idx, double space_penalty) {
		Node& d = g.nodes[(size_t)d_idx];
		Node& e = g.nodes[(size_t)e_idx];
		const int shift = d.diag() - e.diag();
		const int gap_score = shift != 0 ? -sc.gap_open - std::abs(shift) * sc.gap_extend : 0;
		const int space = shift > 0 ? d.j - e.subject_last() : d.i - e.query_last();
		int prefix_score = 0, link_j = 0, diff1 = 0, path_max = 0, path_min = 0, prefix_score_begin = 0;
		if (space <= 0 || space_penalty == 0.0) {
			const size_t edge = g.get_edge((size_t)d_idx, d.j);
			if (g.edge_ok(edge) && g.edges[edge].prefix_score > e.prefix_score + gap_score + d.score) return 0;
			Link link;
			if (get_link(e, d, link, LINK_PADDING) > 0) {
				diff1 = e.score - link.score1;
				const int prefix_e = g.prefix_score((size_t)e_idx, link.subject_pos1, path_max, path_min);
				prefix_score = prefix_e - diff1 + gap_score + link.score2;
				const size_t edge2 = g.get_edge((size_t)d_idx, link.subject_pos2);
				if (g.edge_ok(edge2) && g.edges[edge2].prefix_score > prefix_score) return 0;
				prefix_score_begin = prefix_score - link.score2;
				path_min = std::min(path_min, prefix_score - link.score2);
				if (prefix_e == path_max) path_max -= diff1;
				link_j = link.subject_pos2;
			}
		}
		else {
			prefix_score = e.prefix_score + gap_score - int(space_penalty * std::max(space - 1, 0)) + d.score;
			const size_t edge = g.get_edge((size_t)d_idx, d.j);
			if (g.edge_ok(edge) && g.edges[edge].prefix_score > prefix_score) return 0;
			prefix_score_begin = prefix_score - d.score;
			path_max = e.path_max;
			path_min = e.path_min;
			path_min = std::min(path_min, prefix_score - d.score);
			link_j = d.j;
		}
		if (prefix_score > d.score) {
			path_max = std::max(path_max, prefix_score);
			g.add_edge(Edge{ prefix_score, path_max, link_j, prefix_score == path_max ? prefix_score : path_min, prefix_score_begin,
				(unsigned)d_idx, (unsigned)e_idx });
		}
		return prefix_score;
	}

	// greedy_align.cpp:275-366
	void forward_pass(double space_penalty) {
		window.clear();
		for (unsigned node = 0; node < g.nodes.size(); ++node) {
			g.nodes[node].link_idx = (int)g.edges.size();
			const int dd = g.nodes[node].diag();
			size_t i = window.find_or_insert(dd, node);  // the map entry of this diagonal (value = last node seen on it)
			size_t j = i;
			int max_j = 0;
			if (i != 0) {
				do {
					--j;
					const Node& d = g.nodes[node];
					const Node& e = g.nodes[window.v[j].second];
					if (e.prefix_score - int(space_penalty * (std::max(d.j - e.subject_end(), 0))) <= 0) {
						// std::map::erase(j): the successor keeps its position in the order; with indices it slides into slot j
						const bool was_first = j == 0;
						window.erase(j);
						--i;
						if (was_first) break;
						continue;  // `j = k; continue;` of the reference: the loop test holds (j > 0) and --j steps before the erased slot
					}
					if (e.subject_end() < max_j) continue;
					const unsigned e_idx = window.v[j].second;
					approximate_link((int)node, (int)e_idx, space_penalty);
					{
						const Node& d2 = g.nodes[node];
						const Node& e2 = g.nodes[e_idx];
						max_j = std::max(max_j, std::min(d2.j, e2.subject_end()));
						if (e2.subject_end() - (d2.subject_end() - std::min(e2.diag() - d2.diag(), 0)) >= REVERSE_LINK_MIN_OVERHANG)
							approximate_link((int)e_idx, (int)node, space_penalty);
					}
				} while (j != 0);
			}
			j = i;
			if (window.v[j].second == node) ++j;
			int max_i = 0;
			while (j != window.size()) {
				const Node& d = g.nodes[node];
				const Node& e = g.nodes[window.v[j].second];
				if (e.prefix_score - int(space_penalty * (std::max(d.j - e.subject_end(), 0))) <= 0 && j != i) {
					window.erase(j);  // the successor slides into slot j
					continue;
				}
				if (e.query_end() < max_i) { ++j; continue; }
				const unsigned e_idx = window.v[j].second;
				approximate_link((int)node, (int)e_idx, space_penalty);
				{
					const Node& d2 = g.nodes[node];
					const Node& e2 = g.nodes[e_idx];
					if (e2.i < d2.i) max_i = std::max(max_i, std::min(e2.query_end(), d2.i));
					if (e2.subject_end() - (d2.subject_end() - std::min(e2.diag() - d2.diag(), 0)) >= REVERSE_LINK_MIN_OVERHANG)
						approximate_link((int)e_idx, (int)node, space_penalty);
				}
				++j;
			}
			window.v[i].second = node;
		}
	}

	static double overlap_factor(int b0, int e0, int b1, int e1) {  // Interval(b0,e0).overlap_factor(Interval(b1,e1))
		const int ib = std::max(b0, b1), ie = std::min(e0, e1);
		const unsigned ov = (unsigned)(ie > ib ? ie - ib : 0);
		const int len = e0 > b0 ? e0 - b0 : 0;
		return (double)ov / (double)len;
	}
	// backtrace.cpp:36-76 (both overloads: t = candidate with ranges + score)
	template<typename It>
	static bool disjoint(It begin, It end, int qb, int qe, int sb, int se, int score, int cutoff) {
		for (; begin != end; ++begin) {
			const double ot = overlap_factor(sb, se, begin->s_begin, begin->s_end), oq = overlap_factor(qb, qe, begin->q_begin, begin->q_end);
			if ((1.0 - std::min(ot, oq)) * score / begin->score >= 0.5) continue;  // chaining_stacked_hsp_ratio, config.cpp:603
			if ((1.0 - std::max(ot, oq)) * score < cutoff) return false;
		}
		return true;
	}

	// backtrace.cpp:78-173 with out == nullptr
	bool backtrace_rec(size_t node, int j_end, Chain& t, int score_max, int score_min, int max_shift, unsigned& next) const {
		const Node& d = g.nodes[node];
		const size_t f = g.get_edge(node, j_end);
		bool at_end = !g.edge_ok(f);
		const int prefix_score = at_end ? d.score : g.edges[f].prefix_score;
		if (prefix_score > score_max) return false;
		score_min = std::min(score_min, at_end ? 0 : g.edges[f].prefix_score_begin);
		if (!at_end) {
			const Edge& ed = g.edges[f];
			const Node& e = g.nodes[ed.node_out];
			const int shift = d.diag() - e.diag();
			const int j = ed.j;
			if (std::abs(shift) <= max_shift) {
				const bool bt = backtrace_rec(ed.node_out, shift > 0 ? j : j + shift, t, score_max, score_min, max_shift, next);
				if (!bt) {
					if (ed.prefix_score_begin > score_min) return false;
					at_end = true;
				}
			}
			else { next = ed.node_out; at_end = true; }
		}
		if (at_end) { t.q_begin = d.i; t.s_begin = d.j; t.score = score_max - score_min; }
		const int dd = d.diag();
		t.d_max = std::max(t.d_max, dd);
		t.d_min = std::min(t.d_min, dd);
		return true;
	}

	void run(std::vector<Chain>& ts, double space_penalty, int cutoff, int max_shift) {
		// greedy_align.cpp:368-394 (chaining_maxnodes unset; chaining_len_cap 2.0, chaining_min_nodes 200)
		if (g.nodes.size() > 200) {
			std::sort(g.nodes.begin(), g.nodes.end(), [](const Segment& x, const Segment& y) { return x.score > y.score; });
			const double cap = qlen * 2.0;
			double total_len = 0.0;
			auto it = g.nodes.begin();
			while (it < g.nodes.end() && total_len < cap) { total_len += it->len; ++it; }
			g.nodes.erase(std::max(g.nodes.begin() + 200, it), g.nodes.end());
		}
		std::sort(g.nodes.begin(), g.nodes.end(), [](const Segment& x, const Segment& y) { return x.j < y.j || (x.j == y.j && x.i < y.i); });
		prune();
		forward_pass(space_penalty);
		// backtrace.cpp:327-356
		static thread_local std::vector<const Node*> top;  // scratch reused across calls (no allocation in steady state)
		top.clear();
		for (const Node& d : g.nodes)
			if (d.rel_score() >= cutoff) top.push_back(&d);
		std::sort(top.begin(), top.end(), [](const Node* x, const Node* y) { return x->rel_score() > y->rel_score(); });
		const size_t t_begin = ts.size();
		for (const Node* n : top) {
			if (!disjoint(ts.begin() + (ptrdiff_t)t_begin, ts.end(), n->i, n->i + n->len, n->j, n->j + n->len, n->score, cutoff)) continue;
			// backtrace.cpp:292-325
			size_t top_node = (size_t)(n - g.nodes.data());
			unsigned next;
			int max_j = slen;
			do {
				Chain t;
				next = UINT_MAX;
				if (top_node != NO_EDGE) {
					const Node& d = g.nodes[top_node];
					t.s_end = d.subject_end();
					t.q_end = d.query_end();
					backtrace_rec(top_node, std::min(d.subject_end(), max_j), t, d.prefix_score, d.prefix_score, max_shift, next);
				}
				if (t.score > 0) max_j = t.s_begin;
				if (t.score >= cutoff && disjoint(ts.begin() + (ptrdiff_t)t_begin, ts.end(), t.q_begin, t.q_end, t.s_begin, t.s_end, t.score, cutoff))
					ts.push_back(t);
				top_node = next;
			} while (next != UINT_MAX);
		}
	}

	// greedy_align.cpp:106-127
	void prune() {
		static thread_local std::vector<Node> finished, win;  // scratch reused across calls
		finished.clear(); win.clear();
		finished.reserve(g.nodes.size());
		for (const Node& d : g.nodes) {
			size_t n = 0;
			for (size_t k = 0; k < win.size();) {
				if (win[k].subject_end() > d.j) {
					if (win[k].score >= d.score && win[k].j <= d.j && win[k].subject_end() >= d.subject_end()) ++n;
					++k;
				}
				else { finished.push_back(win[k]); win.erase(win.begin() + (ptrdiff_t)k); }
			}
			if (n <= 8) win.push_back(d);  // config.chaining_range_cover
		}
		for (const Node& d : win) finished.push_back(d);
		g.nodes.swap(finished);  // both buffers stay alive for the next call
	}
};

// greedy_align.cpp:426-437
int merge_score(const Chain& h1, const Chain& h2) {
	const int gq = h2.q_begin - h1.q_end, gt = h2.s_begin - h1.s_end;
	if (gq < 0 || gt < 0) return 0;
	const int s = h1.score + h2.score;
	if (gq > gt) return int(s - gq * 0.5 - gt * SPACE_PENALTY);
	return int(s - gt * 0.5 - gq * SPACE_PENALTY);
}
Chain merge(const Chain& h1, const Chain& h2) {
	Chain h;
	h.d_max = std::max(h1.d_max, h2.d_max);
	h.d_min = std::min(h1.d_min, h2.d_min);
	h.q_begin = h1.q_begin; h.q_end = h2.q_end;
	h.s_begin = h1.s_begin; h.s_end = h2.s_end;
	h.score = merge_score(h1, h2);
	return h;
}

}  // namespace

void chain_segments(const Scoring& sc, const int8_t* query, int qlen, const int8_t* subject, int slen,
                    const std::vector<Segment>& segs, std::vector<Chain>& out) {
	out.clear();
	if (segs.size() == 1) {
		const Segment& s = segs[0];
		Chain c;
		c.d_min = c.d_max = s.diag();
		c.score = s.score;
		c.q_begin = s.i; c.q_end = s.i + s.len; c.s_begin = s.j; c.s_end = s.j + s.len;
		out.push_back(c);
		return;
	}
	static thread_local Graph graph;    // storage reused across calls (the reference keeps them thread_local as well,
	static thread_local Window window;  // chaining/greedy_align.cpp:418-419)
	graph.nodes.clear(); graph.edges.clear();
	Chainer ch{ sc, Seq{ query }, Seq{ subject }, qlen, slen, graph, window };
	// DiagGraph::load, greedy_align.cpp:58-74
	int d = INT_MIN, max_j_end = INT_MIN;
	for (const Segment& s : segs) {
		const int d2 = s.diag();
		if (d2 != d) { d = d2; ch.g.nodes.emplace_back(s); max_j_end = s.subject_end(); }
		else if (max_j_end < s.j) { ch.g.nodes.emplace_back(s); max_j_end = std::max(max_j_end, s.subject_end()); }
	}
	ch.run(out, SPACE_PENALTY, 19, 2000);
	// merge_hsps, greedy_align.cpp:461-480 (std::list erase semantics on a vector)
	for (size_t a = 0; a < out.size(); ++a) {
		size_t b = a + 1;
		while (b < out.size()) {
			if (merge_score(out[a], out[b]) > std::max(out[a].score, out[b].score)) { out[a] = merge(out[a], out[b]); out.erase(out.begin() + (ptrdiff_t)b); }
			else if (merge_score(out[b], out[a]) > std::max(out[a].score, out[b].score)) { out[a] = merge(out[b], out[a]); out.erase(out.begin() + (ptrdiff_t)b); }
			else ++b;
		}
	}
}

}  // namespace dmnd

// C wrapper of the per-(query, target) bridge -- hits -> segments -> chains -> merged bands (pipeline.cpp: produce_round1) -- for the
// TEST-ONLY oracle library's dmnd_hits_chain and for tests: the device code (cuda/chain_kernels.cuh) must return the same bands.
extern "C" int dmnd_host_chain_pair(const int32_t* hit_i, const int32_t* hit_j, const dmnd_segment* hit_seg, int nh, const int8_t* query, int qlen,
                                    const int8_t* subject, int slen, int band, int32_t* d0_out, int32_t* d1_out, int cap) {
	using namespace dmnd;
	static const Scoring sc;
	struct H { int i, j; dmnd_segment s; };
	std::vector<H> hh((size_t)nh);
	for (int k = 0; k < nh; ++k) hh[(size_t)k] = H{ hit_i[k], hit_j[k], hit_seg[k] };
	std::sort(hh.begin(), hh.end(), [](const H& x, const H& y) { const int a = x.i - x.j, b = y.i - y.j; return a < b || (a == b && x.j < y.j); });
	std::vector<Segment> segs;
	for (const H& h : hh) {  // align/ungapped.cpp:81-91
		if (!segs.empty() && segs.back().diag() == h.i - h.j && segs.back().subject_end() >= h.j) continue;
		if (h.s.score > 0) segs.push_back(Segment{ h.s.i, h.s.j, h.s.len, h.s.score });
	}
	if (segs.empty()) return 0;
	std::stable_sort(segs.begin(), segs.end(), [](const Segment& x, const Segment& y) { return x.diag() < y.diag() || (x.diag() == y.diag() && x.j < y.j); });
	std::vector<Chain> chains;
	chain_segments(sc, query, qlen, subject, slen, segs, chains);
	std::stable_sort(chains.begin(), chains.end(), [](const Chain& x, const Chain& y) { return x.d_min < y.d_min; });
	int np = 0, d0 = INT_MAX, d1 = INT_MIN;
	auto emit = [&]() -> bool { if (np >= cap) return false; d0_out[np] = d0; d1_out[np] = d1; ++np; return true; };
	for (const Chain& h : chains) {  // add_dp_targets, align/gapped_score.cpp:107-180
		const int b0 = std::max(h.d_min - band, -(slen - 1)), b1 = std::min(h.d_max + 1 + band, qlen);
		bool merge = false;
		if (d0 != INT_MAX) {
			const int ib = std::max(d0, b0), ie = std::min(d1, b1);
			const double overlap = ie > ib ? ie - ib : 0;
			merge = overlap / (d1 - d0) > 0.0 || overlap / (b1 - b0) > 0.0;
		}
		if (merge) { d0 = std::min(d0, b0); d1 = std::max(d1, b1); }
		else {
			if (d0 != INT_MAX && !emit()) return -1;
			d0 = b0; d1 = b1;
		}
	}
	if (!chains.empty() && !emit()) return -1;
	return np;
}
