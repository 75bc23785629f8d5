// chain_kernels.cuh -- the host bridge between the seed stage and the DP kernels, on the device (sm_100a): one thread per
// (query, target) pair turns the pair's seed hits into banded DP problems:
//   align/load_hits.h:44-122                hits of one target, ordered by (diagonal, subject position)
//   align/ungapped.cpp:62-118               x-drop segments (already evaluated per hit by xdrop_kernel), covered hits skipped
//   chaining/greedy_align.cpp:58-127        DiagGraph::load / sort / prune
//   chaining/greedy_align.cpp:150-273       get_hgap_link / get_vgap_link / get_approximate_link
//   chaining/greedy_align.cpp:275-366       Aligner::forward_pass (std::map window, erase while iterating)
//   chaining/diag_graph.h:123-169           add_edge / get_edge / prefix_score
//   chaining/backtrace.cpp:36-173,269-356   disjoint / backtrace_old / top-node loop
//   chaining/greedy_align.cpp:426-497       merge_hsps / Chaining::run
//   align/gapped_score.cpp:41-72,107-180    Extension::band, add_dp_targets (band merge)
// Device code only (tests/emu_chain.cpp compiles THIS file for the CPU and checks it against host/chaining.cpp, the statement-
// level restatement the CPU pipeline uses).  Containers are fixed-capacity arrays in local memory; a pair that does not fit
// (more than CH_HITS hits, CH_NODES graph nodes, ...) sets an overflow flag and its QUERY is processed by the host path, so
// the capacities bound memory, not correctness.  Every arithmetic step is the reference's: int truncations of double
// products are written with explicit round-to-nearest intrinsics so that no FMA contraction can change a product.
#pragma once
#include "dev_params.h"

namespace dmnd_cuda {

constexpr int CH_HITS = 64, CH_NODES = 32, CH_EDGES = 160, CH_TOP = 16, CH_CHAINS = 8, CH_PROBS = 4;
#ifndef CH_DMUL
#define CH_DMUL(a, b) __dmul_rn((a), (b))
#define CH_DSUB(a, b) __dsub_rn((a), (b))
#define CH_DDIV(a, b) __ddiv_rn((a), (b))
#endif

struct ChSeg {
	int i, j, len, score;
	__device__ __forceinline__ int diag() const { return i - j; }
	__device__ __forceinline__ int subject_last() const { return j + len - 1; }
	__device__ __forceinline__ int query_last() const { return i + len - 1; }
	__device__ __forceinline__ int subject_end() const { return j + len; }
	__device__ __forceinline__ int query_end() const { return i + len; }
};
struct ChNode : ChSeg {
	int link_idx, prefix_score, path_max, path_min;
	__device__ __forceinline__ int rel_score() const { return prefix_score == path_max ? prefix_score : prefix_score - path_min; }
};
struct ChEdge { int prefix_score, path_max, j, path_min, prefix_score_begin, node_in, node_out; };
struct ChLink { int subject_pos1, query_pos1, subject_pos2, query_pos2, score1, score2; };
struct ChChain { int d_min, d_max, score, q_begin, q_end, s_begin, s_end; };

struct Chainer {
	const int8_t* score;  // 32 x 32 matrix
	const int8_t *query, *subject;  // letter 0 of both sequences
	int qlen, slen, gap_open, gap_extend;
	ChNode nodes[CH_NODES]; int nn;
	ChEdge edges[CH_EDGES]; int ne;
	int win_key[CH_NODES], win_val[CH_NODES], nw;
	bool overflow;

	__device__ __forceinline__ int sc(int a, int b) const { return (int)score[(a << 5) | b]; }
	__device__ __forceinline__ int ql(int i) const { return query[i] & 31; }
	__device__ __forceinline__ int sl(int j) const { return subject[j] & 31; }

	// ---- diag_graph.h:123-169
	__device__ void add_edge(const ChEdge& e) {
		if (ne >= CH_EDGES) { overflow = true; return; }
		for (int k = e.node_in + 1; k < nn; ++k) {
			if (nodes[k].link_idx == -1) break;
			++nodes[k].link_idx;
		}
		ChNode& d = nodes[e.node_in];
		if (e.prefix_score > d.prefix_score) { d.prefix_score = e.prefix_score; d.path_max = e.path_max; d.path_min = e.path_min; }
		const int at = d.link_idx++;
		for (int k = ne; k > at; --k) edges[k] = edges[k - 1];
		edges[at] = e;
		++ne;
	}
	// returns the edge index or -1; (size_t)(link_idx - 1) of a zero-score node can be "edge -1" = none as well
	__device__ int get_edge(int node, int j) const {
		const ChNode& d = nodes[node];
		if (d.score == 0) { const int k = d.link_idx - 1; return (k >= 0 && k < ne) ? k : -1; }
		if (ne == 0) return -1;
		int max_score = d.score, max_i = -1;
		for (int i = d.link_idx - 1; i >= 0 && edges[i].node_in == node; --i)
			if (edges[i].j < j && edges[i].prefix_score > max_score) { max_i = i; max_score = edges[i].prefix_score; }
		return max_i;
	}
	__device__ int prefix_score(int node, int j, int& path_max, int& path_min) const {
		const int i = get_edge(node, j);
		const bool none = i < 0;
		path_max = none ? nodes[node].score : max(nodes[node].score, edges[i].path_max);
		path_min = none ? nodes[node].score : edges[i].path_min;
		return none ? nodes[node].score : max(nodes[node].score, edges[i].prefix_score);
	}

	// ---- greedy_align.cpp:150-214; TR = the transposed call (roles of query and subject swapped)
	template<bool TR> __device__ __forceinline__ int pair_score(int i, int j) const { return TR ? sc(sl(i), ql(j)) : sc(ql(i), sl(j)); }  // transposed: the roles swap, the matrix index order with them
	template<bool TR> __device__ int score_range(int i, int j, int j_end) const {
		int v = 0;
		while (j < j_end) { v += pair_score<TR>(i, j); ++i; ++j; }
		return v;
	}
	template<bool TR> __device__ int hgap_link(const ChSeg& d1, const ChSeg& d2, ChLink& l, int padding) const {
		const int d = d1.diag() - d2.diag(),
			j2_end = min(max(d2.j, d1.subject_last() + d + 1 + padding), d2.subject_last());
		int j1;
		bool space;
		if (d1.subject_last() < d2.j - d - 1) { j1 = d1.subject_last(); space = true; }
		else { j1 = max(d2.j - d - 1 - padding, d1.j); space = false; }
		int j2 = j1 + d + 1, i1 = d1.i + (j1 - d1.j), i2 = i1 + 1;
		if (j2 > d2.subject_last()) { l.subject_pos1 = -1; l.score1 = 0; l.score2 = 0; return INT_MIN; }
		int score1 = 0, score2 = score_range<TR>(i2, j2, d2.j) + d2.score - score_range<TR>(d2.i, d2.j, j2);
		int max_score = INT_MIN;
		for (;;) {
			if (score1 + score2 > max_score) {
				max_score = score1 + score2;
				l.query_pos1 = i1; l.subject_pos1 = j1; l.query_pos2 = i2; l.subject_pos2 = j2; l.score1 = score1; l.score2 = score2;
			}
			score2 -= pair_score<TR>(i2, j2);
			++i1; ++i2; ++j1; ++j2;
			if (j2 > j2_end) break;
			score1 += pair_score<TR>(i1, j1);
		}
		const int j1_end = j2_end - d;
		if (space) l.score1 += d1.score;
		else l.score1 += d1.score - score_range<TR>(d1.diag() + j1_end, j1_end, d1.subject_end())
			+ score_range<TR>(d1.query_end(), d1.subject_end(), j1_end) - score1;
		return max_score;
	}
	__device__ int get_link(const ChSeg& d1, const ChSeg& d2, ChLink& l, int padding) const {
		l.subject_pos1 = -1; l.query_pos1 = 0; l.subject_pos2 = 0; l.query_pos2 = 0; l.score1 = 0; l.score2 = 0;
		if (d1.diag() < d2.diag()) {
			const ChSeg t1{ d1.j, d1.i, d1.len, d1.score }, t2{ d2.j, d2.i, d2.len, d2.score };
			const int s = hgap_link<true>(t1, t2, l, padding);
			int t = l.subject_pos1; l.subject_pos1 = l.query_pos1; l.query_pos1 = t;
			t = l.subject_pos2; l.subject_pos2 = l.query_pos2; l.query_pos2 = t;
			return s;
		}
		return hgap_link<false>(d1, d2, l, padding);
	}

	// ---- greedy_align.cpp:220-273
	__device__ int approximate_link(int d_idx, int e_idx, double space_penalty) {
		const ChNode d = nodes[d_idx], e = nodes[e_idx];
		const int shift = d.diag() - e.diag();
		const int gap_score = shift != 0 ? -gap_open - abs(shift) * gap_extend : 0;
		const int space = shift > 0 ? d.j - e.subject_last() : d.i - e.query_last();
		int prefix = 0, link_j = 0, diff1 = 0, path_max = 0, path_min = 0, prefix_begin = 0;
		if (space <= 0 || space_penalty == 0.0) {
			const int edge = get_edge(d_idx, d.j);
			if (edge >= 0 && edges[edge].prefix_score > e.prefix_score + gap_score + d.score) return 0;
			ChLink link;
			if (get_link(e, d, link, 10) > 0) {
				diff1 = e.score - link.score1;
				const int prefix_e = prefix_score(e_idx, link.subject_pos1, path_max, path_min);
				prefix = prefix_e - diff1 + gap_score + link.score2;
				const int edge2 = get_edge(d_idx, link.subject_pos2);
				if (edge2 >= 0 && edges[edge2].prefix_score > prefix) return 0;
				prefix_begin = prefix - link.score2;
				path_min = min(path_min, prefix - link.score2);
				if (prefix_e == path_max) path_max -= diff1;
				link_j = link.subject_pos2;
			}
		}
		else {
			prefix = e.prefix_score + gap_score - (int)CH_DMUL(space_penalty, (double)max(space - 1, 0)) + d.score;
			const int edge = get_edge(d_idx, d.j);
			if (edge >= 0 && edges[edge].prefix_score > prefix) return 0;
			prefix_begin = prefix - d.score;
			path_max = e.path_max;
			path_min = e.path_min;
			path_min = min(path_min, prefix - d.score);
			link_j = d.j;
		}
		if (prefix > d.score) {
			path_max = max(path_max, prefix);
			add_edge(ChEdge{ prefix, path_max, link_j, prefix == path_max ? prefix : path_min, prefix_begin, d_idx, e_idx });
		}
		return prefix;
	}

	// ---- the ordered diagonal -> node window of forward_pass (a std::map<int, unsigned> in the reference)
	__device__ int win_find_or_insert(int key, int val) {
		int lo = 0, hi = nw;
		while (lo < hi) { const int mid = (lo + hi) / 2; if (win_key[mid] < key) lo = mid + 1; else hi = mid; }
		if (lo == nw || win_key[lo] != key) {
			for (int k = nw; k > lo; --k) { win_key[k] = win_key[k - 1]; win_val[k] = win_val[k - 1]; }
			win_key[lo] = key; win_val[lo] = val; ++nw;
		}
		return lo;
	}
	__device__ void win_erase(int i) {
		for (int k = i; k + 1 < nw; ++k) { win_key[k] = win_key[k + 1]; win_val[k] = win_val[k + 1]; }
		--nw;
	}
	__device__ bool expired(const ChNode& d, const ChNode& e, double space_penalty) const {
		return e.prefix_score - (int)CH_DMUL(space_penalty, (double)max(d.j - e.subject_end(), 0)) <= 0;
	}
	// greedy_align.cpp:275-366
	__device__ void forward_pass(double space_penalty) {
		nw = 0;
		for (int node = 0; node < nn && !overflow; ++node) {
			nodes[node].link_idx = ne;
			const int dd = nodes[node].diag();
			int i = win_find_or_insert(dd, node);
			int j = i, max_j = 0;
			if (i != 0) {
				do {
					--j;
					if (expired(nodes[node], nodes[win_val[j]], space_penalty)) {
						const bool was_first = j == 0;
						win_erase(j);
						--i;
						if (was_first) break;
						continue;
					}
					if (nodes[win_val[j]].subject_end() < max_j) continue;
					const int e_idx = win_val[j];
					approximate_link(node, e_idx, space_penalty);
					const ChNode& d2 = nodes[node];
					const ChNode& e2 = nodes[e_idx];
					max_j = max(max_j, min(d2.j, e2.subject_end()));
					if (e2.subject_end() - (d2.subject_end() - min(e2.diag() - d2.diag(), 0)) >= 10) approximate_link(e_idx, node, space_penalty);
				} while (j != 0);
			}
			j = i;
			if (win_val[j] == node) ++j;
			int max_i = 0;
			while (j != nw) {
				if (expired(nodes[node], nodes[win_val[j]], space_penalty) && j != i) { win_erase(j); continue; }
				if (nodes[win_val[j]].query_end() < max_i) { ++j; continue; }
				const int e_idx = win_val[j];
				approximate_link(node, e_idx, space_penalty);
				const ChNode& d2 = nodes[node];
				const ChNode& e2 = nodes[e_idx];
				if (e2.i < d2.i) max_i = max(max_i, min(e2.query_end(), d2.i));
				if (e2.subject_end() - (d2.subject_end() - min(e2.diag() - d2.diag(), 0)) >= 10) approximate_link(e_idx, node, space_penalty);
				++j;
			}
			win_val[i] = node;
		}
	}

	// ---- greedy_align.cpp:106-127 (chaining_range_cover = 8); the node list is rebuilt in place through `tmp`
	__device__ void prune(ChNode* fin, ChNode* win) {
		int nf = 0, nwin = 0;
		for (int x = 0; x < nn; ++x) {
			const ChNode d = nodes[x];
			int n = 0;
			for (int k = 0; k < nwin;) {
				if (win[k].subject_end() > d.j) {
					if (win[k].score >= d.score && win[k].j <= d.j && win[k].subject_end() >= d.subject_end()) ++n;
					++k;
				}
				else {
					fin[nf++] = win[k];
					for (int y = k; y + 1 < nwin; ++y) win[y] = win[y + 1];
					--nwin;
				}
			}
			if (n <= 8) win[nwin++] = d;
		}
		for (int k = 0; k < nwin; ++k) fin[nf++] = win[k];
		for (int k = 0; k < nf; ++k) nodes[k] = fin[k];
		nn = nf;
	}

	__device__ static double overlap_factor(int b0, int e0, int b1, int e1) {
		const int ib = max(b0, b1), ie = min(e0, e1);
		const unsigned ov = (unsigned)(ie > ib ? ie - ib : 0);
		const int len = e0 > b0 ? e0 - b0 : 0;
		return CH_DDIV((double)ov, (double)len);
	}
	// backtrace.cpp:36-76
	__device__ static bool disjoint(const ChChain* ts, int nts, int qb, int qe, int sb, int se, int score, int cutoff) {
		for (int k = 0; k < nts; ++k) {
			const double ot = overlap_factor(sb, se, ts[k].s_begin, ts[k].s_end), oq = overlap_factor(qb, qe, ts[k].q_begin, ts[k].q_end);
			const double lo = oq < ot ? oq : ot, hi = ot < oq ? oq : ot;  // std::min / std::max (their NaN behaviour, not fmin / fmax's)
			if (CH_DDIV(CH_DMUL(CH_DSUB(1.0, lo), (double)score), (double)ts[k].score) >= 0.5) continue;
			if (CH_DMUL(CH_DSUB(1.0, hi), (double)score) < (double)cutoff) return false;
		}
		return true;
	}
	// backtrace.cpp:78-173 (out == nullptr), the recursion unrolled into a descent and an unwinding loop
	__device__ bool backtrace(int node, int j_end, ChChain& t, int score_max, int score_min, int max_shift, int& next) {
		int st_node[CH_NODES + 1], st_edge[CH_NODES + 1], st_min[CH_NODES + 1];
		int sp = 0, cur = node, cur_j = j_end, cur_min = score_min;
		bool ret;
		for (;;) {
			const ChNode& d = nodes[cur];
			const int f = get_edge(cur, cur_j);
			bool at_end = f < 0;
			const int prefix = at_end ? d.score : edges[f].prefix_score;
			if (prefix > score_max) { ret = false; break; }
			cur_min = min(cur_min, at_end ? 0 : edges[f].prefix_score_begin);
			if (!at_end) {
				const ChEdge& ed = edges[f];
				const int shift = d.diag() - nodes[ed.node_out].diag();
				if (abs(shift) <= max_shift) {
					if (sp >= CH_NODES) { overflow = true; ret = false; break; }
					st_node[sp] = cur; st_edge[sp] = f; st_min[sp] = cur_min; ++sp;
					cur_j = shift > 0 ? ed.j : ed.j + shift;
					cur = ed.node_out;
					continue;
				}
				next = ed.node_out;
				at_end = true;
			}
			t.q_begin = d.i; t.s_begin = d.j; t.score = score_max - cur_min;
			t.d_max = max(t.d_max, d.diag()); t.d_min = min(t.d_min, d.diag());
			ret = true;
			break;
		}
		while (sp > 0) {
			--sp;
			const ChNode& d = nodes[st_node[sp]];
			bool at_end = false;
			if (!ret) {
				if (edges[st_edge[sp]].prefix_score_begin > st_min[sp]) continue;  // this frame returns false as well
				at_end = true;
			}
			if (at_end) { t.q_begin = d.i; t.s_begin = d.j; t.score = score_max - st_min[sp]; }
			t.d_max = max(t.d_max, d.diag()); t.d_min = min(t.d_min, d.diag());
			ret = true;
		}
		return ret;
	}

	// ---- Chaining::run for >= 2 segments (greedy_align.cpp:368-394,482-497; backtrace.cpp:327-356); chains into ts, count returned
	__device__ int run(const ChSeg* segs, int nsegs, ChChain* ts, ChNode* tmp1, ChNode* tmp2) {
		const double space_penalty = 0.1;
		const int cutoff = 19, max_shift = 2000;
		nn = 0; ne = 0;
		// DiagGraph::load, greedy_align.cpp:58-74
		int dprev = INT_MIN, max_j_end = INT_MIN;
		for (int k = 0; k < nsegs; ++k) {
			const ChSeg& s = segs[k];
			const int d2 = s.diag();
			bool take = false;
			if (d2 != dprev) { dprev = d2; take = true; max_j_end = s.subject_end(); }
			else if (max_j_end < s.j) { take = true; max_j_end = max(max_j_end, s.subject_end()); }
			if (take) {
				if (nn >= CH_NODES) { overflow = true; return 0; }
				ChNode& n = nodes[nn++];
				n.i = s.i; n.j = s.j; n.len = s.len; n.score = s.score; n.link_idx = -1; n.prefix_score = s.score; n.path_max = s.score; n.path_min = s.score;
			}
		}
		// (chaining_min_nodes = 200 > CH_NODES: the length-cap branch of greedy_align.cpp:372-386 never applies here)
		for (int a = 1; a < nn; ++a) {  // sort by (j, i): keys are unique, so any sort gives the reference's order
			const ChNode x = nodes[a];
			int b = a;
			while (b > 0 && (x.j < nodes[b - 1].j || (x.j == nodes[b - 1].j && x.i < nodes[b - 1].i))) { nodes[b] = nodes[b - 1]; --b; }
			nodes[b] = x;
		}
		prune(tmp1, tmp2);
		forward_pass(space_penalty);
		if (overflow) return 0;
		int top[CH_TOP], ntop = 0;
		for (int k = 0; k < nn; ++k)
			if (nodes[k].rel_score() >= cutoff) {
				if (ntop >= CH_TOP) { overflow = true; return 0; }  // (std::sort is an insertion sort up to 16 elements: stable, reproduced below)
				top[ntop++] = k;
			}
		for (int a = 1; a < ntop; ++a) {
			const int x = top[a];
			int b = a;
			while (b > 0 && nodes[x].rel_score() > nodes[top[b - 1]].rel_score()) { top[b] = top[b - 1]; --b; }
			top[b] = x;
		}
		int nts = 0;
		for (int a = 0; a < ntop; ++a) {
			const ChNode& n = nodes[top[a]];
			if (!disjoint(ts, nts, n.i, n.i + n.len, n.j, n.j + n.len, n.score, cutoff)) continue;
			int top_node = top[a], next, max_j = slen;
			do {
				ChChain t{ INT_MAX, INT_MIN, 0, 0, 0, 0, 0 };
				next = -1;
				const ChNode& d = nodes[top_node];
				t.s_end = d.subject_end();
				t.q_end = d.query_end();
				backtrace(top_node, min(d.subject_end(), max_j), t, d.prefix_score, d.prefix_score, max_shift, next);
				if (overflow) return 0;
				if (t.score > 0) max_j = t.s_begin;
				if (t.score >= cutoff && disjoint(ts, nts, t.q_begin, t.q_end, t.s_begin, t.s_end, t.score, cutoff)) {
					if (nts >= CH_CHAINS) { overflow = true; return 0; }
					ts[nts++] = t;
				}
				top_node = next;
			} while (next != -1);
		}
		return nts;
	}
};

// greedy_align.cpp:426-437
__device__ __forceinline__ int ch_merge_score(const ChChain& h1, const ChChain& h2) {
	const int gq = h2.q_begin - h1.q_end, gt = h2.s_begin - h1.s_end;
	if (gq < 0 || gt < 0) return 0;
	const double s = (double)(h1.score + h2.score);
	if (gq > gt) return (int)CH_DSUB(CH_DSUB(s, CH_DMUL((double)gq, 0.5)), CH_DMUL((double)gt, 0.1));
	return (int)CH_DSUB(CH_DSUB(s, CH_DMUL((double)gt, 0.5)), CH_DMUL((double)gq, 0.1));
}
__device__ __forceinline__ ChChain ch_merge(const ChChain& h1, const ChChain& h2) {
	ChChain h;
	h.d_max = max(h1.d_max, h2.d_max); h.d_min = min(h1.d_min, h2.d_min);
	h.q_begin = h1.q_begin; h.q_end = h2.q_end; h.s_begin = h1.s_begin; h.s_end = h2.s_end;
	h.score = ch_merge_score(h1, h2);
	return h;
}

__device__ __forceinline__ int ch_band_for(int len, bool slow) {  // Extension::band, align/gapped_score.cpp:41-72
	if (!slow) return len < 50 ? 12 : len < 100 ? 16 : len < 250 ? 30 : len < 350 ? 40 : 64;
	return len < 50 ? 15 : len < 100 ? 20 : len < 150 ? 30 : len < 200 ? 50 : len < 250 ? 60 : len < 350 ? 100 : len < 500 ? 120 : 150;
}

// One (query, target) pair: its hits (query offset i, subject offset j, x-drop segment) -> banded DP problems.
// Returns the number of problems written to out_d0 / out_d1 (<= CH_PROBS), or -1 when a capacity was exceeded.
struct ChHit { int i, j; ChSeg seg; };
__device__ int chain_pair(Chainer& C, ChHit* hits, int nh, int band, int* out_d0, int* out_d1, ChSeg* segs, ChNode* tmp1, ChNode* tmp2, ChChain* chains) {
	// load_hits order inside a target is irrelevant: the hits are sorted by (diagonal, j) (SeedHit::operator<, align/target.h)
	for (int a = 1; a < nh; ++a) {
		const ChHit x = hits[a];
		const int dx = x.i - x.j;
		int b = a;
		while (b > 0 && (dx < hits[b - 1].i - hits[b - 1].j || (dx == hits[b - 1].i - hits[b - 1].j && x.j < hits[b - 1].j))) { hits[b] = hits[b - 1]; --b; }
		hits[b] = x;
	}
	// align/ungapped.cpp:81-91
	int ns = 0;
	for (int k = 0; k < nh; ++k) {
		const ChHit& h = hits[k];
		if (ns > 0 && segs[ns - 1].diag() == h.i - h.j && segs[ns - 1].subject_end() >= h.j) continue;
		if (h.seg.score > 0) segs[ns++] = h.seg;
	}
	if (ns == 0) return 0;
	for (int a = 1; a < ns; ++a) {  // stable sort by (diag, j)
		const ChSeg x = segs[a];
		int b = a;
		while (b > 0 && (x.diag() < segs[b - 1].diag() || (x.diag() == segs[b - 1].diag() && x.j < segs[b - 1].j))) { segs[b] = segs[b - 1]; --b; }
		segs[b] = x;
	}
	int nc = 0;
	if (ns == 1) {  // greedy_align.cpp:485-489
		const ChSeg& s = segs[0];
		chains[0] = ChChain{ s.diag(), s.diag(), s.score, s.i, s.i + s.len, s.j, s.j + s.len };
		nc = 1;
	}
	else {
		C.overflow = false;
		nc = C.run(segs, ns, chains, tmp1, tmp2);
		if (C.overflow) return -1;
		// merge_hsps, greedy_align.cpp:461-480
		for (int a = 0; a < nc; ++a) {
			int b = a + 1;
			while (b < nc) {
				bool merged = false;
				if (ch_merge_score(chains[a], chains[b]) > max(chains[a].score, chains[b].score)) { chains[a] = ch_merge(chains[a], chains[b]); merged = true; }
				else if (ch_merge_score(chains[b], chains[a]) > max(chains[a].score, chains[b].score)) { chains[a] = ch_merge(chains[b], chains[a]); merged = true; }
				if (merged) { for (int y = b; y + 1 < nc; ++y) chains[y] = chains[y + 1]; --nc; }
				else ++b;
			}
		}
		for (int a = 1; a < nc; ++a) {  // stable sort by d_min
			const ChChain x = chains[a];
			int b = a;
			while (b > 0 && x.d_min < chains[b - 1].d_min) { chains[b] = chains[b - 1]; --b; }
			chains[b] = x;
		}
	}
	// add_dp_targets, align/gapped_score.cpp:107-180 (min_band_overlap 0: any overlap merges)
	int np = 0, d0 = INT_MAX, d1 = INT_MIN;
	for (int k = 0; k < nc; ++k) {
		const int b0 = max(chains[k].d_min - band, -(C.slen - 1)), b1 = min(chains[k].d_max + 1 + band, C.qlen);
		bool merge = false;
		if (d0 != INT_MAX) {
			const int ib = max(d0, b0), ie = min(d1, b1);
			merge = ie > ib;  // overlap / (d1 - d0) > 0.0 || overlap / (b1 - b0) > 0.0 with positive lengths
		}
		if (merge) { d0 = min(d0, b0); d1 = max(d1, b1); }
		else {
			if (d0 != INT_MAX) { if (np >= CH_PROBS) return -1; out_d0[np] = d0; out_d1[np] = d1; ++np; }
			d0 = b0; d1 = b1;
		}
	}
	if (nc > 0) { if (np >= CH_PROBS) return -1; out_d0[np] = d0; out_d1[np] = d1; ++np; }
	return np;
}

}  // namespace dmnd_cuda
