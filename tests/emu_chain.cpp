// emu_chain.cpp -- the device host-bridge code (diamond_b200/csrc/cuda/chain_kernels.cuh: hits of one (query, target) pair ->
// segments -> greedy chaining -> DP bands) compiled for the CPU and compared with the host restatement the CPU pipeline uses
// (diamond_b200/csrc/host/chaining.cpp + the band merge of pipeline.cpp: produce_round1) on random related sequence pairs:
// substitutions, indels, internal repeats (many nodes on many diagonals), several hits per diagonal.
// usage: emu_chain SEED PAIRS
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <random>
#include <vector>
#define __device__
#define __forceinline__ inline
#define CH_DMUL(a, b) ((a) * (b))
#define CH_DSUB(a, b) ((a) - (b))
#define CH_DDIV(a, b) ((a) / (b))
using std::max; using std::min;
#include "../diamond_b200/csrc/cuda/chain_kernels.cuh"
#include "../diamond_b200/csrc/host/chaining.h"
using namespace dmnd_cuda;

static int band_for(int len) { return len < 50 ? 12 : len < 100 ? 16 : len < 250 ? 30 : len < 350 ? 40 : 64; }

int main(int argc, char** argv) {
	const int seed = argc > 1 ? atoi(argv[1]) : 1, pairs = argc > 2 ? atoi(argv[2]) : 2000;
	std::mt19937 rng((unsigned)seed);
	dmnd::Scoring sc;
	int8_t m8[1024];
	for (int k = 0; k < 1024; ++k) m8[k] = (int8_t)sc.m32[k];
	int fails = 0, overflows = 0, multi = 0, probs = 0;
	for (int it = 0; it < pairs; ++it) {
		const int slen = 30 + (int)(rng() % 600);
		std::vector<int8_t> sb(64, 31), qb(64, 31);
		std::vector<int8_t> s((size_t)slen);
		for (auto& x : s) x = (int8_t)(rng() % 20);
		if (rng() % 3 == 0) {  // internal repeat in the subject: hits on several diagonals
			const int rl = 10 + (int)(rng() % 40), a = (int)(rng() % (unsigned)std::max(1, slen - 2 * rl)), b = a + rl + (int)(rng() % (unsigned)std::max(1, slen - a - 2 * rl));
			for (int k = 0; k < rl && b + k < slen; ++k) s[(size_t)(b + k)] = s[(size_t)(a + k)];
		}
		// query: a mutated window of the subject with indels
		std::vector<int8_t> q;
		const int w0 = (int)(rng() % (unsigned)std::max(1, slen - 20)), wl = std::min(slen - w0, 20 + (int)(rng() % 400));
		const double sub = 0.05 + 0.4 * (rng() % 1000) / 1000.0;
		for (int k = 0; k < wl; ++k) {
			if (rng() % 100 < 2) continue;  // deletion
			q.push_back((rng() % 1000) / 1000.0 < sub ? (int8_t)(rng() % 20) : s[(size_t)(w0 + k)]);
			if (rng() % 100 < 2) { const int n = 1 + (int)(rng() % 6); for (int x = 0; x < n; ++x) q.push_back((int8_t)(rng() % 20)); }
		}
		if (q.size() < 12) continue;
		const int qlen = (int)q.size();
		sb.insert(sb.end(), s.begin(), s.end()); sb.resize(sb.size() + 64, 31);
		qb.insert(qb.end(), q.begin(), q.end()); qb.resize(qb.size() + 64, 31);
		const int8_t *Q = qb.data() + 64, *S = sb.data() + 64;
		// hits: positions where 6 consecutive letters agree (several per diagonal), plus a few random ones
		std::vector<ChHit> hits;
		for (int i = 0; i + 6 <= qlen && hits.size() < 200; ++i)
			for (int j = 0; j + 6 <= slen && hits.size() < 200; ++j)
				if (memcmp(Q + i, S + j, 6) == 0 && rng() % 3 == 0) hits.push_back(ChHit{ i, j, ChSeg{ 0, 0, 0, 0 } });
		for (int x = 0; x < 2; ++x) hits.push_back(ChHit{ (int)(rng() % (unsigned)qlen), (int)(rng() % (unsigned)slen), ChSeg{ 0, 0, 0, 0 } });
		if ((int)hits.size() > CH_HITS) hits.resize(CH_HITS);
		for (auto& h : hits) { const dmnd::Segment g = dmnd::xdrop_ungapped(sc, Q, nullptr, S, h.i, h.j); h.seg = ChSeg{ g.i, g.j, g.len, g.score }; }
		std::shuffle(hits.begin(), hits.end(), rng);
		const int band = band_for(qlen);
		// ---- host path (pipeline.cpp: produce_round1)
		std::vector<ChHit> hh = hits;
		std::sort(hh.begin(), hh.end(), [](const ChHit& x, const ChHit& y) { const int d1 = x.i - x.j, d2 = y.i - y.j; return d1 < d2 || (d1 == d2 && x.j < y.j); });
		std::vector<dmnd::Segment> segs;
		for (const ChHit& h : hh) {
			if (!segs.empty() && segs.back().diag() == h.i - h.j && segs.back().subject_end() >= h.j) continue;
			if (h.seg.score > 0) segs.push_back(dmnd::Segment{ h.seg.i, h.seg.j, h.seg.len, h.seg.score });
		}
		std::vector<std::pair<int, int>> want;
		if (!segs.empty()) {
			std::stable_sort(segs.begin(), segs.end(), [](const dmnd::Segment& x, const dmnd::Segment& y) { return x.diag() < y.diag() || (x.diag() == y.diag() && x.j < y.j); });
			std::vector<dmnd::Chain> chains;
			dmnd::chain_segments(sc, Q, qlen, S, slen, segs, chains);
			std::stable_sort(chains.begin(), chains.end(), [](const dmnd::Chain& x, const dmnd::Chain& y) { return x.d_min < y.d_min; });
			if (segs.size() > 1) ++multi;
			int d0 = INT_MAX, d1 = INT_MIN;
			for (const dmnd::Chain& h : chains) {
				const int b0 = std::max(h.d_min - band, -(slen - 1)), b1 = std::min(h.d_max + 1 + band, qlen);
				bool merge = false;
				if (d0 != INT_MAX) { const int ib = std::max(d0, b0), ie = std::min(d1, b1); const double overlap = ie > ib ? ie - ib : 0; merge = overlap / (d1 - d0) > 0.0 || overlap / (b1 - b0) > 0.0; }
				if (merge) { d0 = std::min(d0, b0); d1 = std::max(d1, b1); }
				else { if (d0 != INT_MAX) want.push_back({ d0, d1 }); d0 = b0; d1 = b1; }
			}
			if (!chains.empty()) want.push_back({ d0, d1 });
		}
		// ---- device code
		static Chainer C;
		C.score = m8; C.query = Q; C.subject = S; C.qlen = qlen; C.slen = slen; C.gap_open = sc.gap_open; C.gap_extend = sc.gap_extend;
		static ChSeg dsegs[CH_HITS]; static ChNode t1[CH_NODES], t2[CH_NODES]; static ChChain ch[CH_CHAINS];
		int o0[CH_PROBS], o1[CH_PROBS];
		std::vector<ChHit> dh = hits;
		const int np = chain_pair(C, dh.data(), (int)dh.size(), band, o0, o1, dsegs, t1, t2, ch);
		if (np < 0) { ++overflows; continue; }
		probs += np;
		bool same = np == (int)want.size();
		for (int k = 0; same && k < np; ++k) same = o0[k] == want[(size_t)k].first && o1[k] == want[(size_t)k].second;
		if (!same) {
			++fails;
			if (fails < 6) {
				printf("MISMATCH pair %d qlen %d slen %d hits %zu: host", it, qlen, slen, hits.size());
				for (auto& w : want) printf(" [%d,%d)", w.first, w.second);
				printf(" | device");
				for (int k = 0; k < np; ++k) printf(" [%d,%d)", o0[k], o1[k]);
				printf("\n");
			}
		}
	}
	printf("pairs=%d multi_segment=%d problems=%d overflows=%d fails=%d \n", pairs, multi, probs, overflows, fails);
	return fails ? 1 : 0;
}
