// pipeline.cpp -- P layer: one (query block, reference block) pass of blastp on top of the K layer.
//
// Host restatement of the reference's control flow around its two hot kernels, batched across queries so that every
// GPU launch sees the DP problems of ALL queries at once (the reference batches only within one query):
//   search/setup.cpp:306-309,338-402       seedp_bits, setup_search (sensitivity traits -> parameters)
//   run/double_indexed.cpp:102-252         run_ref_chunk: search_shape per shape, then align_queries
//   align/align.cpp:203-269                align_queries (sort by query, per-query extend)
//   align/load_hits.h:44-122               load_hits
//   align/extend.cpp:79-119,226-387        ranking_chunk_size, ranking_terminate, extend (chunk loop)
//   align/ungapped.cpp:62-118              ungapped_stage
//   align/gapped_score.cpp:41-72,107-246   band, add_dp_targets, round-1 align
//   align/gapped_final.cpp:64-158          round-2 align
//   align/culling.cpp:34-202               inner_culling, culling, append_hits, output_range
//   dp/dp.h:47-52,121-124                  banded_cols, cells
// Every per-query extend() is a small state machine (produce DP problems -> wait -> consume results), so the chunked
// ranking loop of the reference is replayed exactly while the DP itself runs in device-wide waves.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#include "../../../include/dmnd_b200.h"
#include "chaining.h"
#include "scoring.h"

namespace {

using namespace dmnd;
using Clock = std::chrono::steady_clock;
static double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }


template<typename F>
void parallel_for(size_t n, int threads, size_t grain, F&& f) {
	if (n == 0) return;
	threads = std::max(1, std::min<int>(threads, (int)((n + grain - 1) / grain)));
	if (threads == 1) { f(0, n, 0); return; }
	std::atomic<size_t> next(0);
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; ++t)
		pool.emplace_back([&, t] {
			for (;;) {
				const size_t b = next.fetch_add(grain);
				if (b >= n) break;
				f(b, std::min(n, b + grain), t);
			}
		});
	for (auto& th : pool) th.join();
}

struct SeedHit { int i, j, score; bool operator<(const SeedHit& x) const { const int d1 = i - j, d2 = x.i - x.j; return d1 < d2 || (d1 == d2 && j < x.j); } };
struct TargetScore { uint32_t target; uint16_t score; bool operator<(const TargetScore& x) const { return score > x.score || (score == x.score && target < x.target); } };

struct HspLite { int score; double evalue; int d_begin, d_end; };
inline bool hsp_less(const HspLite& a, const HspLite& b) {  // basic/match.h:199-202 (query_source_range.begin_ is 0 in round 1)
	return a.score > b.score || (a.score == b.score && a.d_begin < b.d_begin);
}
struct Target {  // align/target.h:83-144
	uint32_t block_id; int tlen; int filter_score; double filter_evalue; std::vector<HspLite> hsp;
	static bool comp_evalue(const Target& t, const Target& u) {
		return t.filter_evalue < u.filter_evalue || (t.filter_evalue == u.filter_evalue && (t.filter_score > u.filter_score || (t.filter_score == u.filter_score && t.block_id < u.block_id)));
	}
};
struct Match {  // align/extend.h:36-70
	uint32_t target_block_id; int tlen; int filter_score; double filter_evalue; bool has_hsp; HspLite h; dmnd_dp_result r;
	std::vector<uint8_t> tr;
	static bool cmp_evalue(const Match& m, const Match& n) {
		return m.filter_evalue < n.filter_evalue || (m.filter_evalue == n.filter_evalue && (m.filter_score > n.filter_score || (m.filter_score == n.filter_score && m.target_block_id < n.target_block_id)));
	}
};

template<typename It>
It output_range(It begin, It end, int64_t max_target_seqs) {  // align/culling.cpp:90-109 (no --top)
	if (end <= begin) return begin;
	It i = begin;
	if (i->filter_evalue == DBL_MAX) return begin;
	i += std::min<ptrdiff_t>((ptrdiff_t)max_target_seqs, end - begin);
	while (--i > begin && i->filter_evalue == DBL_MAX);
	++i;
	return i;
}

int band_for(int len) {  // align/gapped_score.cpp:41-72, Mode::BANDED_FAST
	if (len < 50) return 12;
	if (len < 100) return 16;
	if (len < 250) return 30;
	if (len < 350) return 40;
	return 64;
}
inline int banded_cols(int qlen, int tlen, int d_begin, int d_end) {  // dp/dp.h:47-52
	const int pos = std::max(d_end - 1, 0) - (d_end - 1);
	return std::min(qlen - 1 - d_begin, tlen - 1) + 1 - pos;
}

struct Env {
	const Scoring* sc;
	const int8_t *q_letters, *r_letters, *bias;  // bias may be null
	const int64_t *q_limits, *r_limits;
	uint32_t nq, nr;
	int64_t ref_letters;
	int max_target_seqs;
	double max_evalue;
	int qlen(uint32_t q) const { return (int)(q_limits[q + 1] - q_limits[q] - 1); }
	int tlen(uint32_t t) const { return (int)(r_limits[t + 1] - r_limits[t] - 1); }
};

enum Phase { PH_ROUND1_PRODUCE, PH_ROUND1_CONSUME, PH_ROUND2_CONSUME, PH_DONE };

struct QueryState {
	uint32_t qid = 0;
	int qlen = 0;
	Phase phase = PH_ROUND1_PRODUCE;
	// SeedHitList (align/target.h:160-165)
	std::vector<SeedHit> seed_hits;
	std::vector<uint32_t> hit_begin;  // per target, into seed_hits (size targets+1)
	std::vector<uint32_t> target_block_ids;
	std::vector<TargetScore> target_scores;
	// extend() loop state (align/extend.cpp:259-336)
	int64_t chunk_size = 0, i0 = 0, i1 = 0;
	bool new_hits_ev = false;
	int tail_score = 0, previous_tail_score = 0;
	std::vector<Target> aligned_targets;
	std::vector<Match> matches;
	// in-flight
	std::vector<Target> r1;            // targets of the chunk in flight (the `r` of gapped_score.cpp:182)
	std::vector<uint32_t> prob_target; // per DP problem of this query: index into r1 / r2
	std::vector<Match> r2;
	size_t prob_begin = 0, prob_count = 0;  // slice of the wave's problem array
	uint64_t stat_targets = 0;

	void load_hits(const Env& e, dmnd_hit* begin, dmnd_hit* end);
	void start(const Env& e);
	void produce_round1(const Env& e, std::vector<dmnd_dp_problem>& out, uint64_t& cells);
	void consume_round1(const Env& e, const dmnd_dp_problem* probs, const dmnd_dp_result* res);
	void produce_round2(const Env& e, std::vector<dmnd_dp_problem>& out, uint64_t& cells);
	void consume_round2(const Env& e, const dmnd_dp_result* res);
	void finish_outer(const Env& e);
};

void QueryState::load_hits(const Env& e, dmnd_hit* begin, dmnd_hit* end) {
	// align/load_hits.h:44-122
	std::sort(begin, end, [](const dmnd_hit& a, const dmnd_hit& b) {
		const uint64_t sa = DMND_HIT_SUBJECT(a), sb = DMND_HIT_SUBJECT(b);
		return sa < sb || (sa == sb && (a.query < b.query || (a.query == b.query && a.seed_offset < b.seed_offset)));
	});
	uint32_t target = UINT32_MAX;
	uint16_t score = 0;
	for (dmnd_hit* h = begin; h < end; ++h) {
		const uint64_t subj = DMND_HIT_SUBJECT(*h);
		// SequenceSet::local_position: sequence whose [limits[t], limits[t+1]) holds subj
		const uint32_t t = (uint32_t)(std::upper_bound(e.r_limits, e.r_limits + e.nr + 1, (int64_t)subj) - e.r_limits) - 1;
		if (t != target) {
			if (target != UINT32_MAX) { target_scores.push_back({ (uint32_t)target_block_ids.size() - 1, score }); score = 0; }
			hit_begin.push_back((uint32_t)seed_hits.size());
			target = t;
			target_block_ids.push_back(t);
		}
		const uint16_t hs = (uint16_t)DMND_HIT_SCORE(*h);
		seed_hits.push_back({ h->seed_offset, (int)((int64_t)subj - e.r_limits[t]), (int)hs });
		score = std::max(score, hs);
	}
	if (target != UINT32_MAX) target_scores.push_back({ (uint32_t)target_block_ids.size() - 1, score });
	hit_begin.push_back((uint32_t)seed_hits.size());
}

void QueryState::start(const Env& e) {
	// align/extend.cpp:346-387 then :226-258
	qlen = e.qlen(qid);
	const int64_t target_count = (int64_t)target_block_ids.size();
	stat_targets = (uint64_t)target_count;
	if (target_count == 0) { phase = PH_DONE; return; }
	const int64_t block_mult = std::max<int64_t>((int64_t)std::round((double)e.ref_letters / 2e9), 1);
	const int64_t mm = ((int64_t)e.max_target_seqs + 31) / 32 * 32;  // make_multiple(max_target_seqs, 32)
	chunk_size = std::max<int64_t>(128, std::min<int64_t>(mm, 400)) * block_mult;
	if (chunk_size < target_count) std::sort(target_scores.begin(), target_scores.end());
	i0 = 0;
	i1 = std::min<int64_t>(chunk_size, target_count);
	if ((i1 - i0) < e.max_target_seqs)
		while (i1 < target_count && e.sc->evalue(target_scores[(size_t)i1].score, (unsigned)qlen, 50) <= e.max_evalue)
			i1 += std::min<int64_t>(16, target_count - i1);
	phase = PH_ROUND1_PRODUCE;
}

void QueryState::produce_round1(const Env& e, std::vector<dmnd_dp_problem>& out, uint64_t& cells) {
	// extend_chunk -> ungapped_stage -> align (round 1)
	r1.clear();
	prob_target.clear();
	prob_begin = out.size();
	const int8_t* query = e.q_letters + e.q_limits[qid];
	const int8_t* cbs = e.bias ? e.bias + e.q_limits[qid] : nullptr;
	const int band = band_for(qlen);
	std::vector<SeedHit> hits;
	std::vector<Segment> segs;
	std::vector<Chain> chains;
	for (int64_t k = i0; k < i1; ++k) {
		const uint32_t tix = target_scores[(size_t)k].target;
		const uint32_t block_id = target_block_ids[tix];
		const int slen = e.tlen(block_id);
		const int8_t* subject = e.r_letters + e.r_limits[block_id];
		r1.push_back(Target{ block_id, slen, 0, DBL_MAX, {} });
		hits.assign(seed_hits.begin() + hit_begin[tix], seed_hits.begin() + hit_begin[tix + 1]);
		std::sort(hits.begin(), hits.end());
		segs.clear();
		for (const SeedHit& h : hits) {  // align/ungapped.cpp:81-91
			if (!segs.empty() && segs.back().diag() == h.i - h.j && segs.back().subject_end() >= h.j) continue;
			const Segment d = xdrop_ungapped(*e.sc, query, cbs, subject, h.i, h.j);
			if (d.score > 0) segs.push_back(d);
		}
		if (segs.empty()) continue;
		std::stable_sort(segs.begin(), segs.end(), [](const Segment& x, const Segment& y) { return x.diag() < y.diag() || (x.diag() == y.diag() && x.j < y.j); });
		chain_segments(*e.sc, query, qlen, subject, slen, segs, chains);
		std::stable_sort(chains.begin(), chains.end(), [](const Chain& x, const Chain& y) { return x.d_min < y.d_min; });
		// add_dp_targets, align/gapped_score.cpp:107-180
		int d0 = INT_MAX, d1 = INT_MIN;
		auto emit = [&] {
			out.push_back(dmnd_dp_problem{ qid, block_id, d0, d1 });
			prob_target.push_back((uint32_t)r1.size() - 1);
			cells += (uint64_t)(d1 - d0) * (uint64_t)banded_cols(qlen, slen, d0, d1);
		};
		for (const Chain& h : chains) {
			const int b0 = std::max(h.d_min - band, -(slen - 1)), b1 = std::min(h.d_max + 1 + band, qlen);
			bool merge = false;
			if (d0 != INT_MAX) {
				const int ib = std::max(d0, b0), ie = std::min(d1, b1);
				const double overlap = ie > ib ? ie - ib : 0;
				merge = overlap / (d1 - d0) > 0.0 || overlap / (b1 - b0) > 0.0;
			}
			if (merge) { d0 = std::min(d0, b0); d1 = std::max(d1, b1); }
			else {
				if (d0 != INT_MAX) emit();
				d0 = b0; d1 = b1;
			}
		}
		if (!chains.empty()) emit();
	}
	prob_count = out.size() - prob_begin;
	phase = PH_ROUND1_CONSUME;
}

static void culling_targets(std::vector<Target>& targets, bool sort_only, int max_target_seqs) {  // align/culling.cpp:187-191
	std::sort(targets.begin(), targets.end(), Target::comp_evalue);
	if (!sort_only) targets.erase(output_range(targets.begin(), targets.end(), max_target_seqs), targets.end());
}

void QueryState::consume_round1(const Env& e, const dmnd_dp_problem* probs, const dmnd_dp_result* res) {
	// tail of align() round 1 (gapped_score.cpp:231-246): add_hit, inner_culling, drop targets without hits
	for (size_t k = 0; k < prob_count; ++k) {
		const int score = res[prob_begin + k].score;
		Target& t = r1[prob_target[k]];
		const double ev = e.sc->evalue(score, (unsigned)qlen, (unsigned)t.tlen);
		if (score > 0 && ev <= e.max_evalue) {  // banded_swipe.h:341-342, ScoreMatrix::report_cutoff
			const dmnd_dp_problem& pr = probs[prob_begin + k];
			t.hsp.push_back(HspLite{ score, ev, pr.d_begin, pr.d_end });  // Hsp::d_begin/d_end = the DpTarget's band
			if (score > t.filter_score) { t.filter_evalue = ev; t.filter_score = score; }  // Target::add_hit(list,it), target.h:104-112
		}
	}
}

}  // namespace

// The remaining state-machine steps need the wave's problem array, so they live in the driver below.
namespace {

struct Driver {
	dmnd_ctx* ctx;
	dmnd_block *qb;
	const dmnd_block* rb;
	Env env;
	dmnd_search_opts opts;
	int host_threads;
	dmnd_run_stats stats;

	std::vector<QueryState> qs;

	int run_waves(std::vector<dmnd_match>& out_matches, std::vector<uint8_t>& out_tr);
};

int Driver::run_waves(std::vector<dmnd_match>& out_matches, std::vector<uint8_t>& out_tr) {
	const Env& e = env;
	std::vector<size_t> active(qs.size());
	for (size_t i = 0; i < qs.size(); ++i) active[i] = i;
	const int T = host_threads;
	std::vector<std::vector<dmnd_dp_problem>> tl_p1((size_t)T), tl_p2((size_t)T);
	std::vector<uint64_t> tl_c1((size_t)T), tl_c2((size_t)T);
	std::vector<dmnd_dp_problem> p1, p2;
	std::vector<dmnd_dp_result> res1, res2;
	std::vector<uint8_t> tr;

	while (!active.empty()) {
		auto t0 = Clock::now();
		for (auto& v : tl_p1) v.clear();
		for (auto& v : tl_p2) v.clear();
		std::fill(tl_c1.begin(), tl_c1.end(), 0);
		std::fill(tl_c2.begin(), tl_c2.end(), 0);
		// ---- produce: every active query emits the DP problems of its next step into a per-thread list
		std::vector<int> owner(active.size());
		parallel_for(active.size(), T, 64, [&](size_t b, size_t en, int t) {
			for (size_t a = b; a < en; ++a) {
				QueryState& q = qs[active[a]];
				owner[a] = t;
				if (q.phase == PH_ROUND1_PRODUCE) q.produce_round1(e, tl_p1[(size_t)t], tl_c1[(size_t)t]);
				else if (q.phase == PH_ROUND2_CONSUME) {
					// round-2 problems (gapped_final.cpp:64-78,118-124): one per kept HSP of every culled target
					q.prob_begin = tl_p2[(size_t)t].size();
					q.prob_target.clear();
					q.r2.clear();
					for (const Target& tg : q.aligned_targets) {
						for (const HspLite& h : tg.hsp) {
							tl_p2[(size_t)t].push_back(dmnd_dp_problem{ q.qid, tg.block_id, h.d_begin, h.d_end });
							q.prob_target.push_back((uint32_t)q.r2.size());
							tl_c2[(size_t)t] += (uint64_t)(h.d_end - h.d_begin) * (uint64_t)banded_cols(q.qlen, tg.tlen, h.d_begin, h.d_end);
						}
						Match m{};
						m.target_block_id = tg.block_id; m.tlen = tg.tlen; m.filter_score = 0; m.filter_evalue = DBL_MAX; m.has_hsp = false;
						q.r2.push_back(m);
					}
					q.prob_count = tl_p2[(size_t)t].size() - q.prob_begin;
				}
			}
		});
		// ---- concatenate (per-thread offsets)
		std::vector<size_t> off1((size_t)T + 1, 0), off2((size_t)T + 1, 0);
		for (int t = 0; t < T; ++t) { off1[(size_t)t + 1] = off1[(size_t)t] + tl_p1[(size_t)t].size(); off2[(size_t)t + 1] = off2[(size_t)t] + tl_p2[(size_t)t].size(); }
		p1.resize(off1[(size_t)T]); p2.resize(off2[(size_t)T]);
		for (int t = 0; t < T; ++t) {
			if (!tl_p1[(size_t)t].empty()) std::memcpy(p1.data() + off1[(size_t)t], tl_p1[(size_t)t].data(), tl_p1[(size_t)t].size() * sizeof(dmnd_dp_problem));
			if (!tl_p2[(size_t)t].empty()) std::memcpy(p2.data() + off2[(size_t)t], tl_p2[(size_t)t].data(), tl_p2[(size_t)t].size() * sizeof(dmnd_dp_problem));
			stats.cells_round1 += tl_c1[(size_t)t]; stats.cells_round2 += tl_c2[(size_t)t];
		}
		for (size_t a = 0; a < active.size(); ++a) {
			QueryState& q = qs[active[a]];
			if (q.phase == PH_ROUND1_CONSUME) q.prob_begin += off1[(size_t)owner[a]];
			else if (q.phase == PH_ROUND2_CONSUME) q.prob_begin += off2[(size_t)owner[a]];
		}
		stats.dp_problems_round1 += p1.size(); stats.dp_problems_round2 += p2.size();
		stats.host_bridge_ms += ms_since(t0);
		// ---- device waves
		res1.resize(p1.size()); res2.resize(p2.size());
		t0 = Clock::now();
		if (!p1.empty() && dmnd_banded_swipe(ctx, qb, rb, p1.data(), p1.size(), DMND_DP_SCORE_ONLY, res1.data(), nullptr, 0)) return 1;
		stats.dp1_ms += ms_since(t0);
		t0 = Clock::now();
		if (!p2.empty()) {
			uint8_t* trp = nullptr; size_t cap = 0;
			if (opts.want_transcript) {
				for (const dmnd_dp_problem& pr : p2) cap += (size_t)e.qlen(pr.query) + (size_t)e.tlen(pr.target);
				tr.resize(cap); trp = tr.data();
			}
			if (dmnd_banded_swipe(ctx, qb, rb, p2.data(), p2.size(), DMND_DP_TRACEBACK, res2.data(), trp, cap)) return 1;
		}
		stats.dp2_ms += ms_since(t0);
		t0 = Clock::now();
		// ---- consume
		parallel_for(active.size(), T, 64, [&](size_t b, size_t en, int) {
			for (size_t a = b; a < en; ++a) {
				QueryState& q = qs[active[a]];
				if (q.phase == PH_ROUND1_CONSUME) {
					q.consume_round1(e, p1.data(), res1.data());
					// inner_culling (max_hsps == 1), keep targets with hits
					std::vector<Target> v;
					for (Target& t : q.r1) {
						if (t.filter_evalue == DBL_MAX) continue;
						std::stable_sort(t.hsp.begin(), t.hsp.end(), hsp_less);  // Target::inner_culling, culling.cpp:60-70
						t.hsp.resize(1);
						v.push_back(std::move(t));
					}
					q.r1.clear();
					// align/extend.cpp:307-320
					const int64_t n_targets = (int64_t)q.target_scores.size();
					const bool multi_chunk = (q.i1 - q.i0) < n_targets;
					bool new_hits = q.new_hits_ev = !v.empty();
					if (multi_chunk) {
						// append_hits, align/culling.cpp:111-141 (with_culling = first_round_culling = true)
						if (v.empty()) new_hits = false;
						else {
							new_hits = (int64_t)q.aligned_targets.size() < e.max_target_seqs;
							bool append = new_hits;
							culling_targets(q.aligned_targets, append, e.max_target_seqs);
							double min_evalue = DBL_MAX;
							for (const Target& t : v) min_evalue = std::min(min_evalue, t.filter_evalue);
							auto range_end = output_range(q.aligned_targets.begin(), q.aligned_targets.end(), e.max_target_seqs);
							if (q.aligned_targets.empty() || min_evalue <= (range_end - 1)->filter_evalue) { append = true; new_hits = true; }
							if (append) for (Target& t : v) q.aligned_targets.push_back(std::move(t));
						}
					}
					else q.aligned_targets = std::move(v);
					q.i0 = q.i1;
					q.i1 += std::min<int64_t>(q.chunk_size, n_targets - q.i1);
					q.previous_tail_score = q.tail_score;
					if (new_hits) q.tail_score = q.target_scores[(size_t)(q.i1 - 1)].score;
					bool terminate = false;
					if (q.i0 < n_targets) {  // ranking_terminate, align/extend.cpp:111-119
						const int ts = q.target_scores[(size_t)(q.i1 - 1)].score;
						terminate = !new_hits && (q.previous_tail_score == 0 || double(ts) / (double)q.previous_tail_score <= 0.95 || e.sc->bitscore(ts) < 25.0);
					}
					if (q.i0 < n_targets && !terminate) { q.phase = PH_ROUND1_PRODUCE; continue; }
					culling_targets(q.aligned_targets, false, e.max_target_seqs);  // extend.cpp:331
					if (q.aligned_targets.empty()) {
						// round 2 over nothing: align() returns no matches; evaluate the outer loop condition
						q.finish_outer(e);
					}
					else q.phase = PH_ROUND2_CONSUME;  // problems are produced at the top of the next wave
				}
				else if (q.phase == PH_ROUND2_CONSUME && q.prob_target.size() == q.prob_count && !q.r2.empty()) {
					// gapped_final.cpp:140-149
					for (size_t k = 0; k < q.prob_count; ++k) {
						const dmnd_dp_result& r = res2[q.prob_begin + k];
						Match& m = q.r2[q.prob_target[k]];
						const double ev = e.sc->evalue(r.score, (unsigned)q.qlen, (unsigned)m.tlen);
						if (r.score > 0 && ev <= e.max_evalue) {
							const dmnd_dp_problem& pr = p2[q.prob_begin + k];
							const HspLite h{ r.score, ev, pr.d_begin, pr.d_end };
							if (!m.has_hsp || hsp_less(h, m.h)) {
								m.h = h; m.r = r;
								if (opts.want_transcript && r.status == 0) m.tr.assign(tr.begin() + r.transcript_off, tr.begin() + r.transcript_off + r.transcript_len);
							}
							m.has_hsp = true;
							if (r.score > m.filter_score) { m.filter_evalue = ev; m.filter_score = r.score; }
						}
					}
					for (Match& m : q.r2) if (m.has_hsp) { m.filter_evalue = m.h.evalue; m.filter_score = m.h.score; }  // Match::inner_culling
					std::sort(q.r2.begin(), q.r2.end(), Match::cmp_evalue);  // culling(r, cfg), culling.cpp:199-202
					q.r2.erase(output_range(q.r2.begin(), q.r2.end(), e.max_target_seqs), q.r2.end());
					for (Match& m : q.r2) q.matches.push_back(m);
					q.r2.clear();
					q.aligned_targets.clear();
					q.finish_outer(e);
				}
			}
		});
		std::vector<size_t> next_active;
		for (size_t a : active) if (qs[a].phase != PH_DONE) next_active.push_back(a);
		active.swap(next_active);
		stats.host_bridge_ms += ms_since(t0);
	}
	// ---- emit
	for (QueryState& q : qs) {
		if (q.matches.empty()) continue;
		++stats.queries_aligned;
		for (const Match& m : q.matches) {
			dmnd_match o{};
			o.query = q.qid; o.target = m.target_block_id; o.score = m.h.score; o.evalue = m.h.evalue;
			o.bit_score = e.sc->bitscore(m.h.score);
			o.q_begin = m.r.q_begin; o.q_end = m.r.q_end; o.t_begin = m.r.t_begin; o.t_end = m.r.t_end;
			o.identities = m.r.identities; o.mismatches = m.r.mismatches; o.gap_openings = m.r.gap_openings;
			o.length = m.r.length; o.gaps = m.r.gaps; o.positives = m.r.positives;
			o.transcript_off = out_tr.size(); o.transcript_len = (uint32_t)m.tr.size();
			out_tr.insert(out_tr.end(), m.tr.begin(), m.tr.end());
			out_matches.push_back(o);
		}
	}
	stats.matches = out_matches.size();
	return 0;
}

}  // namespace

void QueryState::finish_outer(const Env& e) {
	// outer do-while of extend(), align/extend.cpp:336, then the final culling (:341)
	const int64_t n_targets = (int64_t)target_scores.size();
	if ((int64_t)matches.size() < e.max_target_seqs && i0 < n_targets && new_hits_ev) { phase = PH_ROUND1_PRODUCE; return; }
	std::sort(matches.begin(), matches.end(), Match::cmp_evalue);
	matches.erase(output_range(matches.begin(), matches.end(), e.max_target_seqs), matches.end());
	phase = PH_DONE;
}

// ------------------------------------------------------------------------------------------------------------
struct dmnd_result {
	std::vector<dmnd_match> matches;
	std::vector<uint8_t> transcripts;
	dmnd_run_stats stats;
};

extern "C" {

void dmnd_search_opts_default(dmnd_search_opts* o) {
	std::memset(o, 0, sizeof *o);
	o->sensitivity = 0; o->threads = 8; o->index_chunks = 0; o->comp_based_stats = 1; o->max_target_seqs = 25;
	o->max_evalue = 0.001; o->db_letters = 0; o->want_transcript = 0;
}

int dmnd_params_init(const dmnd_search_opts* o, dmnd_params* p) {
	std::memset(p, 0, sizeof *p);
	Scoring sc;
	std::memcpy(p->score, sc.m8, sizeof p->score);
	p->gap_open = sc.gap_open; p->gap_extend = sc.gap_extend;
	// Reduction("A KR EDNQ C G H ILVM FYW P ST"), stats/stats.cpp:48, basic/basic.cpp:267-296
	static const char* groups[10] = { "A", "KR", "EDNQ", "C", "G", "H", "ILVM", "FYW", "P", "ST" };
	const char* alph = letter_alphabet();
	for (int c = 0; c < 10; ++c)
		for (const char* s = groups[c]; *s; ++s) {
			const int l = (int)(std::strchr(alph, *s) - alph);
			p->reduction[l] = (uint8_t)c; p->map8[l] = (uint8_t)c; p->map8b[l] = (uint8_t)c;
		}
	p->reduction[MASK_LETTER] = MASK_LETTER; p->reduction[STOP_LETTER] = MASK_LETTER;
	p->map8[MASK_LETTER] = p->map8[STOP_LETTER] = p->map8[DMND_DELIMITER] = 10;
	p->map8b[MASK_LETTER] = p->map8b[STOP_LETTER] = p->map8b[DMND_DELIMITER] = 11;
	p->reduction_size = 10;
	if (o->sensitivity != 0) { dmnd_set_last_error("only --fast (sensitivity 0) is wired in this build"); return 1; }
	// Sensitivity::FAST: shape_codes (search/setup.cpp:211-212), traits (:43)
	const char* codes[] = { "1101110101101111" };
	p->n_shapes = 1;
	for (int s = 0; s < p->n_shapes; ++s) {
		int w = 0, len = 0;
		for (const char* c = codes[s]; *c; ++c, ++len)
			if (*c == '1') { p->shape_pos[s][w++] = len; p->shape_mask[s] |= 1u << len; }
		p->shape_len[s] = len; p->shape_weight = w;
	}
	p->hamming_id = 11;
	p->index_chunks = o->index_chunks > 0 ? o->index_chunks : 4;
	{  // seedp_bits, search/setup.cpp:306-309
		auto bit_length = [](int64_t x) { int n = 0; while (x > 0) { ++n; x >>= 1; } return n; };
		int64_t pw = 1; for (int i = 0; i < p->shape_weight; ++i) pw *= p->reduction_size;
		const int threads = o->threads > 0 ? o->threads : 1;
		p->seedp_bits = std::max(std::max(bit_length(pw - 1) - 32, bit_length((int64_t)threads * 4 * p->index_chunks - 1)), 8);
	}
	p->seed_cut = 0.9 * std::log(2.0) * p->shape_weight;
	p->left_most_interval = 32; p->ungapped_window = 48; p->ungapped_evalue = 0.0;
	return 0;
}

static int blastp_impl(dmnd_ctx* ctx, dmnd_block* qb, const dmnd_block* rb, const int8_t* q_letters, const int64_t* q_limits,
                       uint32_t nq, const int8_t* r_letters, const int64_t* r_limits, uint32_t nr, const dmnd_search_opts* opts,
                       dmnd_result** out) {
	auto t_total = Clock::now();
	std::unique_ptr<dmnd_result> res(new dmnd_result());
	std::memset(&res->stats, 0, sizeof res->stats);
	Scoring sc;
	int64_t ref_letters = 0;
	for (uint32_t i = 0; i < nr; ++i) ref_letters += r_limits[i + 1] - r_limits[i] - 1;
	sc.db_letters = opts->db_letters ? (double)opts->db_letters : (double)ref_letters;
	int host_threads = (int)std::thread::hardware_concurrency();
	if (host_threads < 1) host_threads = 1;
	if (const char* ev = std::getenv("DMND_HOST_THREADS")) host_threads = std::max(1, std::atoi(ev));

	Driver d;
	d.ctx = ctx; d.qb = qb; d.rb = rb; d.opts = *opts; d.host_threads = host_threads;
	std::memset(&d.stats, 0, sizeof d.stats);
	dmnd_timing tm0; dmnd_timing_fetch(ctx, &tm0, 1);

	// ---- seed stage (run_ref_chunk: one search_shape per shape; FAST has one shape)
	auto t0 = Clock::now();
	dmnd_hits* hits = nullptr;
	if (dmnd_search_shape(ctx, qb, rb, 0, &hits, &d.stats.seed)) return 1;
	const size_t nh = dmnd_hits_count(hits);
	std::vector<dmnd_hit> hv(nh);
	if (nh && dmnd_hits_download(ctx, hits, hv.data(), nh)) { dmnd_hits_free(ctx, hits); return 1; }
	dmnd_hits_free(ctx, hits);
	if (dmnd_block_clear_seed_mask(ctx, qb)) return 1;  // run/double_indexed.cpp:211-212
	d.stats.seed_ms = ms_since(t0);
	d.stats.hits = nh;

	// ---- group by query (hits arrive grouped by ascending query id)
	t0 = Clock::now();
	std::vector<size_t> qstart;
	for (size_t i = 0; i < nh; ++i)
		if (i == 0 || hv[i].query != hv[i - 1].query) {
			if (i > 0 && hv[i].query < hv[i - 1].query) { dmnd_set_last_error("dmnd_blastp: hits not grouped by ascending query"); return 1; }
			qstart.push_back(i);
		}
	qstart.push_back(nh);
	const size_t nqh = qstart.size() - 1;
	d.qs.resize(nqh);

	std::vector<int8_t> bias;
	if (opts->comp_based_stats == 1) bias.assign((size_t)(q_limits[nq] + DMND_PERIMETER_PADDING), 0);
	Env& e = d.env;
	e.sc = &sc; e.q_letters = q_letters; e.r_letters = r_letters; e.bias = bias.empty() ? nullptr : bias.data();
	e.q_limits = q_limits; e.r_limits = r_limits; e.nq = nq; e.nr = nr; e.ref_letters = ref_letters;
	e.max_target_seqs = opts->max_target_seqs; e.max_evalue = opts->max_evalue;

	parallel_for(nqh, host_threads, 64, [&](size_t b, size_t en, int) {
		std::vector<int8_t> hc;
		for (size_t k = b; k < en; ++k) {
			QueryState& q = d.qs[k];
			q.qid = hv[qstart[k]].query;
			q.load_hits(e, hv.data() + qstart[k], hv.data() + qstart[k + 1]);
			q.start(e);
			if (!bias.empty()) {  // HauserCorrection per query, align/extend.cpp:247-250
				hauser_correction(sc, q_letters + q_limits[q.qid], q.qlen, hc);
				std::memcpy(bias.data() + q_limits[q.qid], hc.data(), (size_t)q.qlen);
			}
		}
	});
	for (const QueryState& q : d.qs) d.stats.targets += q.stat_targets;
	if (dmnd_block_set_bias(ctx, qb, bias.empty() ? nullptr : bias.data(), (size_t)(q_limits[nq] + DMND_PERIMETER_PADDING))) return 1;
	d.stats.host_bridge_ms += ms_since(t0);

	if (d.run_waves(res->matches, res->transcripts)) return 1;
	d.stats.total_ms = ms_since(t_total);
	dmnd_timing_fetch(ctx, &d.stats.device, 0);
	res->stats = d.stats;
	*out = res.release();
	return 0;
}

int dmnd_blastp_resident(dmnd_ctx* ctx, dmnd_block* query, const dmnd_block* ref, const int8_t* q_letters, const int64_t* q_limits,
                         uint32_t nq, const int8_t* r_letters, const int64_t* r_limits, uint32_t nr, const dmnd_search_opts* opts,
                         dmnd_result** out) {
	return blastp_impl(ctx, query, ref, q_letters, q_limits, nq, r_letters, r_limits, nr, opts, out);
}

int dmnd_blastp(dmnd_ctx* ctx, const int8_t* q_letters, size_t q_raw_len, const int64_t* q_limits, uint32_t nq,
                const int8_t* r_letters, size_t r_raw_len, const int64_t* r_limits, uint32_t nr, const dmnd_search_opts* opts,
                dmnd_result** out) {
	dmnd_block *qb = nullptr, *rb = nullptr;
	if (dmnd_block_upload(ctx, q_letters, q_raw_len, q_limits, nq, &qb)) return 1;
	if (dmnd_block_upload(ctx, r_letters, r_raw_len, r_limits, nr, &rb)) { dmnd_block_free(ctx, qb); return 1; }
	const int rc = blastp_impl(ctx, qb, rb, q_letters, q_limits, nq, r_letters, r_limits, nr, opts, out);
	dmnd_block_free(ctx, qb); dmnd_block_free(ctx, rb);
	return rc;
}

const dmnd_match* dmnd_result_matches(const dmnd_result* r, size_t* n) { *n = r->matches.size(); return r->matches.data(); }
const uint8_t* dmnd_result_transcripts(const dmnd_result* r, size_t* n) { *n = r->transcripts.size(); return r->transcripts.data(); }
const dmnd_run_stats* dmnd_result_stats(const dmnd_result* r) { return &r->stats; }
void dmnd_result_free(dmnd_result* r) { delete r; }

}  // extern "C"
