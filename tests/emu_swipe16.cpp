// emu_swipe16.cpp -- the packed 16-bit banded SWIPE kernel and the traceback walk (diamond_b200/csrc/cuda/swipe16.cuh, the source
// the GPU library is built from) on the CPU behind tests/emu_cuda.h, four problems per warp in lock step, against the oracle's
// dmnd_banded_swipe: score, coordinates, identities / mismatches / gap openings / length and the transcript of every problem.
// Random problems with real neighbouring sequences on both sides, Hauser-like biases, masked letters, corner bands, every
// register tile (R = 4, 8, 12, 16) and warps whose four problems differ in size.
// usage: emu_swipe16 SEED MAXQ MAXBAND PROBLEMS [TRACE=1]
#include <cstdint>
#include <cstring>
static int8_t g_emu_smem[1 << 18];  // the CTA's dynamic shared memory (blocks run one after the other)
#define DMND_DYN_SMEM(name) int8_t* name = g_emu_smem
#define __cvta_generic_to_shared(p) ((size_t)((const int8_t*)(p) - g_emu_smem))
#define DMND_S16_LDS
static inline unsigned s16_lds(unsigned addr) { return (uint8_t)g_emu_smem[addr]; }
#include "emu_cuda.h"
#include "../diamond_b200/csrc/cuda/swipe16.cuh"
#include <algorithm>
#include <numeric>
#include <random>
#include <vector>
using namespace dmnd_cuda;

template<int R> static void run_group(const SwipeArgs& a, const DevParams* P, const S16Args& sa, bool trace) {
	const unsigned warps = (a.n + 3) / 4;
	if (trace) emu::launch(std::min(warps, 3u), 32, [&] { swipe16_kernel<R, true>(a, P, sa); });
	else emu::launch(std::min(warps, 3u), 32, [&] { swipe16_kernel<R, false>(a, P, sa); });
}

int main(int argc, char** argv) {
	const int seed = argc > 1 ? atoi(argv[1]) : 1, maxq = argc > 2 ? atoi(argv[2]) : 300, maxw = argc > 3 ? atoi(argv[3]) : 128, nprob = argc > 4 ? atoi(argv[4]) : 200;
	const bool trace = argc > 5 ? atoi(argv[5]) != 0 : true;
	dmnd_search_opts o; dmnd_search_opts_default(&o);
	dmnd_params hp; dmnd_params_init(&o, &hp);
	static DevParams P; memset(&P, 0, sizeof P);
	memcpy(P.score, hp.score, 1024); P.gap_open = hp.gap_open; P.gap_extend = hp.gap_extend; P.one = 1; P.k65536 = 65536; P.neg2 = 0x80008000u;
	dmnd_ctx* ctx; if (dmnd_create(0, &hp, &ctx)) return 2;
	std::mt19937 rng((unsigned)seed);
	// ---- blocks: nprob queries and nprob targets, real letters around every sequence
	std::vector<int8_t> qb(256, 31), tb(256, 31), bias;
	std::vector<int64_t> qlim, tlim;
	std::vector<dmnd_dp_problem> probs;
	for (int it = 0; it < nprob; ++it) {
		const int qlen = 5 + (int)(rng() % (unsigned)maxq), tlen = 5 + (int)(rng() % (unsigned)(maxq * 4 / 3));
		qlim.push_back((int64_t)qb.size()); tlim.push_back((int64_t)tb.size());
		const size_t qo = qb.size(), to = tb.size();
		for (int i = 0; i < tlen; ++i) tb.push_back((int8_t)(rng() % 20));
		const int off = (int)(rng() % (unsigned)std::max(1, tlen - qlen + 1));
		for (int i = 0; i < qlen; ++i) qb.push_back((rng() % 100 < 75 && off + i < tlen) ? tb[to + (size_t)(off + i)] : (int8_t)(rng() % 20));
		if (rng() % 5 == 0) qb[qo + rng() % (unsigned)qlen] = 23;
		if (rng() % 7 == 0) tb[to + rng() % (unsigned)tlen] = (int8_t)(20 + rng() % 6);
		qb.push_back(31); tb.push_back(31);
		const int lo = -(tlen - 1), hi = qlen, c = -off + (int)(rng() % 41) - 20, w = 1 + (int)(rng() % (unsigned)maxw), kind = (int)(rng() % 10);
		int d0, d1;
		if (kind < 7) { d0 = std::max(lo, c - w / 2); d1 = std::min(hi, d0 + w); }
		else if (kind == 7) { d0 = lo; d1 = std::min(hi, lo + w); }
		else if (kind == 8) { d1 = hi; d0 = std::max(lo, hi - w); }
		else { d0 = lo + (int)(rng() % (unsigned)(hi - lo)); d1 = d0 + 1; }
		if (d1 <= d0) d1 = d0 + 1;
		probs.push_back(dmnd_dp_problem{ (uint32_t)it, (uint32_t)it, d0, d1 });
	}
	qlim.push_back((int64_t)qb.size()); tlim.push_back((int64_t)tb.size());
	qb.resize(qb.size() + 256, 31); tb.resize(tb.size() + 256, 31);
	bias.resize(qb.size());
	for (auto& x : bias) x = (int8_t)((int)(rng() % 9) - 5);
	dmnd_block *bq, *bt;
	if (dmnd_block_upload(ctx, qb.data(), qb.size(), qlim.data(), (uint32_t)nprob, &bq) || dmnd_block_upload(ctx, tb.data(), tb.size(), tlim.data(), (uint32_t)nprob, &bt)) return 2;
	dmnd_block_set_bias(ctx, bq, bias.data(), bias.size());
	// ---- oracle
	std::vector<dmnd_dp_result> want((size_t)nprob);
	size_t tcap = 0;
	for (int i = 0; i < nprob; ++i) tcap += (size_t)(qlim[i + 1] - qlim[i] - 1) + (size_t)(tlim[i + 1] - tlim[i] - 1);
	std::vector<uint8_t> want_ts(tcap + 16);
	if (dmnd_banded_swipe(ctx, bq, bt, probs.data(), (size_t)nprob, trace ? 1 : 0, want.data(), trace ? want_ts.data() : nullptr, trace ? want_ts.size() : 0)) { printf("oracle: %s\n", dmnd_last_error()); return 2; }
	// ---- the kernel under emulation: group by register tile, order by macro steps (as prep_kernel + the device sort do)
	std::vector<uint8_t> table((size_t)S16_TABLE_ENTRIES);
	unsigned bad = 0;
	emu::launch((S16_TABLE_ENTRIES + 255) / 256, 256, [&] { s16_table_kernel(&P, table.data(), &bad); });
	if (bad) { printf("table out of int8\n"); return 2; }
	std::vector<int32_t> score((size_t)nprob, -1), endc((size_t)nprob * 2, -1);
	std::vector<dmnd_dp_result> got((size_t)nprob);
	std::vector<uint64_t> ts_off((size_t)nprob);
	{ uint64_t run = 0; for (int i = 0; i < nprob; ++i) { ts_off[(size_t)i] = run; run += (uint64_t)(qlim[i + 1] - qlim[i] - 1) + (uint64_t)(tlim[i + 1] - tlim[i] - 1); } }
	std::vector<uint8_t> got_ts(tcap + 16);
	unsigned overflow = 0;
	int done = 0, skipped = 0;
	for (int R = 4; R <= 16; R += 4) {
		std::vector<uint32_t> order; std::vector<int> nmacro((size_t)nprob, 0);
		int maxqlen = 0;
		for (int i = 0; i < nprob; ++i) {
			const int B = probs[(size_t)i].d_end - probs[(size_t)i].d_begin;
			if (B > S16_MAX_BAND) { if (R == 4) ++skipped; continue; }
			if (s16_rows(B) != R) continue;
			const int qlen = (int)(qlim[i + 1] - qlim[i] - 1), tlen = (int)(tlim[i + 1] - tlim[i] - 1);
			const int i1 = std::max(probs[(size_t)i].d_end - 1, 0), j0 = i1 - (probs[(size_t)i].d_end - 1), cols = std::min(qlen - 1 - probs[(size_t)i].d_begin, tlen - 1) + 1 - j0;
			nmacro[(size_t)i] = cols > 0 ? (2 * (cols - 1) + B + 1) >> 1 : 0;
			maxqlen = std::max(maxqlen, qlen);
			order.push_back((uint32_t)i);
		}
		if (order.empty()) continue;
		// every fourth warp mixes sizes on purpose; the rest is sorted heaviest first
		std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return nmacro[x] > nmacro[y]; });
		if (order.size() > 8) std::swap(order[1], order[order.size() - 1]);
		std::vector<uint64_t> excl(order.size() + 1, 0);
		for (size_t k = 0; k < order.size(); ++k) excl[k + 1] = excl[k] + s16_trace_bytes(R, (unsigned long long)nmacro[order[k]]);
		std::vector<uint8_t> tr((size_t)excl.back() + 64, 0xA5);
		unsigned work = 0;
		SwipeArgs a;
		a.q_letters = qb.data(); a.q_bias = bias.data(); a.r_letters = tb.data(); a.q_limits = qlim.data(); a.r_limits = tlim.data();
		a.probs = probs.data(); a.order = order.data(); a.n = (uint32_t)order.size(); a.score = score.data(); a.end_cell = endc.data();
		a.trace = tr.data(); a.trace_excl = excl.data(); a.trace_base = 0; a.order_pos0 = 0; a.work = &work;
		S16Args sa; sa.table = table.data(); sa.qstride = (maxqlen + 8 * R + 4 + S16_TILE + 7) & ~7; sa.overflow = &overflow;
		if (R == 4) run_group<4>(a, &P, sa, trace); else if (R == 8) run_group<8>(a, &P, sa, trace); else if (R == 12) run_group<12>(a, &P, sa, trace); else run_group<16>(a, &P, sa, trace);
		if (trace) {
			WalkArgs wa;
			wa.q_letters = a.q_letters; wa.q_bias = a.q_bias; wa.r_letters = a.r_letters; wa.q_limits = a.q_limits; wa.r_limits = a.r_limits;
			wa.probs = a.probs; wa.order = a.order; wa.n = a.n; wa.score = a.score; wa.end_cell = a.end_cell; wa.trace = a.trace; wa.trace_excl = a.trace_excl;
			wa.trace_base = 0; wa.order_pos0 = 0; wa.res = got.data(); wa.s16 = 1; wa.transcripts = got_ts.data(); wa.transcript_off = ts_off.data();
			emu::launch((a.n + 127) / 128, 128, [&] { walk_kernel(wa, &P); });
		}
		done += (int)order.size();
	}
	if (overflow) { printf("overflow flag raised\n"); return 1; }
	int fails = 0, pos = 0;
	for (int i = 0; i < nprob; ++i) {
		const dmnd_dp_problem& pr = probs[(size_t)i];
		if (pr.d_end - pr.d_begin > S16_MAX_BAND) continue;
		const dmnd_dp_result &ro = want[(size_t)i], &re = got[(size_t)i];
		bool same;
		if (!trace) same = ro.score == score[(size_t)i];
		else {
			same = ro.score == re.score && ro.q_begin == re.q_begin && ro.q_end == re.q_end && ro.t_begin == re.t_begin && ro.t_end == re.t_end && ro.identities == re.identities
				&& ro.mismatches == re.mismatches && ro.gap_openings == re.gap_openings && ro.length == re.length && re.status == 0 && ro.transcript_len == re.transcript_len;
			if (same && ro.transcript_len) same = memcmp(want_ts.data() + ro.transcript_off, got_ts.data() + re.transcript_off, ro.transcript_len) == 0;
		}
		if (ro.score > 0) ++pos;
		if (!same) {
			++fails;
			if (fails < 8) printf("MISMATCH p=%d band [%d,%d) oracle score=%d q[%d,%d) t[%d,%d) len %d | kernel score=%d (end cell %d,%d) q[%d,%d) t[%d,%d) len %d status=%d\n", i, pr.d_begin, pr.d_end, ro.score, ro.q_begin, ro.q_end,
				ro.t_begin, ro.t_end, ro.length, trace ? re.score : score[(size_t)i], endc[2 * (size_t)i], endc[2 * (size_t)i + 1], re.q_begin, re.q_end, re.t_begin, re.t_end, re.length, re.status);
		}
	}
	printf("problems=%d skipped=%d fails=%d positives=%d \n", done, skipped, fails, pos);
	return fails ? 1 : 0;
}
