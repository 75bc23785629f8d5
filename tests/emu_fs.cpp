// emu_fs.cpp -- the 3-frame banded DP and its traceback (diamond_b200/csrc/cuda/fs_kernels.cuh, the source the GPU library is
// built from) on the CPU behind tests/emu_cuda.h against the oracle's dmnd_banded_3frame_swipe: score, frames, coordinates, counts
// and the transcript of every problem.  Random DNA-length-consistent frame triples, targets stitched from pieces of the three
// frames (so that alignments really change frame, forwards and backwards) with substitutions and indels, bands at the corners of
// the matrix, every register tile (R = 2 .. 32).
// usage: emu_fs SEED MAXDNA MAXBAND PROBLEMS [TRACE=1]
#include <cstdint>
#include <cstring>
#include "emu_cuda.h"
#include "../diamond_b200/csrc/cuda/fs_kernels.cuh"
#include <algorithm>
#include <random>
#include <vector>
using namespace dmnd_cuda;

template<int R> static void run_group(const FsArgs& a, const DevParams* P, bool trace) {
	const unsigned warps = a.n;
	if (trace) emu::launch(std::min(warps, 3u), 32, [&] { fs_swipe_kernel<R, true>(a, P); });
	else emu::launch(std::min(warps, 3u), 32, [&] { fs_swipe_kernel<R, false>(a, P); });
}

int main(int argc, char** argv) {
	const int seed = argc > 1 ? atoi(argv[1]) : 1, maxdna = argc > 2 ? atoi(argv[2]) : 600, maxw = argc > 3 ? atoi(argv[3]) : 128, nprob = argc > 4 ? atoi(argv[4]) : 100;
	const bool trace = argc > 5 ? atoi(argv[5]) != 0 : true;
	const int FS = 15;
	dmnd_search_opts o; dmnd_search_opts_default(&o);
	dmnd_params hp; dmnd_params_init(&o, &hp);
	static DevParams P; memset(&P, 0, sizeof P);
	memcpy(P.score, hp.score, 1024); P.gap_open = hp.gap_open; P.gap_extend = hp.gap_extend;
	dmnd_ctx* ctx; if (dmnd_create(0, &hp, &ctx)) return 2;
	std::mt19937 rng((unsigned)seed);
	std::vector<int8_t> qb(256, 31), tb(256, 31);
	std::vector<int64_t> qlim, tlim;
	std::vector<dmnd_dp_problem> probs;
	for (int it = 0; it < nprob; ++it) {
		const int L = 3 + (int)(rng() % (unsigned)maxdna);
		std::vector<int8_t> fr[3];
		for (int f = 0; f < 3; ++f) {
			fr[f].resize((size_t)((L - f) / 3));
			for (auto& x : fr[f]) x = (int8_t)(rng() % 100 < 3 ? 24 : rng() % 20);  // a few stop codons
		}
		// target: pieces of the frames, in order, changing frame now and then
		std::vector<int8_t> t;
		int f = (int)(rng() % 3), i = (int)(rng() % (unsigned)std::max<size_t>(1, fr[0].size() / 3));
		const int pre = (int)(rng() % 40);
		for (int k = 0; k < pre; ++k) t.push_back((int8_t)(rng() % 20));
		while (i < (int)fr[f].size() && t.size() < 2000) {
			const unsigned u = rng() % 1000;
			if (u < 25) { f = (f + 1 + (int)(rng() % 2)) % 3; continue; }                 // frame change
			if (u < 40) { i += 1 + (int)(rng() % 3); continue; }                           // letters missing from the target
			if (u < 55) { for (int k = 0, n = 1 + (int)(rng() % 3); k < n; ++k) t.push_back((int8_t)(rng() % 20)); continue; }  // extra letters
			t.push_back(u < 300 ? (int8_t)(rng() % 20) : fr[f][(size_t)i]);
			++i;
		}
		for (int k = 0, n = (int)(rng() % 40); k < n; ++k) t.push_back((int8_t)(rng() % 20));
		if (t.empty()) t.push_back(0);
		for (int ff = 0; ff < 3; ++ff) { qlim.push_back((int64_t)qb.size()); qb.insert(qb.end(), fr[ff].begin(), fr[ff].end()); qb.push_back(31); }
		tlim.push_back((int64_t)tb.size()); tb.insert(tb.end(), t.begin(), t.end()); tb.push_back(31);
		const int qlen = (int)fr[0].size(), tlen = (int)t.size();
		const int lo = -(tlen - 1), hi = std::max(qlen, lo + 1), w = 1 + (int)(rng() % (unsigned)maxw), kind = (int)(rng() % 10), c = -pre + (int)(rng() % 41) - 20;
		int d0, d1;
		if (kind < 7) { d0 = std::max(lo, c - w / 2); d1 = std::min(hi, d0 + w); }
		else if (kind == 7) { d0 = lo; d1 = std::min(hi, lo + w); }
		else if (kind == 8) { d1 = hi; d0 = std::max(lo, hi - w); }
		else { d0 = lo + (int)(rng() % (unsigned)(hi - lo)); d1 = d0 + 1; }
		if (d1 <= d0) d1 = d0 + 1;
		probs.push_back(dmnd_dp_problem{ (uint32_t)(3 * it), (uint32_t)it, d0, d1 });
	}
	qlim.push_back((int64_t)qb.size()); tlim.push_back((int64_t)tb.size());
	qb.resize(qb.size() + 256, 31); tb.resize(tb.size() + 256, 31);
	dmnd_block *bq, *bt;
	if (dmnd_block_upload(ctx, qb.data(), qb.size(), qlim.data(), (uint32_t)(3 * nprob), &bq) || dmnd_block_upload(ctx, tb.data(), tb.size(), tlim.data(), (uint32_t)nprob, &bt)) return 2;
	// ---- oracle
	std::vector<dmnd_fs_result> want((size_t)nprob), got((size_t)nprob);
	std::vector<uint64_t> toff((size_t)nprob + 1, 0), moff((size_t)nprob + 1, 0);
	std::vector<uint32_t> tcap((size_t)nprob);
	for (int i = 0; i < nprob; ++i) {
		const int ql0 = (int)(qlim[3 * i + 1] - qlim[3 * i] - 1), tlen = (int)(tlim[i + 1] - tlim[i] - 1);
		tcap[(size_t)i] = (uint32_t)(2 * tlen + ql0 + 8); toff[(size_t)i + 1] = toff[(size_t)i] + tcap[(size_t)i];
		const dmnd_dp_problem& pr = probs[(size_t)i];
		const int B = pr.d_end - pr.d_begin, i1 = std::max(pr.d_end - 1, 0), i0 = i1 + 1 - B, pos0 = i1 - (pr.d_end - 1);
		const int ncol = (B > 0 && ql0 > 0) ? std::max(std::min(tlen - pos0, ql0 - i0), 0) : 0;
		moff[(size_t)i + 1] = moff[(size_t)i] + fs_matrix_ints(B, ncol);
	}
	std::vector<uint8_t> want_ts((size_t)toff.back() + 16), got_ts((size_t)toff.back() + 16);
	if (dmnd_banded_3frame_swipe(ctx, bq, bt, probs.data(), (size_t)nprob, FS, trace ? 1 : 0, want.data(), trace ? want_ts.data() : nullptr, trace ? want_ts.size() : 0)) { printf("oracle: %s\n", dmnd_last_error()); return 2; }
	// ---- the kernels under emulation, one launch per register-tile class
	std::vector<int32_t> score((size_t)nprob, -1), maxcol((size_t)nprob, -1);
	std::vector<int32_t> matrix((size_t)moff.back() + 16, 0);
	for (int R = 2; R <= 32; R *= 2) {
		std::vector<uint32_t> order;
		for (int i = 0; i < nprob; ++i) if (fs_tile_rows(probs[(size_t)i].d_end - probs[(size_t)i].d_begin) == R) order.push_back((uint32_t)i);
		if (order.empty()) continue;
		unsigned work = 0;
		FsArgs a;
		a.q_letters = qb.data(); a.r_letters = tb.data(); a.q_limits = qlim.data(); a.r_limits = tlim.data(); a.probs = probs.data(); a.order = order.data(); a.n = (uint32_t)order.size();
		a.frame_shift = FS; a.score = score.data(); a.max_col = maxcol.data(); a.matrix = matrix.data(); a.matrix_off = moff.data(); a.matrix_base = 0; a.work = &work;
		if (R == 2) run_group<2>(a, &P, trace); else if (R == 4) run_group<4>(a, &P, trace); else if (R == 8) run_group<8>(a, &P, trace); else if (R == 16) run_group<16>(a, &P, trace); else run_group<32>(a, &P, trace);
	}
	if (trace) {
		FsWalkArgs wa;
		wa.q_letters = qb.data(); wa.r_letters = tb.data(); wa.q_limits = qlim.data(); wa.r_limits = tlim.data(); wa.probs = probs.data(); wa.n = (uint32_t)nprob; wa.pos0 = 0;
		wa.frame_shift = FS; wa.score = score.data(); wa.max_col = maxcol.data(); wa.matrix = matrix.data(); wa.matrix_off = moff.data(); wa.matrix_base = 0;
		wa.res = got.data(); wa.transcripts = got_ts.data(); wa.transcript_off = toff.data(); wa.transcript_cap = tcap.data();
		emu::launch(((unsigned)nprob + 127) / 128, 128, [&] { fs_walk_kernel(wa, &P); });
	}
	int fails = 0, pos = 0, shifted = 0;
	for (int i = 0; i < nprob; ++i) {
		const dmnd_fs_result &ro = want[(size_t)i], &re = got[(size_t)i];
		bool same;
		if (!trace) {  // score and the first column that reaches it (dmnd_fs_result.t_end of the score-only mode)
			const dmnd_dp_problem& pr = probs[(size_t)i];
			const int i1 = std::max(pr.d_end - 1, 0), pos0 = i1 - (pr.d_end - 1);
			same = ro.score == score[(size_t)i] && (ro.score <= 0 || ro.t_end == pos0 + maxcol[(size_t)i] + 1);
		}
		else {
			same = ro.score == re.score && ro.status == re.status && ro.q_begin == re.q_begin && ro.q_end == re.q_end && ro.frame_begin == re.frame_begin && ro.frame_end == re.frame_end
				&& ro.t_begin == re.t_begin && ro.t_end == re.t_end && ro.identities == re.identities && ro.mismatches == re.mismatches && ro.gap_openings == re.gap_openings
				&& ro.length == re.length && ro.gaps == re.gaps && ro.positives == re.positives && ro.transcript_len == re.transcript_len;
			if (same && ro.transcript_len) same = memcmp(want_ts.data() + ro.transcript_off, got_ts.data() + re.transcript_off, ro.transcript_len) == 0;
			if (ro.transcript_len) for (uint32_t k = 0; k < ro.transcript_len; ++k) if (want_ts[ro.transcript_off + k] == DMND_TR_FRAMESHIFT_FWD || want_ts[ro.transcript_off + k] == DMND_TR_FRAMESHIFT_REV) { ++shifted; break; }
		}
		if (ro.score > 0) ++pos;
		if (!same) {
			++fails;
			const dmnd_dp_problem& pr = probs[(size_t)i];
			if (fails < 8) printf("MISMATCH p=%d band [%d,%d) oracle score=%d q[%d,%d) f %d>%d t[%d,%d) len %d st %d | kernel score=%d maxcol %d q[%d,%d) f %d>%d t[%d,%d) len %d st %d\n", i, pr.d_begin, pr.d_end, ro.score, ro.q_begin, ro.q_end,
				ro.frame_begin, ro.frame_end, ro.t_begin, ro.t_end, ro.length, ro.status, trace ? re.score : score[(size_t)i], maxcol[(size_t)i], re.q_begin, re.q_end, re.frame_begin, re.frame_end, re.t_begin, re.t_end, re.length, re.status);
		}
	}
	printf("problems=%d positive=%d with_frameshift=%d fails=%d \n", nprob, pos, shifted, fails);
	return fails ? 1 : 0;
}
