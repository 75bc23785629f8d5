// fs.cu -- host side of dmnd_banded_3frame_swipe (frameshift alignment, fs_kernels.cuh): register-tile classes, score-matrix arena
// in slices, launches.  The geometry of every problem is known on the host (block limits), so there is no device preparation.
#include "ctx.cuh"
#include "fs_kernels.cuh"
#include <algorithm>
#include <cstring>

namespace dmnd_cuda {

template<bool TRACE>
static void launch_fs(int R, const FsArgs& a, const DevParams* P, int grid, cudaStream_t st) {
	switch (R) {
	case 2: fs_swipe_kernel<2, TRACE><<<grid, 128, 0, st>>>(a, P); break;
	case 4: fs_swipe_kernel<4, TRACE><<<grid, 128, 0, st>>>(a, P); break;
	case 8: fs_swipe_kernel<8, TRACE><<<grid, 128, 0, st>>>(a, P); break;
	case 16: fs_swipe_kernel<16, TRACE><<<grid, 128, 0, st>>>(a, P); break;
	default: fs_swipe_kernel<32, TRACE><<<grid, 128, 0, st>>>(a, P); break;
	}
}

static int fs_impl(dmnd_ctx* ctx, const dmnd_block* query, const dmnd_block* ref, const dmnd_dp_problem* problems, size_t n, int frame_shift, int mode,
                   dmnd_fs_result* results, uint8_t* transcripts, size_t transcript_cap) {
	if (n == 0) return 0;
	if (n > 0x7ffffff0ull) { set_error("dmnd_banded_3frame_swipe: too many problems in one call"); return 1; }
	if (frame_shift <= 0) { set_error("dmnd_banded_3frame_swipe: the frame shift penalty must be positive"); return 1; }
	const bool trace = mode == DMND_DP_TRACEBACK;
	cudaStream_t st = ctx->stream;
	PhaseTimer timer(ctx, trace ? PH_DP_TRACE : PH_DP_SCORE);
	// ---- geometry on the host
	const int64_t* ql = query->h_limits.data(); const int64_t* rl = ref->h_limits.data();
	std::vector<uint64_t> moff(n + 1, 0), toff(n + 1, 0);
	std::vector<uint32_t> tcap(n, 0);
	std::vector<uint8_t> cls(n, 0);
	uint64_t cells = 0;
	for (size_t k = 0; k < n; ++k) {
		const dmnd_dp_problem& pr = problems[k];
		if ((uint64_t)pr.query + 3 > query->nseq || pr.target >= ref->nseq) { set_error("dmnd_banded_3frame_swipe: sequence index out of range"); return 1; }
		const int B = pr.d_end - pr.d_begin;
		if (B > DMND_FS_MAX_BAND) { set_error("dmnd_banded_3frame_swipe: band wider than 1024 diagonals is not supported by this build"); return 1; }
		const int ql0 = (int)(ql[pr.query + 1] - ql[pr.query] - 1), tlen = (int)(rl[pr.target + 1] - rl[pr.target] - 1);
		const int i1 = std::max(pr.d_end - 1, 0), i0 = i1 + 1 - B, pos0 = i1 - (pr.d_end - 1);
		const int ncol = (B > 0 && ql0 > 0) ? std::max(std::min(tlen - pos0, ql0 - i0), 0) : 0;
		const int R = fs_tile_rows(std::max(B, 1));
		cls[k] = (uint8_t)(R == 2 ? 0 : R == 4 ? 1 : R == 8 ? 2 : R == 16 ? 3 : 4);
		moff[k + 1] = moff[k] + (trace ? fs_matrix_ints(std::max(B, 0), ncol) : 0);
		tcap[k] = (uint32_t)(2 * tlen + ql0 + 8);
		toff[k + 1] = toff[k] + (trace && transcripts ? tcap[k] : 0);
		cells += (uint64_t)(3 * std::max(B, 0)) * (uint64_t)ncol;
	}
	if (trace && transcripts && toff[n] > transcript_cap) { set_error("dmnd_banded_3frame_swipe: transcript buffer too small (2 * target length + query codons + 8 bytes per problem)"); return 1; }
	if (trace && toff[n] > 0xffffffffull) { set_error("dmnd_banded_3frame_swipe: more than 4 GiB of transcripts in one call"); return 1; }
	(trace ? ctx->dp_cells_trace : ctx->dp_cells_score) += cells;
	ctx->dp_cells_padded += cells;
	// ---- device buffers: problems, per-problem offsets, scores, results
	if (ctx->b_probs.ensure(n * sizeof(dmnd_dp_problem)) || ctx->b_prep.ensure(n * (8 + 8 + 4 + 4 + 4 + 4) + 256) || ctx->b_results.ensure(n * sizeof(dmnd_fs_result)) || ctx->b_work.ensure(64)) return 1;
	dmnd_dp_problem* d_probs = ctx->b_probs.as<dmnd_dp_problem>();
	uint64_t* d_moff = ctx->b_prep.as<uint64_t>();
	uint64_t* d_toff = d_moff + n;
	uint32_t* d_tcap = reinterpret_cast<uint32_t*>(d_toff + n);
	int32_t* d_score = reinterpret_cast<int32_t*>(d_tcap + n);
	int32_t* d_maxcol = d_score + n;
	uint32_t* d_order = reinterpret_cast<uint32_t*>(d_maxcol + n);
	DMND_CUDA_CHECK(cudaMemcpyAsync(d_probs, problems, n * sizeof(dmnd_dp_problem), cudaMemcpyHostToDevice, st));
	DMND_CUDA_CHECK(cudaMemcpyAsync(d_moff, moff.data(), n * 8, cudaMemcpyHostToDevice, st));
	DMND_CUDA_CHECK(cudaMemcpyAsync(d_toff, toff.data(), n * 8, cudaMemcpyHostToDevice, st));
	DMND_CUDA_CHECK(cudaMemcpyAsync(d_tcap, tcap.data(), n * 4, cudaMemcpyHostToDevice, st));
	ctx->h2d_bytes += n * (sizeof(dmnd_dp_problem) + 20);
	uint8_t* d_tr = nullptr;
	if (trace && transcripts) { if (ctx->b_tr.ensure((size_t)toff[n] + 16)) return 1; d_tr = ctx->b_tr.as<uint8_t>(); }
	// ---- slices: consecutive problems whose score matrices fit the arena (40 % of the free memory, at most 16 GiB)
	size_t free_b = 0, total_b = 0;
	DMND_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
	const uint64_t arena_ints = std::max<uint64_t>((uint64_t)1 << 22, std::min<uint64_t>((uint64_t)(free_b + ctx->b_trace.cap) * 2 / 5, (uint64_t)16 << 30) / 4);
	FsArgs a;
	a.q_letters = query->letters; a.r_letters = ref->letters; a.q_limits = query->limits; a.r_limits = ref->limits;
	a.probs = d_probs; a.frame_shift = frame_shift; a.score = d_score; a.max_col = d_maxcol; a.matrix = nullptr; a.matrix_off = d_moff; a.matrix_base = 0;
	a.work = ctx->b_work.as<unsigned int>();
	std::vector<uint32_t> order;
	order.reserve(n);
	for (size_t p0 = 0; p0 < n;) {
		size_t p1 = p0 + 1;
		while (p1 < n && moff[p1 + 1] - moff[p0] <= arena_ints) ++p1;
		const uint64_t ints = moff[p1] - moff[p0];
		if (trace) {
			if (ints > arena_ints && ints * 4 > (uint64_t)(free_b + ctx->b_trace.cap) * 9 / 10) { set_error("dmnd_banded_3frame_swipe: the score matrix of one problem does not fit the device memory"); return 1; }
			if (ctx->b_trace.ensure((size_t)ints * 4 + 16)) return 1;
			DMND_CUDA_CHECK(cudaMemsetAsync(ctx->b_trace.p, 0, (size_t)ints * 4, st));  // cells the kernel never writes read as 0 (column 0, the top entry of a column, rows outside the query)
			a.matrix = ctx->b_trace.as<int32_t>(); a.matrix_base = moff[p0];
		}
		for (int c = 0; c < 5; ++c) {
			order.clear();
			for (size_t k = p0; k < p1; ++k) if (cls[k] == c) order.push_back((uint32_t)k);
			if (order.empty()) continue;
			DMND_CUDA_CHECK(cudaMemcpyAsync(d_order, order.data(), order.size() * 4, cudaMemcpyHostToDevice, st));
			DMND_CUDA_CHECK(stream_wait(ctx, st));  // `order` is reused by the next class
			DMND_CUDA_CHECK(cudaMemsetAsync(a.work, 0, sizeof(unsigned int), st));
			a.order = d_order; a.n = (uint32_t)order.size();
			const int grid = (int)std::min<size_t>((order.size() + 3) / 4, (size_t)ctx->sm_count * 8);
			if (trace) launch_fs<true>(2 << c, a, ctx->d_params, grid, st); else launch_fs<false>(2 << c, a, ctx->d_params, grid, st);
			++ctx->launches;
		}
		if (trace) {
			FsWalkArgs wa;
			wa.q_letters = a.q_letters; wa.r_letters = a.r_letters; wa.q_limits = a.q_limits; wa.r_limits = a.r_limits; wa.probs = d_probs;
			wa.n = (uint32_t)(p1 - p0); wa.pos0 = (uint32_t)p0; wa.frame_shift = frame_shift; wa.score = d_score; wa.max_col = d_maxcol;
			wa.matrix = a.matrix; wa.matrix_off = d_moff; wa.matrix_base = a.matrix_base;
			wa.res = ctx->b_results.as<dmnd_fs_result>(); wa.transcripts = d_tr; wa.transcript_off = d_toff; wa.transcript_cap = d_tcap;
			fs_walk_kernel<<<(unsigned)((p1 - p0 + 127) / 128), 128, 0, st>>>(wa, ctx->d_params);
			++ctx->launches;
		}
		if (cudaError_t le = cudaGetLastError()) { set_error(std::string("dmnd_banded_3frame_swipe: kernel launch: ") + cudaGetErrorString(le)); return 1; }
		p0 = p1;
	}
	// ---- results
	if (trace) {
		DMND_CUDA_CHECK(cudaMemcpyAsync(results, ctx->b_results.p, n * sizeof(dmnd_fs_result), cudaMemcpyDeviceToHost, st));
		if (d_tr && toff[n]) DMND_CUDA_CHECK(cudaMemcpyAsync(transcripts, d_tr, (size_t)toff[n], cudaMemcpyDeviceToHost, st));
		ctx->d2h_bytes += n * sizeof(dmnd_fs_result) + (d_tr ? toff[n] : 0);
		timer.stop();
		DMND_CUDA_CHECK(stream_wait(ctx, st));
	}
	else {
		std::vector<int32_t> sc(2 * n);  // scores, then the first column that reaches each
		DMND_CUDA_CHECK(cudaMemcpyAsync(sc.data(), d_score, 2 * n * 4, cudaMemcpyDeviceToHost, st));
		ctx->d2h_bytes += 2 * n * 4;
		timer.stop();
		DMND_CUDA_CHECK(stream_wait(ctx, st));
		for (size_t k = 0; k < n; ++k) {
			std::memset(&results[k], 0, sizeof results[k]);
			results[k].score = sc[k];
			if (sc[k] > 0) { const int i1 = std::max(problems[k].d_end - 1, 0); results[k].t_end = i1 - (problems[k].d_end - 1) + sc[n + k] + 1; }
		}
	}
	return 0;
}

}  // namespace dmnd_cuda

extern "C" int dmnd_banded_3frame_swipe(dmnd_ctx* ctx, const dmnd_block* query, const dmnd_block* ref, const dmnd_dp_problem* problems, size_t n,
                                        int frame_shift, int mode, dmnd_fs_result* results, uint8_t* transcripts, size_t transcript_cap) {
	if (!ctx || !query || !ref || (n && (!problems || !results))) { dmnd_cuda::set_error("dmnd_banded_3frame_swipe: null argument"); return 1; }
	cudaSetDevice(ctx->device);
	return dmnd_cuda::fs_impl(ctx, query, ref, problems, n, frame_shift, mode, results, transcripts, transcript_cap);
}
