// chain.cu -- dmnd_hits_chain: the host bridge between the seed stage and the DP kernels, on the device.
//
// The reference (and round 1 of this repository) turns seed hits into DP problems on the CPU, query by query:
// load_hits (align/load_hits.h:44-122) groups a query's hits by target, ungapped_stage (align/ungapped.cpp:62-118) filters the
// x-drop segments, Chaining::run (chaining/greedy_align.cpp) links them, add_dp_targets (align/gapped_score.cpp:107-180) merges
// the chains' bands.  Here the hits never leave the device for that:
//   xdrop_kernel (seed_kernels.cuh)    segment + (target, j) of every hit
//   device radix sort                  hits ordered by (query, target)
//   chain_pair_kernel                  one thread per (query, target) pair: chain_kernels.cuh -> <= CH_PROBS bands
//   chain_query_kernel / emit          per query: target count, problem count, "host" flag; problems compacted in (query, target) order
// What comes back to the host is 16 bytes per query with hits and the DP problem list (which the DP call reads anyway);
// the hits / segments of the queries flagged for the host path (more than 64 targets = ranking chunks, or a pair beyond the
// fixed capacities of chain_kernels.cuh) come back as well and take the CPU code of host/pipeline.cpp unchanged.
#include "ctx.cuh"
#include "chain_kernels.cuh"
#include <cub/cub.cuh>
#include <algorithm>

namespace dmnd_cuda {

// sort key: query and target in the fewest bits (the radix sort runs over tbits + qbits bits); keys[] afterwards = query << 32 | target
__global__ void chain_key_kernel(const dmnd_hit* __restrict__ hits, const dmnd_hit_site* __restrict__ sites, size_t n, int tbits, uint64_t* keys, uint32_t* idx) {
	const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	keys[k] = ((uint64_t)hits[k].query << tbits) | (uint64_t)sites[k].target;
	idx[k] = (uint32_t)k;
}
__global__ void chain_unpack_kernel(uint64_t* keys, size_t n, int tbits) {
	const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const uint64_t v = keys[k];
	keys[k] = ((v >> tbits) << 32) | (v & (((uint64_t)1 << tbits) - 1));
}
__global__ void chain_head_kernel(const uint64_t* __restrict__ keys, size_t n, uint8_t* pair_head, uint8_t* query_head) {
	const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const bool ph = k == 0 || keys[k] != keys[k - 1];
	pair_head[k] = ph ? 1 : 0;
	query_head[k] = (k == 0 || (keys[k] >> 32) != (keys[k - 1] >> 32)) ? 1 : 0;
}

struct PairOut { int32_t d0[CH_PROBS], d1[CH_PROBS]; int32_t np; };  // np = -1: capacity exceeded

__global__ void __launch_bounds__(64) chain_pair_kernel(const int8_t* __restrict__ q_letters, const int64_t* __restrict__ q_limits, const int8_t* __restrict__ r_letters, const int64_t* __restrict__ r_limits,
                                                        const dmnd_hit* __restrict__ hits, const dmnd_segment* __restrict__ segs, const dmnd_hit_site* __restrict__ sites,
                                                        const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ pair_start, uint32_t npairs, uint32_t nhits,
                                                        const DevParams* __restrict__ P, int band_slow, PairOut* out) {
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= npairs) return;
	const uint32_t b = pair_start[g], e = g + 1 < npairs ? pair_start[g + 1] : nhits;
	PairOut po;
	po.np = -1;
	const uint32_t nh = e - b;
	if (nh <= (uint32_t)CH_HITS) {
		const uint64_t key = keys[b];
		const uint32_t query = (uint32_t)(key >> 32), target = (uint32_t)key;
		Chainer C;
		C.score = P->score;
		C.query = q_letters + q_limits[query]; C.subject = r_letters + r_limits[target];
		C.qlen = (int)(q_limits[query + 1] - q_limits[query] - 1); C.slen = (int)(r_limits[target + 1] - r_limits[target] - 1);
		C.gap_open = P->gap_open; C.gap_extend = P->gap_extend;
		ChHit h[CH_HITS];
		for (uint32_t k = 0; k < nh; ++k) {
			const uint32_t x = idx[b + k];
			const dmnd_segment s = segs[x];
			h[k] = ChHit{ hits[x].seed_offset, sites[x].j, ChSeg{ s.i, s.j, s.len, s.score } };
		}
		ChSeg sg[CH_HITS];
		ChNode t1[CH_NODES], t2[CH_NODES];
		ChChain ch[CH_CHAINS];
		po.np = chain_pair(C, h, (int)nh, ch_band_for(C.qlen, band_slow != 0), po.d0, po.d1, sg, t1, t2, ch);
	}
	out[g] = po;
}

// one thread per query with hits: counts, host flag
__global__ void chain_query_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ pair_start, const uint32_t* __restrict__ qpair_start /* first pair of query */,
                                   uint32_t nqueries, uint32_t npairs, uint32_t nhits, const PairOut* __restrict__ pairs, int max_targets, dmnd_chain_query* out, uint32_t* nprob, uint32_t* nfb_hits) {
	const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
	if (qi >= nqueries) return;
	const uint32_t pb = qpair_start[qi], pe = qi + 1 < nqueries ? qpair_start[qi + 1] : npairs;
	const uint32_t hb = pair_start[pb], he = pe < npairs ? pair_start[pe] : nhits;
	uint32_t np = 0;
	bool host = (pe - pb) > (uint32_t)max_targets;
	for (uint32_t g = pb; g < pe; ++g) {
		if (pairs[g].np < 0) host = true; else np += (uint32_t)pairs[g].np;
	}
	dmnd_chain_query q;
	q.query = (uint32_t)(keys[hb] >> 32);
	q.n_targets = pe - pb;
	q.n_problems = host ? 0u : np;
	q.first = 0;
	q.n_hits = he - hb;
	q.flags = host ? DMND_CHAIN_HOST : 0u;
	out[qi] = q;
	nprob[qi] = host ? 0u : np;
	nfb_hits[qi] = host ? he - hb : 0u;
}
// second pass: problems of the device-chained queries at their offsets, hits of the host-path queries gathered in sorted order
__global__ void chain_emit_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ pair_start, const uint32_t* __restrict__ qpair_start,
                                  uint32_t nqueries, uint32_t npairs, uint32_t nhits, const PairOut* __restrict__ pairs, const uint32_t* __restrict__ prob_off, const uint32_t* __restrict__ fb_off,
                                  dmnd_chain_query* qout, dmnd_dp_problem* problems, const dmnd_hit* __restrict__ hits, const dmnd_segment* __restrict__ segs, const dmnd_hit_site* __restrict__ sites,
                                  dmnd_hit* fb_hits, dmnd_segment* fb_segs, dmnd_hit_site* fb_sites) {
	const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
	if (qi >= nqueries) return;
	const uint32_t pb = qpair_start[qi], pe = qi + 1 < nqueries ? qpair_start[qi + 1] : npairs;
	dmnd_chain_query q = qout[qi];
	if (q.flags & DMND_CHAIN_HOST) {
		const uint32_t hb = pair_start[pb];
		q.first = fb_off[qi];
		for (uint32_t k = 0; k < q.n_hits; ++k) {
			const uint32_t x = idx[hb + k];
			fb_hits[q.first + k] = hits[x]; fb_segs[q.first + k] = segs[x]; fb_sites[q.first + k] = sites[x];
		}
	}
	else {
		q.first = prob_off[qi];
		uint32_t o = q.first;
		for (uint32_t g = pb; g < pe; ++g) {
			const PairOut& po = pairs[g];
			const uint32_t target = (uint32_t)keys[pair_start[g]];
			for (int k = 0; k < po.np; ++k) problems[o++] = dmnd_dp_problem{ q.query, target, po.d0[k], po.d1[k] };
		}
	}
	qout[qi] = q;
}

// first PAIR of every query = rank of its first hit among the pair heads (lower bound in pair_start)
__global__ void chain_rank_kernel(const uint32_t* __restrict__ pair_start, uint32_t npairs, const uint32_t* __restrict__ qhit_start, uint32_t nq, uint32_t* out) {
	const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
	if (qi >= nq) return;
	const uint32_t hpos = qhit_start[qi];
	uint32_t lo = 0, hi = npairs;
	while (lo < hi) { const uint32_t mid = (lo + hi) / 2; if (pair_start[mid] < hpos) lo = mid + 1; else hi = mid; }
	out[qi] = lo;
}
__global__ void chain_total_kernel(const uint32_t* a, const uint32_t* a_off, const uint32_t* b, const uint32_t* b_off, uint32_t n, uint32_t* out) {
	out[0] = n ? a_off[n - 1] + a[n - 1] : 0u;
	out[1] = n ? b_off[n - 1] + b[n - 1] : 0u;
}

int hits_chain_impl(dmnd_ctx* ctx, const dmnd_block* query, const dmnd_block* ref, const dmnd_hits* h, int raw_xdrop, int band_slow, int max_targets, dmnd_chain_out* out) {
	std::memset(out, 0, sizeof *out);
	const size_t n = h->n;
	if (n == 0) return 0;
	if (n > 0xfffffff0ull) { set_error("dmnd_hits_chain: too many hits in one call"); return 1; }
	cudaStream_t st = ctx->stream;
	PhaseTimer t(ctx, PH_SEED);
	// ---- x-drop segment and (target, j) of every hit
	if (ctx->b_pairs.ensure(n * (sizeof(dmnd_segment) + sizeof(dmnd_hit_site)))) return 1;
	dmnd_segment* d_segs = ctx->b_pairs.as<dmnd_segment>();
	dmnd_hit_site* d_sites = reinterpret_cast<dmnd_hit_site*>(d_segs + n);
	if (launch_xdrop(ctx, query, ref, h, raw_xdrop, d_segs, d_sites)) return 1;
	// ---- order by (query, target)
	int qbits = 1, tbits = 1;
	while (((uint64_t)1 << qbits) < (uint64_t)query->nseq) ++qbits;
	while (((uint64_t)1 << tbits) < (uint64_t)ref->nseq) ++tbits;
	size_t tmp_sort = 0, tmp_sel = 0, tmp_scan = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 64, st);
	cub::DeviceSelect::Flagged(nullptr, tmp_sel, cub::CountingInputIterator<uint32_t>(0), (const uint8_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, n, st);
	cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, (const uint32_t*)nullptr, (uint32_t*)nullptr, n, st);
	if (ctx->b_cub.ensure(std::max(std::max(tmp_sort, tmp_sel), tmp_scan)) || ctx->b_keys.ensure(n * 8) || ctx->b_keys2.ensure(n * 8) || ctx->b_vals.ensure(n * 4) || ctx->b_vals2.ensure(n * 4)
	    || ctx->b_entries.ensure(n * 2 + n * 4 * 2 + 64) || ctx->b_counters.ensure(64))
		return 1;
	uint64_t *d_k0 = ctx->b_keys.as<uint64_t>(), *d_keys = ctx->b_keys2.as<uint64_t>();
	uint32_t *d_i0 = ctx->b_vals.as<uint32_t>(), *d_idx = ctx->b_vals2.as<uint32_t>();
	uint8_t* d_phead = ctx->b_entries.as<uint8_t>();
	uint8_t* d_qhead = d_phead + n;
	uint32_t* d_pair_start = reinterpret_cast<uint32_t*>(d_qhead + n + ((8 - (2 * n) % 8) % 8));
	uint32_t* d_qhit_start = d_pair_start + n;  // first hit of every query with hits
	uint32_t* d_cnt = ctx->b_counters.as<uint32_t>();
	const unsigned nb = (unsigned)((n + 255) / 256);
	chain_key_kernel<<<nb, 256, 0, st>>>(h->d, d_sites, n, tbits, d_k0, d_i0);
	DMND_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ctx->b_cub.p, tmp_sort, d_k0, d_keys, d_i0, d_idx, n, 0, tbits + qbits, st));
	chain_unpack_kernel<<<nb, 256, 0, st>>>(d_keys, n, tbits);
	chain_head_kernel<<<nb, 256, 0, st>>>(d_keys, n, d_phead, d_qhead);
	DMND_CUDA_CHECK(cub::DeviceSelect::Flagged(ctx->b_cub.p, tmp_sel, cub::CountingInputIterator<uint32_t>(0), d_phead, d_pair_start, d_cnt, n, st));
	DMND_CUDA_CHECK(cub::DeviceSelect::Flagged(ctx->b_cub.p, tmp_sel, cub::CountingInputIterator<uint32_t>(0), d_qhead, d_qhit_start, d_cnt + 1, n, st));
	ctx->launches += 9;
	DMND_CUDA_CHECK(cudaMemcpyAsync(ctx->h_pinned, d_cnt, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	DMND_CUDA_CHECK(stream_wait(ctx, st));
	const uint32_t npairs = ((uint32_t*)ctx->h_pinned)[0], nqueries = ((uint32_t*)ctx->h_pinned)[1];
	// ---- one thread per (query, target) pair
	if (ctx->b_hits2.ensure((size_t)npairs * sizeof(PairOut) + (size_t)nqueries * (sizeof(dmnd_chain_query) + 5 * sizeof(uint32_t)) + 256)) return 1;
	PairOut* d_pairs = ctx->b_hits2.as<PairOut>();
	dmnd_chain_query* d_q = reinterpret_cast<dmnd_chain_query*>(d_pairs + npairs);
	uint32_t* d_qpair_start = reinterpret_cast<uint32_t*>(d_q + nqueries);  // first PAIR of every query = rank of its first hit among the pair heads
	uint32_t *d_nprob = d_qpair_start + nqueries, *d_nfb = d_nprob + nqueries, *d_prob_off = d_nfb + nqueries, *d_fb_off = d_prob_off + nqueries;
	chain_pair_kernel<<<(npairs + 63) / 64, 64, 0, st>>>(query->letters, query->limits, ref->letters, ref->limits, h->d, d_segs, d_sites, d_keys, d_idx, d_pair_start, npairs, (uint32_t)n,
		ctx->d_params, band_slow, d_pairs);
	chain_rank_kernel<<<(nqueries + 255) / 256, 256, 0, st>>>(d_pair_start, npairs, d_qhit_start, nqueries, d_qpair_start);
	chain_query_kernel<<<(nqueries + 255) / 256, 256, 0, st>>>(d_keys, d_pair_start, d_qpair_start, nqueries, npairs, (uint32_t)n, d_pairs, max_targets, d_q, d_nprob, d_nfb);
	DMND_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ctx->b_cub.p, tmp_scan, d_nprob, d_prob_off, nqueries, st));
	DMND_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ctx->b_cub.p, tmp_scan, d_nfb, d_fb_off, nqueries, st));
	// totals = last offset + last count
	chain_total_kernel<<<1, 1, 0, st>>>(d_nprob, d_prob_off, d_nfb, d_fb_off, nqueries, d_cnt + 2);
	DMND_CUDA_CHECK(cudaMemcpyAsync(ctx->h_pinned, d_cnt + 2, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	DMND_CUDA_CHECK(stream_wait(ctx, st));
	const uint32_t nproblems = ((uint32_t*)ctx->h_pinned)[0], nfb = ((uint32_t*)ctx->h_pinned)[1];
	if (ctx->b_chain_probs.ensure((size_t)nproblems * sizeof(dmnd_dp_problem) + 64)
	    || ctx->b_hits.ensure((size_t)nfb * (sizeof(dmnd_hit) + sizeof(dmnd_segment) + sizeof(dmnd_hit_site)) + 64))
		return 1;
	dmnd_dp_problem* d_probs = ctx->b_chain_probs.as<dmnd_dp_problem>();
	dmnd_hit* d_fbh = ctx->b_hits.as<dmnd_hit>();
	dmnd_segment* d_fbs = reinterpret_cast<dmnd_segment*>(d_fbh + nfb);
	dmnd_hit_site* d_fbt = reinterpret_cast<dmnd_hit_site*>(d_fbs + nfb);
	chain_emit_kernel<<<(nqueries + 127) / 128, 128, 0, st>>>(d_keys, d_idx, d_pair_start, d_qpair_start, nqueries, npairs, (uint32_t)n, d_pairs, d_prob_off, d_fb_off, d_q, d_probs,
		h->d, d_segs, d_sites, d_fbh, d_fbs, d_fbt);
	ctx->launches += 7;
	DMND_CUDA_CHECK(cudaGetLastError());
	t.stop();
	out->n_queries = nqueries; out->n_pairs = npairs; out->n_problems = nproblems; out->n_host_hits = nfb;
	ctx->chain_q = d_q; ctx->chain_probs = d_probs; ctx->chain_fb_hits = d_fbh; ctx->chain_fb_segs = d_fbs; ctx->chain_fb_sites = d_fbt;
	ctx->chain_counts[0] = nqueries; ctx->chain_counts[1] = nproblems; ctx->chain_counts[2] = nfb;
	return 0;
}

int hits_chain_fetch_impl(dmnd_ctx* ctx, dmnd_chain_query* queries, dmnd_dp_problem* problems, dmnd_hit* hits, dmnd_segment* segs, dmnd_hit_site* sites) {
	cudaStream_t st = ctx->stream;
	PhaseTimer t(ctx, PH_D2H);
	const size_t nq = ctx->chain_counts[0], np = ctx->chain_counts[1], nf = ctx->chain_counts[2];
	if (nq) DMND_CUDA_CHECK(cudaMemcpyAsync(queries, ctx->chain_q, nq * sizeof(dmnd_chain_query), cudaMemcpyDeviceToHost, st));
	if (np) DMND_CUDA_CHECK(cudaMemcpyAsync(problems, ctx->chain_probs, np * sizeof(dmnd_dp_problem), cudaMemcpyDeviceToHost, st));
	if (nf) {
		DMND_CUDA_CHECK(cudaMemcpyAsync(hits, ctx->chain_fb_hits, nf * sizeof(dmnd_hit), cudaMemcpyDeviceToHost, st));
		DMND_CUDA_CHECK(cudaMemcpyAsync(segs, ctx->chain_fb_segs, nf * sizeof(dmnd_segment), cudaMemcpyDeviceToHost, st));
		DMND_CUDA_CHECK(cudaMemcpyAsync(sites, ctx->chain_fb_sites, nf * sizeof(dmnd_hit_site), cudaMemcpyDeviceToHost, st));
	}
	t.stop();
	ctx->d2h_bytes += nq * sizeof(dmnd_chain_query) + np * sizeof(dmnd_dp_problem) + nf * (sizeof(dmnd_hit) + sizeof(dmnd_segment) + sizeof(dmnd_hit_site));
	return 0;
}

}  // namespace dmnd_cuda
