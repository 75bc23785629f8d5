// seed.cu -- stages 0-2 of the double-indexed seed search (Search::search_shape, search/stage0.cpp:101-228) for sm_100a.
//
// The reference materialises both seed arrays per index chunk, radix-partitions them and hash-joins each partition
// on the CPU.  Here only the REFERENCE side is indexed (reduced-alphabet seed -> bijective 40-bit mix -> device radix
// sort -> 2^24-entry bucket directory that stays L2 resident); every QUERY position probes that directory directly, so
// the 10x larger query side is never sorted or written out.  A probe that finds a shared key becomes a compact
// "entry" (query loc, run begin, run length, seed partition); the rest of the stage works on entries only:
//   per index chunk, in the reference's order (search/stage0.cpp:109-121: masking of ALL keys of the chunk first, then
//   the search): entropy masking (seed_complexity.cpp:37-51,77-127) -> pair expansion -> 48-byte Hamming fingerprint
//   (hamming/finger_print.h:180-215, kernel.h:29-50) -> left-most filter (stage2.h:73-154, left_most.h:30-110,
//   sse_dist.h:105-190 SSE branch, pattern_matcher.h:23-65) -> Hit append (hit.h:30-48).
// Chunk order matters and is reproduced: SEED_MASK bits set while processing chunk k are visible to the left-most
// filter of chunks >= k, and verify_hit compares seed partitions against the current chunk's range (left_most.h:32-41).
#include "ctx.cuh"
#include <cmath>
#include "seed_kernels.cuh"
#include "gf_kernels.cuh"
#include <cub/cub.cuh>
#include <algorithm>



namespace dmnd_cuda {

int hits_xdrop_impl(dmnd_ctx* ctx, const dmnd_block* query, const dmnd_block* ref, const dmnd_hits* h, int raw_xdrop, dmnd_segment* host, dmnd_hit_site* sites, size_t cap) {
	if (cap < h->n) { set_error("dmnd_hits_xdrop: buffer too small"); return 1; }
	if (h->n == 0) return 0;
	if (ctx->b_pairs.ensure(h->n * (sizeof(dmnd_segment) + sizeof(dmnd_hit_site)))) return 1;
	dmnd_hit_site* d_sites = sites ? reinterpret_cast<dmnd_hit_site*>(ctx->b_pairs.as<dmnd_segment>() + h->n) : nullptr;
	PhaseTimer t(ctx, PH_SEED);
	xdrop_kernel<<<(unsigned)((h->n + 127) / 128), 128, 0, ctx->stream>>>(query->letters, query->bias, query->limits, ref->letters, ref->limits, ref->nseq,
		h->d, h->n, ctx->d_params, raw_xdrop, ctx->b_pairs.as<dmnd_segment>(), d_sites);
	++ctx->launches;
	DMND_CUDA_CHECK(cudaGetLastError());
	DMND_CUDA_CHECK(cudaMemcpyAsync(host, ctx->b_pairs.p, h->n * sizeof(dmnd_segment), cudaMemcpyDeviceToHost, ctx->stream));
	if (sites) DMND_CUDA_CHECK(cudaMemcpyAsync(sites, d_sites, h->n * sizeof(dmnd_hit_site), cudaMemcpyDeviceToHost, ctx->stream));
	t.stop();
	DMND_CUDA_CHECK(stream_wait(ctx, ctx->stream));  // the caller's buffers are complete on return whether or not the timer synchronises
	ctx->d2h_bytes += h->n * (sizeof(dmnd_segment) + (sites ? sizeof(dmnd_hit_site) : 0));
	return 0;
}

// xdrop_kernel on the context's stream, results left on the device (dmnd_hits_chain, chain.cu)
int launch_xdrop(dmnd_ctx* ctx, const dmnd_block* query, const dmnd_block* ref, const dmnd_hits* h, int raw_xdrop, dmnd_segment* d_segs, dmnd_hit_site* d_sites) {
	xdrop_kernel<<<(unsigned)((h->n + 127) / 128), 128, 0, ctx->stream>>>(query->letters, query->bias, query->limits, ref->letters, ref->limits, ref->nseq,
		h->d, h->n, ctx->d_params, raw_xdrop, d_segs, d_sites);
	++ctx->launches;
	DMND_CUDA_CHECK(cudaGetLastError());
	return 0;
}

int hits_gapped_filter_impl(dmnd_ctx* ctx, const dmnd_block* query, const dmnd_block* ref, const dmnd_hits* h, uint8_t* pass, size_t cap) {
	if (cap < h->n) { set_error("dmnd_hits_gapped_filter: buffer too small"); return 1; }
	if (ctx->params.gapped_filter_evalue <= 0.0) { set_error("dmnd_hits_gapped_filter: this sensitivity mode has no gapped filter"); return 1; }
	if (h->n == 0) return 0;
	if (ctx->b_pairs.ensure(h->n + 64)) return 1;
	PhaseTimer t(ctx, PH_SEED);
	gapped_filter_kernel<<<(unsigned)((h->n + 3) / 4), 128, 0, ctx->stream>>>(query->letters, query->bias, query->limits, ref->letters, ref->limits, ref->nseq,
		h->d, h->n, ctx->d_params, ctx->b_pairs.as<uint8_t>());
	++ctx->launches;
	DMND_CUDA_CHECK(cudaGetLastError());
	DMND_CUDA_CHECK(cudaMemcpyAsync(pass, ctx->b_pairs.p, h->n, cudaMemcpyDeviceToHost, ctx->stream));
	t.stop();
	DMND_CUDA_CHECK(stream_wait(ctx, ctx->stream));
	ctx->d2h_bytes += h->n;
	return 0;
}

static ShapeArg shape_arg(const dmnd_params& hp, int sid) {
	ShapeArg s;
	for (int k = 0; k < DMND_MAX_WEIGHT; ++k) s.pos[k] = (int8_t)hp.shape_pos[sid][k];
	s.weight = hp.shape_weight; s.span = hp.shape_len[sid]; s.rsize = hp.reduction_size; s.seedp_bits = hp.seedp_bits;
	return s;
}

static int fetch_u64(dmnd_ctx* ctx, const unsigned long long* d, unsigned long long* h) {
	DMND_CUDA_CHECK(cudaMemcpyAsync(ctx->h_pinned, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
	DMND_CUDA_CHECK(stream_wait(ctx, ctx->stream));
	*h = *(unsigned long long*)ctx->h_pinned;
	return 0;
}

int build_ref_index(dmnd_ctx* ctx, const dmnd_block* ref, int sid, RefIndex& ix) {
	cudaStream_t st = ctx->stream;
	const DevParams* P = ctx->d_params;
	const int seed_bits = ctx->h_dev_params.seed_bits;  // 40 for 10^12
	const int bucket_bits = std::min(24, seed_bits), shift = 40 - bucket_bits;  // keys are 40-bit mixes
	const size_t nbuckets = (size_t)1 << bucket_bits;
	const size_t rpos = ref->raw_len - 2 * DMND_PERIMETER_PADDING;
	ix.valid = false;
	if (ctx->b_counters.ensure(16 * sizeof(unsigned long long))) return 1;
	unsigned long long* d_cnt = ctx->b_counters.as<unsigned long long>();
	DMND_CUDA_CHECK(cudaMemsetAsync(d_cnt + 4, 0, sizeof(unsigned long long), st));
	if (ctx->b_keys.ensure(rpos * 8) || ix.keys.ensure(rpos * 8) || ctx->b_vals.ensure(rpos * 4) || ix.locs.ensure(rpos * 4)
	    || ix.bucket.ensure((nbuckets + 1) * 4 * 2))
		return 1;
	ref_enum_kernel<<<(unsigned)((rpos + 255) / 256), 256, 0, st>>>(ref->letters, ref->raw_len, P, sid, ref->has_soft ? ref->soft : nullptr, ctx->b_keys.as<uint64_t>(), ctx->b_vals.as<uint32_t>(), d_cnt + 4);
	++ctx->launches;
	unsigned long long nref = 0;
	DMND_CUDA_CHECK(cudaMemcpyAsync(ctx->h_pinned, d_cnt + 4, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	DMND_CUDA_CHECK(stream_wait(ctx, st));
	nref = *(unsigned long long*)ctx->h_pinned;
	uint64_t* d_keys = ix.keys.as<uint64_t>();
	uint32_t* d_locs = ix.locs.as<uint32_t>();
	uint32_t* d_bucket = ix.bucket.as<uint32_t>();
	uint32_t* d_hist = d_bucket + nbuckets + 1;
	{
		size_t tmp = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->b_keys.as<uint64_t>(), d_keys, ctx->b_vals.as<uint32_t>(), d_locs, (size_t)nref, 0, 40, st);
		size_t tmp2 = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, tmp2, d_hist, d_bucket, nbuckets + 1, st);
		size_t tmp3 = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, tmp3, ctx->b_vals.as<uint32_t>(), d_locs, ctx->b_keys.as<uint64_t>(), d_keys, (size_t)nref, 0, 32, st);
		if (ctx->b_cub.ensure(std::max(std::max(tmp, tmp2), tmp3))) return 1;
		if (nref && ctx->params.ungapped_evalue != 0.0) {
			// Modes with the stage-2 ungapped window filter: the locations of a seed key must come in ASCENDING order, as in the
			// reference's seed arrays, because the order decides which survivors share a window_ungapped_best call (stage2_window_kernel).
			// ref_enum_kernel hands out slots warp by warp (unordered); the key sort below is stable, so sorting by location first
			// yields (key, location) order.  --fast never looks at the order and skips this pass.
			DMND_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ctx->b_cub.p, tmp3, ctx->b_vals.as<uint32_t>(), d_locs, ctx->b_keys.as<uint64_t>(), d_keys, (size_t)nref, 0, 32, st));
			DMND_CUDA_CHECK(cudaMemcpyAsync(ctx->b_vals.p, d_locs, (size_t)nref * 4, cudaMemcpyDeviceToDevice, st));
			DMND_CUDA_CHECK(cudaMemcpyAsync(ctx->b_keys.p, d_keys, (size_t)nref * 8, cudaMemcpyDeviceToDevice, st));
			ctx->launches += 4;
		}
		if (nref) {
			DMND_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ctx->b_cub.p, tmp, ctx->b_keys.as<uint64_t>(), d_keys, ctx->b_vals.as<uint32_t>(), d_locs, (size_t)nref, 0, 40, st));
			ctx->launches += 6;
		}
		DMND_CUDA_CHECK(cudaMemsetAsync(d_hist, 0, (nbuckets + 1) * 4, st));
		if (nref) { bucket_hist_kernel<<<(unsigned)((nref + 255) / 256), 256, 0, st>>>(d_keys, (size_t)nref, shift, d_hist); ++ctx->launches; }
		DMND_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ctx->b_cub.p, tmp2, d_hist, d_bucket, nbuckets + 1, st));
		ctx->launches += 2;
	}
	// Bloom filter: >= 12 keys' worth of 256-bit blocks per 12 keys, i.e. >= 21 bits per key (false positives < 1 %)
	uint32_t bloom_blocks = 1024;
	while ((unsigned long long)bloom_blocks * 32ull < nref && bloom_blocks < (1u << 26)) bloom_blocks <<= 1;
	if (ix.bloom.ensure((size_t)bloom_blocks * 32)) return 1;
	DMND_CUDA_CHECK(cudaMemsetAsync(ix.bloom.p, 0, (size_t)bloom_blocks * 32, st));
	// first-level bitmap: >= 4 bits per reference key
	uint64_t bitmap_bits = (uint64_t)1 << 20;
	while (bitmap_bits < 4ull * nref && bitmap_bits < ((uint64_t)1 << 31)) bitmap_bits <<= 1;
	if (ix.bitmap.ensure((size_t)(bitmap_bits / 8))) return 1;
	DMND_CUDA_CHECK(cudaMemsetAsync(ix.bitmap.p, 0, (size_t)(bitmap_bits / 8), st));
	ix.bitmap_mask = (uint32_t)(bitmap_bits - 1);
	if (nref) { bloom_build_kernel<<<(unsigned)((nref + 255) / 256), 256, 0, st>>>(d_keys, (size_t)nref, ix.bloom.as<uint32_t>(), bloom_blocks - 1, ix.bitmap.as<uint32_t>(), ix.bitmap_mask); ++ctx->launches; }
	DMND_CUDA_CHECK(stream_wait(ctx, st));  // other lanes may use the index from their own streams
	ix.nref = nref; ix.sid = sid; ix.shift = shift; ix.bloom_blocks = bloom_blocks; ix.valid = true;
	return 0;
}

// One query range [q_begin, q_end): its hits, grouped by query, are APPENDED to the context's hit arena after the `out_offset` hits of the
// ranges before it (ascending ranges keep the whole list grouped by ascending query).
struct HasPairs { __host__ __device__ uint32_t operator()(const uint64_t& v) const { return v > 0 ? 1u : 0u; } };

static int search_slice(dmnd_ctx* ctx, dmnd_block* query, const dmnd_block* ref, int sid, uint32_t q_begin, uint32_t q_end, size_t out_offset, size_t* n_out, dmnd_stage_counters* counters) {
	const dmnd_params& hp = ctx->params;
	if (query->raw_len >= 0xffffffffull || ref->raw_len >= 0xffffffffull) { set_error("dmnd_search_shape: blocks of 4 G letters or more are not supported"); return 1; }
	if (sid < 0 || sid >= hp.n_shapes) { set_error("dmnd_search_shape: bad shape id"); return 1; }
	cudaStream_t st = ctx->stream;
	const DevParams* P = ctx->d_params;
	PhaseTimer timer(ctx, PH_SEED);

	// counters: [0] seeds_hit (filled on host) [1] seed_hits [2] tm1 [3] tm3 [4] ref count [5] entry count [6] hit count [7] masked
	// ... [16 + chunk] (q,s) pairs of each chunk
	constexpr int NCNT = 16 + 64;
	if (ctx->b_counters.ensure(NCNT * sizeof(unsigned long long))) return 1;
	unsigned long long* d_cnt = ctx->b_counters.as<unsigned long long>();
	DMND_CUDA_CHECK(cudaMemsetAsync(d_cnt, 0, NCNT * sizeof(unsigned long long), st));

	// ---- reference index: the block's own (dmnd_block_build_index, shared by all lanes) or a private one built now
	RefIndex& own = ctx->own_index;
	const RefIndex* ixp = &ref->idx;
	if (!(ref->idx.valid && ref->idx.sid == sid && ref->idx.content_epoch == ref->content_epoch)) {
		// the context's private index: rebuilt unless it already holds this shape of this block (the slices of one call share it)
		if (!(own.valid && own.sid == sid && ctx->own_index_block == ref && own.content_epoch == ref->content_epoch)) {
			if (build_ref_index(ctx, ref, sid, own)) return 1;
			own.content_epoch = ref->content_epoch; ctx->own_index_block = ref;
		}
		ixp = &own;
	}
	const RefIndex& ix = *ixp;
	const unsigned long long nref = ix.nref;
	const int shift = ix.shift;
	const uint64_t* d_keys = ix.keys.as<uint64_t>();
	const uint32_t* d_locs = ix.locs.as<uint32_t>();
	const uint32_t* d_bucket = ix.bucket.as<uint32_t>();
	const uint32_t* d_bloom = ix.bloom.as<uint32_t>();
	const uint32_t bloom_blocks = ix.bloom_blocks;
	const size_t qp_begin = (size_t)query->h_limits[q_begin], qp_end = (size_t)query->h_limits[q_end], qpos = qp_end - qp_begin;

	// ---- probe every query position
	size_t ecap = std::max<size_t>(1 << 20, qpos / 8);
	unsigned long long nent = 0, pairs_bound = 0;
	for (; qpos > 0;) {
		if (ctx->b_entries.ensure(ecap * sizeof(Entry))) return 1;
		DMND_CUDA_CHECK(cudaMemsetAsync(d_cnt + 5, 0, 2 * sizeof(unsigned long long), st));
		probe_kernel<<<(unsigned)((qpos + SEED_TILE - 1) / SEED_TILE), 256, 0, st>>>(query->letters, query->has_soft ? query->soft : nullptr, qp_begin, qp_end, P, shape_arg(hp, sid), d_keys, d_bucket, shift, d_bloom, bloom_blocks - 1, getenv("DMND_NO_BITMAP") ? nullptr : ix.bitmap.as<uint32_t>(), ix.bitmap_mask, ctx->b_entries.as<Entry>(), d_cnt + 5, ecap, getenv("DMND_NO_BULK") ? 0 : 1);
		++ctx->launches;
		// entry count and the (q,s) pair bound of this pass (count + 1 == d_cnt + 6) in one round trip
		DMND_CUDA_CHECK(cudaMemcpyAsync(ctx->h_pinned, d_cnt + 5, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
		DMND_CUDA_CHECK(stream_wait(ctx, st));
		nent = ((unsigned long long*)ctx->h_pinned)[0]; pairs_bound = ((unsigned long long*)ctx->h_pinned)[1];
		if (nent <= ecap) break;
		ecap = (size_t)nent + 1024;  // rare: rerun with an exact capacity
	}
	Entry* d_entries = ctx->b_entries.as<Entry>();
	if (query->has_soft && qpos > 0) {  // before any chunk's left-most filter looks at SEED_MASK bits
		motif_seedmask_kernel<<<(unsigned)((qpos + 255) / 256), 256, 0, st>>>(query->letters, query->soft, qp_begin, qp_end, hp.shape_len[sid]);
		++ctx->launches;
	}
	DMND_CUDA_CHECK(cudaMemsetAsync(d_cnt + 6, 0, sizeof(unsigned long long), st));
	ctx->last_pairs_bound = pairs_bound;
	if (ctx->b_hits.ensure((size_t)(pairs_bound + 1) * sizeof(dmnd_hit))) return 1;
	// stage-1 survivors of one chunk: 1.6 % of the pairs on the --sensitive bench workload; the list holds 1/8 of the slice's pairs and a
	// chunk that overflows it is scored by the grid over all its pairs instead
	const unsigned long long surv_cap = pairs_bound / 8 + (1ull << 20);
	if (hp.ungapped_evalue != 0.0 && (ctx->b_keys2.ensure(((size_t)pairs_bound / 32 + 8) * 4) || ctx->b_surv.ensure((size_t)surv_cap * sizeof(Survivor)))) return 1;  // stage-1 survivor bits + list
	const size_t bm_words = ((size_t)nref + 31) / 32 + 1;
	if (ctx->b_vals.ensure(bm_words * 4)) return 1;  // b_vals (unsorted reference locs) is dead after the sort
	uint32_t* d_key_seen = ctx->b_vals.as<uint32_t>();
	DMND_CUDA_CHECK(cudaMemsetAsync(d_key_seen, 0, bm_words * 4, st));

	// ---- chunks, in the reference's order
	const uint32_t parts_total = 1u << hp.seedp_bits;
	const uint32_t nchunks = std::min<uint32_t>((uint32_t)hp.index_chunks, parts_total);
	if (nchunks > 64) { set_error("dmnd_search_shape: more than 64 index chunks are not supported"); return 1; }
	const uint32_t psize = parts_total / nchunks, prem = parts_total % nchunks;
	// per entry: pairs in this chunk, their exclusive prefix, the rank among the chunk's active entries; then the active list itself
	if (ctx->b_pairs.ensure((nent + 2) * (8 + 8 + 4 + 4 + 8) + 64)) return 1;
	uint64_t* d_pairs = ctx->b_pairs.as<uint64_t>();
	uint64_t* d_pair_off = d_pairs + nent + 1;
	uint64_t* d_act_off = d_pair_off + nent + 1;
	uint32_t* d_rank = reinterpret_cast<uint32_t*>(d_act_off + nent + 2);
	uint32_t* d_act = d_rank + nent + 2;
	uint32_t* d_nact = reinterpret_cast<uint32_t*>(d_cnt + 15);
	cub::TransformInputIterator<uint32_t, HasPairs, const uint64_t*> flag_it(d_pairs, HasPairs());
	size_t scan_tmp = 0, scan_tmp2 = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, d_pairs, d_pair_off, (size_t)nent + 1, st);
	cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp2, flag_it, d_rank, (size_t)nent + 1, st);
	if (ctx->b_cub.ensure(std::max(scan_tmp, scan_tmp2))) return 1;
	size_t hits_total = 0;
	uint64_t seed_hits_total = 0;
	for (uint32_t chunk = 0; chunk < nchunks && nent > 0; ++chunk) {
		const uint32_t bsel = std::min(chunk, prem);
		const uint32_t pb = bsel * (psize + 1) + (chunk - bsel) * psize, pe = pb + (chunk < prem ? psize + 1 : psize);
		DMND_CUDA_CHECK(cudaMemsetAsync(d_pairs + nent, 0, 8, st));
		mask_kernel<<<(unsigned)((nent + 255) / 256), 256, 0, st>>>(query->letters, P, sid, d_entries, (size_t)nent, pb, pe, d_pairs, d_key_seen, d_cnt);
		DMND_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ctx->b_cub.p, scan_tmp, d_pairs, d_pair_off, (size_t)nent + 1, st));
		DMND_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ctx->b_cub.p, scan_tmp2, flag_it, d_rank, (size_t)nent + 1, st));
		active_scatter_kernel<<<(unsigned)((nent + 1 + 255) / 256), 256, 0, st>>>(d_pairs, d_pair_off, d_rank, (size_t)nent, d_act, d_act_off, d_nact);
		ctx->launches += 6;
		// no host round trip inside the chunk loop: the chunk's pair total stays on the device (the stage kernels read it, a copy goes
		// to the counters), the grid covers the bound over all chunks and surplus CTAs leave at once
		DMND_CUDA_CHECK(cudaMemcpyAsync(d_cnt + 16 + chunk, d_pair_off + nent, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
		if (pairs_bound == 0) continue;
		LmCtx x;
		x.P = P; x.sid = sid; x.chunked = hp.index_chunks > 1; x.range_begin = pb; x.range_end = pe;
		x.cur_matcher = ctx->d_matcher[sid + 1]; x.cur_minlen = ctx->matcher_minlen[sid + 1]; x.cur_suffix = ctx->matcher_suffix[sid + 1];
		x.prev_matcher = ctx->d_matcher[sid]; x.prev_minlen = ctx->matcher_minlen[sid]; x.prev_suffix = ctx->matcher_suffix[sid];
		const PairLookup L{ d_act, d_act_off, d_nact };
		const unsigned grid = (unsigned)((pairs_bound + STAGE_CTA - 1) / STAGE_CTA);
		if (hp.ungapped_evalue == 0.0)
			stage12_kernel<<<grid, STAGE_CTA, 0, st>>>(query->letters, query->limits, query->nseq, ref->letters, d_entries, L, d_locs, x, ctx->b_hits.as<dmnd_hit>(), d_cnt + 6, d_cnt);
		else {
			DMND_CUDA_CHECK(cudaMemsetAsync(d_cnt + 9, 0, sizeof(unsigned long long), st));  // survivors of this chunk
			stage1_flags_kernel<<<grid, STAGE_CTA, 0, st>>>(query->letters, ref->letters, d_entries, L, d_locs, (unsigned)hp.hamming_id, ctx->b_keys2.as<uint32_t>(),
				ctx->b_surv.as<Survivor>(), d_cnt + 9, surv_cap, d_cnt);
			stage2_window_kernel<<<(unsigned)std::min<unsigned long long>(grid, (unsigned long long)ctx->sm_count * 16), STAGE_CTA, 0, st>>>(query->letters, query->limits, query->nseq, ref->letters, d_entries, L, d_locs,
				ctx->b_keys2.as<uint32_t>(), x, ctx->b_surv.as<Survivor>(), d_cnt + 9, surv_cap, ctx->b_hits.as<dmnd_hit>(), d_cnt + 6, d_cnt);
			stage2_window_full_kernel<<<grid, STAGE_CTA, 0, st>>>(query->letters, query->limits, query->nseq, ref->letters, d_entries, L, d_locs,
				ctx->b_keys2.as<uint32_t>(), x, d_cnt + 9, surv_cap, ctx->b_hits.as<dmnd_hit>(), d_cnt + 6, d_cnt);  // leaves at once unless the list overflowed
			++ctx->launches;
			++ctx->launches;
		}
		++ctx->launches;
	}
	{
		DMND_CUDA_CHECK(cudaMemcpyAsync(ctx->h_pinned, d_cnt, NCNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
		DMND_CUDA_CHECK(stream_wait(ctx, st));
		const unsigned long long* hcn = (const unsigned long long*)ctx->h_pinned;
		hits_total = (size_t)hcn[6];
		for (uint32_t chunk = 0; chunk < nchunks; ++chunk) seed_hits_total += hcn[16 + chunk];
	}

	// ---- group hits by query (stable order inside a query is not required)
	*n_out = hits_total;
	if (hits_total) {
		if (ctx->b_hits_out.ensure_keep((out_offset + hits_total) * sizeof(dmnd_hit), out_offset * sizeof(dmnd_hit))) return 1;
		dmnd_hit* const h_d = ctx->b_hits_out.as<dmnd_hit>() + out_offset;
		if (ctx->b_keys.ensure(hits_total * 4) || ctx->b_vals.ensure(hits_total * 4)) return 1;
		uint32_t* k_in = ctx->b_keys.as<uint32_t>();
		uint32_t* k_out = ctx->b_vals.as<uint32_t>();
		extract_query_kernel<<<(unsigned)((hits_total + 255) / 256), 256, 0, st>>>(ctx->b_hits.as<dmnd_hit>(), hits_total, k_in);
		size_t tmp = 0;
		int end_bit = 1;
		while (((uint64_t)1 << end_bit) < (uint64_t)query->nseq) ++end_bit;
		cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, ctx->b_hits.as<dmnd_hit>(), h_d, hits_total, 0, end_bit, st);
		if (ctx->b_cub.ensure(tmp)) return 1;
		DMND_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ctx->b_cub.p, tmp, k_in, k_out, ctx->b_hits.as<dmnd_hit>(), h_d, hits_total, 0, end_bit, st));
		ctx->launches += 4;
	}
	unsigned long long hc[16];
	DMND_CUDA_CHECK(cudaMemcpyAsync(ctx->h_pinned, d_cnt, sizeof hc, cudaMemcpyDeviceToHost, st));
	timer.stop();
	DMND_CUDA_CHECK(stream_wait(ctx, st));
	std::memcpy(hc, ctx->h_pinned, sizeof hc);
	if (counters) {
		counters->seeds_hit += hc[0];
		counters->seed_hits += seed_hits_total;
		counters->tentative_matches1 += hc[2];
		counters->tentative_matches2 += hp.ungapped_evalue == 0.0 ? hc[2] : hc[8];
		counters->tentative_matches3 += hits_total;
		counters->masked_seeds += hc[7];
	}
	return 0;
}

// Search::search_shape for the query range [q_begin, q_end).  Shapes of low weight match at (nearly) every query position, so the
// entry and pair lists of a whole block would take tens of GB (C4: 3*10^8 positions x 16 shapes of weight 8): the range is cut at
// sequence boundaries into slices of at most DMND_SEED_SLICE letters (default 4*10^7 for weights below 10, one slice otherwise).
// Queries are independent of each other in every stage (SEED_MASK bits and left-most windows stay inside a sequence), so the
// slices' hit lists, concatenated in range order, are the hits of the unsliced search.  "Seeds hit" / "masked seeds" count a key
// once per slice it occurs in (a diagnostic counter only; the hit-level counters are exact).
int search_shape_impl(dmnd_ctx* ctx, dmnd_block* query, const dmnd_block* ref, int sid, uint32_t q_begin, uint32_t q_end, dmnd_hits** out, dmnd_stage_counters* counters) {
	dmnd_stage_counters cn;
	std::memset(&cn, 0, sizeof cn);
	// The (q, s) pair buffers of a slice (16 bytes per pair for the hits, 2 for the survivor list) are held to ~5*10^8 pairs: the first
	// slice of a call is small (the pair density of real seed frequencies is 50-100 x what letters / 10^weight suggests: 292 pairs per
	// query letter for --sensitive on 1.5*10^8 reference letters), every further slice is sized from the density measured so far
	const bool adaptive = ctx->params.shape_weight < 10 && getenv("DMND_SEED_SLICE") == nullptr;
	const double target_pairs = 5e8;
	size_t slice = adaptive ? (size_t)2000000 : (size_t)-1;
	double letters_done = 0.0, pairs_done = 0.0;
	if (const char* ev = getenv("DMND_SEED_SLICE")) slice = std::max<size_t>(1024, strtoull(ev, nullptr, 10));
	size_t total = 0;
	uint32_t b = q_begin;
	do {
		uint32_t e = b;
		const size_t p0 = (size_t)query->h_limits[b];
		while (e < q_end && ((size_t)query->h_limits[e + 1] - p0 <= slice || e == b)) ++e;
		size_t n = 0;
		if (search_slice(ctx, query, ref, sid, b, e, total, &n, &cn)) return 1;
		total += n;
		if (adaptive) {
			letters_done += (double)((size_t)query->h_limits[e] - p0); pairs_done += (double)ctx->last_pairs_bound;
			slice = (size_t)std::min(4e7, std::max(2e5, target_pairs * letters_done / std::max(pairs_done, 1.0)));
		}
		b = e;
	} while (b < q_end);
	dmnd_hits* h = new dmnd_hits();
	h->n = total;
	h->d = total ? ctx->b_hits_out.as<dmnd_hit>() : nullptr;
	if (counters) *counters = cn;
	*out = h;
	return 0;
}


// ---- diagnostics (tools/seed_stage_diag.py): the left-most filter of ONE (query loc, reference loc) pair with its intermediates,
// evaluated on the device with the production device functions, on the block's current SEED_MASK state.
__global__ void lm_debug_kernel(const int8_t* q_letters, const int64_t* q_limits, uint32_t nq, const int8_t* r_letters, uint32_t qloc, uint32_t sloc, LmCtx x, unsigned long long* out) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	const int8_t *qp = q_letters + qloc, *sp = r_letters + sloc;
	uint32_t a = 0, b = nq;
	while (b - a > 1) { const uint32_t mid = a + (b - a) / 2; if ((uint64_t)q_limits[mid] <= (uint64_t)qloc) a = mid; else b = mid; }
	const int seed_offset0 = (int)((int64_t)qloc - q_limits[a]);
	const int window0 = x.P->ungapped_window;
	int cb, ce;
	clip(qp - window0, 2 * window0, window0, cb, ce);
	const int window_left0 = window0 - cb, window_clipped = ce - cb;
	const int8_t* qc = qp - window0 + cb;
	const int interval_mod = x.P->left_most_interval > 0 ? seed_offset0 % x.P->left_most_interval : window_left0;
	const int overhang = max(window_left0 - interval_mod, 0);
	const int8_t* query = qc + overhang; const int query_len = window_clipped - overhang; const int8_t* subject = sp - window_left0 + overhang;
	const int seed_offset = window_left0 - overhang, seed_len = x.P->shape_len[x.sid];
	out[0] = left_most_filter(x, query, query_len, subject, seed_offset, seed_len) ? 1 : 0;
	const DevParams* P = x.P;
	int d = max(seed_offset - 16, 0), window_left = min(16, seed_offset);
	const int8_t *q = query + d, *s = subject + d;
	int window = min(query_len - d, window_left + 1 + 32);
	clip(s, window, window_left, cb, ce);
	window = ce; d = cb;
	q += d; s += d; window_left -= d; window -= d;
	uint64_t match_mask = 0, seed_bits = 0;
	for (int k = 0; k < window && k < 64; ++k) {
		if (P->map8[q[k] & 31] == P->map8b[s[k] & 31]) match_mask |= (uint64_t)1 << k;
		if (q[k] & DMND_SEED_MASK) seed_bits |= (uint64_t)1 << k;
	}
	const uint64_t query_seed_mask = ~seed_bits;
	const uint32_t len_left = (uint32_t)(window_left + seed_len - 1),
		match_mask_left = (uint32_t)((((uint64_t)1 << len_left) - 1) & match_mask), query_mask_left = (uint32_t)((((uint64_t)1 << len_left) - 1) & query_seed_mask);
	const uint32_t raw_left = matcher_hit(x.cur_matcher, x.cur_minlen, x.cur_suffix, match_mask_left, len_left);
	const uint32_t left_hit = raw_left & query_mask_left;
	const uint32_t len_right = (uint32_t)(window - window_left - 1), match_mask_right = (uint32_t)(match_mask >> (window_left + 1)), query_mask_right = (uint32_t)(query_seed_mask >> (window_left + 1));
	const uint32_t right_hit = (x.chunked ? matcher_hit(x.cur_matcher, x.cur_minlen, x.cur_suffix, match_mask_right, len_right) : matcher_hit(x.prev_matcher, x.prev_minlen, x.prev_suffix, match_mask_right, len_right)) & query_mask_right;
	out[1] = match_mask; out[2] = seed_bits; out[3] = ((uint64_t)raw_left << 32) | left_hit; out[4] = right_hit;
	out[5] = ((uint64_t)(uint32_t)seed_offset << 32) | (uint32_t)window_left; out[6] = ((uint64_t)(uint32_t)window << 32) | len_left;
	out[7] = left_hit ? (verify_hits(x, left_hit, q, s, true, match_mask_left) ? 1 : 0) : 2;
	out[8] = right_hit ? (verify_hits(x, right_hit, q + window_left + 1, s + window_left + 1, false, match_mask_right) ? 1 : 0) : 2;
	// every left candidate: position, verify_hit, fingerprint count, partition of the subject seed of the current shape
	int n = 0;
	for (int pos = 0; pos < 32 && n < 20; ++pos) {
		if (!((left_hit >> pos) & 1u)) continue;
		const uint32_t mm = match_mask_left >> pos;
		uint64_t seed = 0; int valid = 1;
		for (int k = 0; k < P->shape_weight; ++k) { const int l = s[pos + P->shape_pos[x.sid][k]] & 31; if (l == 23 || l == 31 || l == 24) valid = 0; seed = seed * (uint64_t)P->reduction_size + P->reduction[l]; }
		const uint32_t part = (uint32_t)(seed & (((uint64_t)1 << P->seedp_bits) - 1));
		out[9 + n] = ((uint64_t)pos << 56) | ((uint64_t)(verify_hit(x, q + pos, s + pos, true, mm) ? 1 : 0) << 48) | ((uint64_t)(((P->shape_mask[x.sid] & mm) == P->shape_mask[x.sid]) ? 1 : 0) << 40)
			| ((uint64_t)valid << 36) | ((uint64_t)fingerprint_match(q + pos, s + pos) << 24) | part;
		++n;
	}
	out[29] = (uint64_t)n;
}

int debug_left_most_impl(dmnd_ctx* ctx, const dmnd_block* query, const dmnd_block* ref, int sid, int chunk, uint32_t qloc, uint32_t sloc, unsigned long long* out30) {
	const dmnd_params& hp = ctx->params;
	const uint32_t parts_total = 1u << hp.seedp_bits;
	const uint32_t nchunks = std::min<uint32_t>((uint32_t)hp.index_chunks, parts_total);
	const uint32_t psize = parts_total / nchunks, prem = parts_total % nchunks;
	const uint32_t bsel = std::min((uint32_t)chunk, prem);
	const uint32_t pb = bsel * (psize + 1) + ((uint32_t)chunk - bsel) * psize, pe = pb + ((uint32_t)chunk < prem ? psize + 1 : psize);
	LmCtx x;
	x.P = ctx->d_params; x.sid = sid; x.chunked = hp.index_chunks > 1; x.range_begin = pb; x.range_end = pe;
	x.cur_matcher = ctx->d_matcher[sid + 1]; x.cur_minlen = ctx->matcher_minlen[sid + 1]; x.cur_suffix = ctx->matcher_suffix[sid + 1];
	x.prev_matcher = ctx->d_matcher[sid]; x.prev_minlen = ctx->matcher_minlen[sid]; x.prev_suffix = ctx->matcher_suffix[sid];
	if (ctx->b_counters.ensure(128 * sizeof(unsigned long long))) return 1;
	unsigned long long* d = ctx->b_counters.as<unsigned long long>();
	DMND_CUDA_CHECK(cudaMemsetAsync(d, 0, 30 * 8, ctx->stream));
	lm_debug_kernel<<<1, 32, 0, ctx->stream>>>(query->letters, query->limits, query->nseq, ref->letters, qloc, sloc, x, d);
	DMND_CUDA_CHECK(cudaGetLastError());
	DMND_CUDA_CHECK(cudaMemcpyAsync(out30, d, 30 * 8, cudaMemcpyDeviceToHost, ctx->stream));
	DMND_CUDA_CHECK(stream_wait(ctx, ctx->stream));
	return 0;
}

}  // namespace dmnd_cuda
