// emu_seed.cpp -- the seed stage's kernels (diamond_b200/csrc/cuda/seed_kernels.cuh, the source the GPU library is built from)
// on the CPU behind tests/emu_cuda.h, driven by a host sequence that mirrors build_ref_index + search_shape_impl (seed.cu) with
// the CUB sorts / scans replaced by their std:: equivalents, compared with the oracle's dmnd_search_shape: hits (incl. the
// ungapped window scores of the modes that have the stage-2 filter), stage counters, SEED_MASK bits -- for every shape of the
// mode, in order, on the same blocks (bits set by one shape stay for the next).
// usage: emu_seed DIR SENSITIVITY MASKING     (DIR holds q.i8 q.i64 r.i8 r.i64)
#include "emu_cuda.h"
#include "../diamond_b200/csrc/cuda/seed_kernels.cuh"
#include <algorithm>
#include <numeric>
#include <string>
using namespace dmnd_cuda;
extern "C" int dmnd_oracle_block_soft(const dmnd_block* b, uint8_t* out, size_t raw_len);

template<typename T> static std::vector<T> slurp(const std::string& path) {
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
	fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
	std::vector<T> v((size_t)n / sizeof(T));
	if (fread(v.data(), 1, (size_t)n, f) != (size_t)n) exit(2);
	fclose(f);
	return v;
}

struct Blk { std::vector<int8_t> letters; std::vector<int64_t> limits; std::vector<uint32_t> soft; bool has_soft = false; size_t raw_len; uint32_t nseq; };

static void fill_dev_params(const dmnd_params& hp, DevParams& d) {
	memset(&d, 0, sizeof d);
	memcpy(d.score, hp.score, 1024); memcpy(d.reduction, hp.reduction, 32); memcpy(d.map8, hp.map8, 32); memcpy(d.map8b, hp.map8b, 32);
	memcpy(d.shape_pos, hp.shape_pos, sizeof d.shape_pos); memcpy(d.shape_mask, hp.shape_mask, sizeof d.shape_mask); memcpy(d.shape_len, hp.shape_len, sizeof d.shape_len);
	d.n_shapes = hp.n_shapes; d.shape_weight = hp.shape_weight; d.reduction_size = hp.reduction_size; d.hamming_id = hp.hamming_id; d.seedp_bits = hp.seedp_bits;
	d.index_chunks = hp.index_chunks; d.left_most_interval = hp.left_most_interval; d.ungapped_window = hp.ungapped_window; d.gap_open = hp.gap_open; d.gap_extend = hp.gap_extend;
	d.seed_cut = hp.seed_cut;
	memcpy(d.ungapped_cutoff, hp.ungapped_cutoff, sizeof d.ungapped_cutoff); d.short_query_ungapped_cutoff = hp.short_query_ungapped_cutoff; d.short_query_max_len = hp.short_query_max_len;
	d.query_contexts = hp.query_contexts > 1 ? hp.query_contexts : 1; memcpy(d.ungapped_cutoff_short, hp.ungapped_cutoff_short, sizeof d.ungapped_cutoff_short);
	unsigned long long pw = 1; for (int i = 0; i < hp.shape_weight; ++i) pw *= (unsigned long long)hp.reduction_size;
	int bits = 0; for (unsigned long long x = pw - 1; x > 0; x >>= 1) ++bits;
	d.seed_bits = bits;
	static const double LNFACT[13] = { 0.000000, 0.000000, 0.693147, 1.791759, 3.178054, 4.787492, 6.579251, 8.525161, 10.604603, 12.801827, 15.104413, 17.502308, 19.987214 };
	memcpy(d.lnfact, LNFACT, sizeof LNFACT);
}

struct Matchers { std::vector<std::vector<uint8_t>> t; std::vector<uint32_t> minlen, suffix; };
static Matchers build_matchers(const dmnd_params& hp) {  // dmnd_create (ctx.cu): PatternMatcher tables over shapes [0, k)
	Matchers m; m.t.resize(hp.n_shapes + 1); m.minlen.resize(hp.n_shapes + 1); m.suffix.resize(hp.n_shapes + 1);
	for (int k = 0; k <= hp.n_shapes; ++k) {
		uint32_t minl = 32, maxl = 0;
		for (int i = 0; i < k; ++i) { const uint32_t len = 32 - (uint32_t)__builtin_clz(hp.shape_mask[i]); maxl = std::max(maxl, len); minl = std::min(minl, len); }
		m.minlen[k] = minl; m.suffix[k] = (1u << maxl) - 1;
		m.t[k].assign((size_t)m.suffix[k] + 1, 0);
		for (uint32_t s = 0; s <= m.suffix[k]; ++s) for (int i = 0; i < k; ++i) if ((s & hp.shape_mask[i]) == hp.shape_mask[i]) m.t[k][s] = 1;
	}
	return m;
}

struct Index { std::vector<uint64_t> keys; std::vector<uint32_t> locs, bucket, bloom, bitmap; int shift; uint32_t bloom_blocks, bitmap_mask; unsigned long long nref; };

// build_ref_index (seed.cu)
static void build_index(const Blk& ref, const DevParams& P, const dmnd_params& hp, int sid, Index& ix) {
	const int bucket_bits = std::min(24, P.seed_bits), shift = 40 - bucket_bits;
	const size_t nbuckets = (size_t)1 << bucket_bits, rpos = ref.raw_len - 2 * DMND_PERIMETER_PADDING;
	std::vector<uint64_t> keys(rpos); std::vector<uint32_t> vals(rpos);
	unsigned long long nref = 0;
	emu::launch((unsigned)((rpos + 255) / 256), 256, [&] { ref_enum_kernel(ref.letters.data(), ref.raw_len, &P, sid, ref.has_soft ? ref.soft.data() : nullptr, keys.data(), vals.data(), &nref); });
	std::vector<size_t> ord((size_t)nref); std::iota(ord.begin(), ord.end(), 0);
	if (hp.ungapped_evalue != 0.0) std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return vals[a] < vals[b]; });  // the location pre-sort
	else std::reverse(ord.begin(), ord.end());  // --fast never looks at the order inside a key: make it visibly arbitrary
	std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });  // stable key sort (cub::DeviceRadixSort)
	ix.keys.resize((size_t)nref); ix.locs.resize((size_t)nref);
	for (size_t i = 0; i < (size_t)nref; ++i) { ix.keys[i] = keys[ord[i]]; ix.locs[i] = vals[ord[i]]; }
	std::vector<uint32_t> hist(nbuckets + 1, 0);
	if (nref) emu::launch((unsigned)((nref + 255) / 256), 256, [&] { bucket_hist_kernel(ix.keys.data(), (size_t)nref, shift, hist.data()); });
	ix.bucket.assign(nbuckets + 1, 0);
	for (size_t b = 0; b < nbuckets; ++b) ix.bucket[b + 1] = ix.bucket[b] + hist[b];
	uint32_t bloom_blocks = 1024;
	while ((unsigned long long)bloom_blocks * 32ull < nref && bloom_blocks < (1u << 26)) bloom_blocks <<= 1;
	ix.bloom.assign((size_t)bloom_blocks * 8, 0);
	uint64_t bitmap_bits = (uint64_t)1 << 20;
	while (bitmap_bits < 4ull * nref && bitmap_bits < ((uint64_t)1 << 31)) bitmap_bits <<= 1;
	ix.bitmap.assign((size_t)(bitmap_bits / 32), 0); ix.bitmap_mask = (uint32_t)(bitmap_bits - 1);
	if (nref) emu::launch((unsigned)((nref + 255) / 256), 256, [&] { bloom_build_kernel(ix.keys.data(), (size_t)nref, ix.bloom.data(), bloom_blocks - 1, ix.bitmap.data(), ix.bitmap_mask); });
	ix.shift = shift; ix.bloom_blocks = bloom_blocks; ix.nref = nref;
}

// search_shape_impl (seed.cu)
static void search_shape(Blk& query, const Blk& ref, const DevParams& P, const dmnd_params& hp, const Matchers& M, int sid, std::vector<dmnd_hit>& hits, dmnd_stage_counters& cn) {
	Index ix; build_index(ref, P, hp, sid, ix);
	unsigned long long cnt[16 + 64]; memset(cnt, 0, sizeof cnt);
	const size_t qp_begin = (size_t)query.limits[0], qp_end = (size_t)query.limits[query.nseq], qpos = qp_end - qp_begin;
	ShapeArg sh; for (int k = 0; k < DMND_MAX_WEIGHT; ++k) sh.pos[k] = (int8_t)hp.shape_pos[sid][k];
	sh.weight = hp.shape_weight; sh.span = hp.shape_len[sid]; sh.rsize = hp.reduction_size; sh.seedp_bits = hp.seedp_bits;
	size_t ecap = std::max<size_t>(1 << 20, qpos / 8);
	std::vector<Entry> entries(ecap);
	emu::launch((unsigned)((qpos + SEED_TILE - 1) / SEED_TILE), 256, [&] { probe_kernel(query.letters.data(), query.has_soft ? query.soft.data() : nullptr, qp_begin, qp_end, &P, sh, ix.keys.data(),
		ix.bucket.data(), ix.shift, ix.bloom.data(), ix.bloom_blocks - 1, ix.bitmap.data(), ix.bitmap_mask, entries.data(), cnt + 5, ecap, 1); });
	const unsigned long long nent = cnt[5], pairs_bound = cnt[6];
	if (nent > ecap) { printf("FAIL entry capacity\n"); exit(1); }
	if (query.has_soft && qpos > 0) emu::launch((unsigned)((qpos + 255) / 256), 256, [&] { motif_seedmask_kernel(query.letters.data(), query.soft.data(), qp_begin, qp_end, hp.shape_len[sid]); });
	cnt[6] = 0;
	std::vector<dmnd_hit> dh((size_t)pairs_bound + 1);
	std::vector<uint32_t> key_seen(((size_t)ix.nref + 31) / 32 + 1, 0), flags((size_t)pairs_bound / 32 + 8, 0xdeadbeefu);
	const uint32_t parts_total = 1u << hp.seedp_bits, nchunks = std::min<uint32_t>((uint32_t)hp.index_chunks, parts_total);
	const uint32_t psize = parts_total / nchunks, prem = parts_total % nchunks;
	std::vector<uint64_t> pairs((size_t)nent + 1), pair_off((size_t)nent + 1), act_off((size_t)nent + 2);
	std::vector<uint32_t> rank((size_t)nent + 2), act((size_t)nent + 2);
	uint32_t nact = 0;
	uint64_t seed_hits_total = 0;
	for (uint32_t chunk = 0; chunk < nchunks && nent > 0; ++chunk) {
		const uint32_t bsel = std::min(chunk, prem);
		const uint32_t pb = bsel * (psize + 1) + (chunk - bsel) * psize, pe = pb + (chunk < prem ? psize + 1 : psize);
		pairs[(size_t)nent] = 0;
		emu::launch((unsigned)((nent + 255) / 256), 256, [&] { mask_kernel(query.letters.data(), &P, sid, entries.data(), (size_t)nent, pb, pe, pairs.data(), key_seen.data(), cnt); });
		uint64_t run = 0;
		for (size_t i = 0; i <= (size_t)nent; ++i) { pair_off[i] = run; run += pairs[i]; }  // cub::DeviceScan::ExclusiveSum
		{ uint32_t r = 0; for (size_t i = 0; i <= (size_t)nent; ++i) { rank[i] = r; r += pairs[i] > 0 ? 1u : 0u; } }  // ... and the scan of (pairs > 0)
		emu::launch((unsigned)((nent + 1 + 255) / 256), 256, [&] { active_scatter_kernel(pairs.data(), pair_off.data(), rank.data(), (size_t)nent, act.data(), act_off.data(), &nact); });
		seed_hits_total += pair_off[(size_t)nent];
		if (pairs_bound == 0) continue;
		LmCtx x;
		x.P = &P; x.sid = sid; x.chunked = hp.index_chunks > 1; x.range_begin = pb; x.range_end = pe;
		x.cur_matcher = M.t[sid + 1].data(); x.cur_minlen = M.minlen[sid + 1]; x.cur_suffix = M.suffix[sid + 1];
		x.prev_matcher = M.t[sid].data(); x.prev_minlen = M.minlen[sid]; x.prev_suffix = M.suffix[sid];
		const PairLookup L{ act.data(), act_off.data(), &nact };
		const unsigned grid = (unsigned)((pairs_bound + STAGE_CTA - 1) / STAGE_CTA);
		if (hp.ungapped_evalue == 0.0)
			emu::launch(grid, STAGE_CTA, [&] { stage12_kernel(query.letters.data(), query.limits.data(), query.nseq, ref.letters.data(), entries.data(), L, ix.locs.data(), x, dh.data(), cnt + 6, cnt); });
		else {
			// every other chunk with a list that is too small: the window pass must then come from the grid over all pairs
			const unsigned long long surv_cap = (chunk & 1) ? 40 : pairs_bound + 64;
			std::vector<Survivor> surv((size_t)surv_cap);
			cnt[9] = 0;
			emu::launch(grid, STAGE_CTA, [&] { stage1_flags_kernel(query.letters.data(), ref.letters.data(), entries.data(), L, ix.locs.data(), (unsigned)hp.hamming_id, flags.data(), surv.data(), cnt + 9, surv_cap, cnt); });
			emu::launch(std::min(grid, 7u), STAGE_CTA, [&] { stage2_window_kernel(query.letters.data(), query.limits.data(), query.nseq, ref.letters.data(), entries.data(), L, ix.locs.data(), flags.data(), x,
				surv.data(), cnt + 9, surv_cap, dh.data(), cnt + 6, cnt); });
			if (cnt[9] + 32 > surv_cap)  // (the library launches it unconditionally and the kernel leaves at once unless the list overflowed; the emulation saves the coroutines)
			emu::launch(grid, STAGE_CTA, [&] { stage2_window_full_kernel(query.letters.data(), query.limits.data(), query.nseq, ref.letters.data(), entries.data(), L, ix.locs.data(), flags.data(), x,
				cnt + 9, surv_cap, dh.data(), cnt + 6, cnt); });
		}
	}
	hits.assign(dh.begin(), dh.begin() + (ptrdiff_t)cnt[6]);
	cn.seeds_hit = cnt[0]; cn.seed_hits = seed_hits_total; cn.tentative_matches1 = cnt[2]; cn.tentative_matches2 = hp.ungapped_evalue == 0.0 ? cnt[2] : cnt[8];
	cn.tentative_matches3 = cnt[6]; cn.masked_seeds = cnt[7];
}

static bool hit_less(const dmnd_hit& a, const dmnd_hit& b) {
	if (a.query != b.query) return a.query < b.query;
	if (a.subject_score != b.subject_score) return a.subject_score < b.subject_score;
	return a.seed_offset < b.seed_offset;
}

int main(int argc, char** argv) {
	if (argc < 4) return 2;
	const std::string dir = argv[1];
	dmnd_search_opts o; dmnd_search_opts_default(&o); o.sensitivity = atoi(argv[2]);
	const int masking = atoi(argv[3]);
	if (argc > 4) o.query_contexts = atoi(argv[4]);  // 6: the query block holds translated frames (blastx)
	Blk q, r;
	q.letters = slurp<int8_t>(dir + "/q.i8"); q.limits = slurp<int64_t>(dir + "/q.i64"); r.letters = slurp<int8_t>(dir + "/r.i8"); r.limits = slurp<int64_t>(dir + "/r.i64");
	q.raw_len = q.letters.size(); r.raw_len = r.letters.size(); q.nseq = (uint32_t)q.limits.size() - 1; r.nseq = (uint32_t)r.limits.size() - 1;
	dmnd_params hp; if (dmnd_params_init(&o, &hp)) { fprintf(stderr, "%s\n", dmnd_last_error()); return 2; }
	DevParams P; fill_dev_params(hp, P);
	const Matchers M = build_matchers(hp);
	dmnd_ctx* ctx; if (dmnd_create(0, &hp, &ctx)) return 2;
	dmnd_block *qb, *rb;
	if (dmnd_block_upload(ctx, q.letters.data(), q.raw_len, q.limits.data(), q.nseq, &qb) || dmnd_block_upload(ctx, r.letters.data(), r.raw_len, r.limits.data(), r.nseq, &rb)) return 2;
	if (masking) {  // the blocks as dmnd_block_mask leaves them (that path has its own emulation test): letters from the oracle, soft bits from its table
		uint64_t n;
		if (dmnd_block_mask(ctx, qb, 5, 0, q.nseq, &n) || dmnd_block_mask(ctx, rb, 5, 0, r.nseq, &n)) return 2;
		for (Blk* b : { &q, &r }) {
			dmnd_block* ob = b == &q ? qb : rb;
			dmnd_block_download_letters(ctx, ob, b->letters.data(), b->raw_len);
			std::vector<uint8_t> soft(b->raw_len);
			dmnd_oracle_block_soft(ob, soft.data(), soft.size());
			b->soft.assign(b->raw_len / 32 + 8, 0);
			for (size_t p = 0; p < b->raw_len; ++p) if (soft[p]) b->soft[p >> 5] |= 1u << (p & 31);
			b->has_soft = true;
		}
	}
	q.letters.resize(((q.raw_len + 63) & ~(size_t)63) + 64 + SEED_TILE + 64, (int8_t)DMND_DELIMITER);  // the device allocation's slack (+ what the last probe tile reads)
	r.letters.resize(((r.raw_len + 63) & ~(size_t)63) + 64, (int8_t)DMND_DELIMITER);
	if (q.soft.size()) q.soft.resize(q.letters.size() / 32 + 8, 0);
	int fails = 0; size_t total_hits = 0;
	for (int sid = 0; sid < hp.n_shapes; ++sid) {
		dmnd_hits* h; dmnd_stage_counters want;
		if (dmnd_search_shape(ctx, qb, rb, sid, &h, &want)) { fprintf(stderr, "%s\n", dmnd_last_error()); return 2; }
		std::vector<dmnd_hit> oh(dmnd_hits_count(h));
		if (!oh.empty()) dmnd_hits_download(ctx, h, oh.data(), oh.size());
		dmnd_hits_free(ctx, h);
		std::vector<int8_t> olet(q.raw_len); dmnd_block_download_letters(ctx, qb, olet.data(), olet.size());
		std::vector<dmnd_hit> eh; dmnd_stage_counters got; memset(&got, 0, sizeof got);
		search_shape(q, r, P, hp, M, sid, eh, got);
		std::sort(oh.begin(), oh.end(), hit_less); std::sort(eh.begin(), eh.end(), hit_less);
		total_hits += oh.size();
		if (memcmp(&want, &got, sizeof want) != 0) { ++fails; printf("FAIL shape %d counters: seeds_hit %llu/%llu seed_hits %llu/%llu tm1 %llu/%llu tm2 %llu/%llu tm3 %llu/%llu masked %llu/%llu (oracle/emulated)\n", sid,
			(unsigned long long)want.seeds_hit, (unsigned long long)got.seeds_hit, (unsigned long long)want.seed_hits, (unsigned long long)got.seed_hits, (unsigned long long)want.tentative_matches1, (unsigned long long)got.tentative_matches1,
			(unsigned long long)want.tentative_matches2, (unsigned long long)got.tentative_matches2, (unsigned long long)want.tentative_matches3, (unsigned long long)got.tentative_matches3, (unsigned long long)want.masked_seeds, (unsigned long long)got.masked_seeds); }
		if (oh.size() != eh.size() || memcmp(oh.data(), eh.data(), oh.size() * sizeof(dmnd_hit)) != 0) {
			++fails; size_t sc = 0, same_set = 0;
			for (size_t k = 0; k < std::min(oh.size(), eh.size()); ++k) { sc += (oh[k].subject_score >> 48) != (eh[k].subject_score >> 48); same_set += oh[k].query == eh[k].query && ((oh[k].subject_score ^ eh[k].subject_score) << 16) == 0; }
			printf("FAIL shape %d hits: oracle %zu emulated %zu; position-wise %zu with another score, %zu same (query, subject)\n", sid, oh.size(), eh.size(), sc, same_set);
		}
		if (memcmp(olet.data(), q.letters.data(), q.raw_len) != 0) { ++fails; printf("FAIL shape %d SEED_MASK bits differ\n", sid); }
	}
	printf("shapes=%d hits=%zu fails=%d \n", hp.n_shapes, total_hits, fails);
	return fails ? 1 : 0;
}
