// seed_kernels.cuh -- device code of the seed stage (dmnd_search_shape, dmnd_hits_xdrop) for sm_100a: kernels and device
// functions only.  The launch sequence is in seed.cu; tests/emu_seed.cpp compiles THIS file for the CPU behind tests/emu_cuda.h
// and runs the whole stage (index build, probe, entropy masking, stage 1/2, left-most filter) against the oracle.
#pragma once
#include "dev_params.h"
#include "mask_kernels.cuh"

namespace dmnd_cuda {

struct Entry { uint32_t qloc, lo, cnt, part; };

__device__ __forceinline__ uint64_t mix40(uint64_t seed) { return (seed * 0x9E3779B97F4A7C15ull) & 0xFFFFFFFFFFull; }

// basic/shape.h:113-171 on the reduced sequence; a window that contains a delimiter is past the end of its sequence
// (search/seed_array/seed_iterator.h:30-33).
__device__ __forceinline__ bool seed_at(const DevParams* P, int sid, const int8_t* s, const uint32_t* __restrict__ soft, size_t p, uint64_t& out) {
	const int span = P->shape_len[sid];
	bool ok = true;
	for (int k = 0; k < span; ++k) ok &= (s[k] != DMND_DELIMITER);
	if (!ok) return false;
	uint64_t v = 0;
	for (int k = 0; k < P->shape_weight; ++k) {
		if (soft && soft_bit(soft, p + (size_t)P->shape_pos[sid][k])) return false;
		const unsigned r = P->reduction[s[P->shape_pos[sid][k]] & 31];
		if (r == 23) return false;
		v = v * (uint64_t)P->reduction_size + r;
	}
	out = v;
	return true;
}

__global__ void ref_enum_kernel(const int8_t* __restrict__ letters, size_t raw_len, const DevParams* __restrict__ P, int sid,
                                const uint32_t* __restrict__ soft, uint64_t* keys, uint32_t* vals, unsigned long long* count) {
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x + DMND_PERIMETER_PADDING;
	uint64_t seed = 0;
	const bool ok = p + DMND_PERIMETER_PADDING < raw_len && letters[p] != DMND_DELIMITER && seed_at(P, sid, letters + p, soft, p, seed);
	const unsigned m = __ballot_sync(0xffffffffu, ok);
	if (m == 0) return;
	const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
	unsigned long long base = 0;
	if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(m));
	base = __shfl_sync(0xffffffffu, base, leader);
	if (ok) {
		const unsigned long long idx = base + __popc(m & ((1u << lane) - 1));
		keys[idx] = mix40(seed);
		vals[idx] = (uint32_t)p;
	}
}

// Blocked Bloom filter over the reference keys: one 32-byte block (= one L2 sector) per key, 4 bits set.  99 % of the query
// positions have no partner in the reference; the filter answers them from a structure sized to stay L2 resident
// (<= 32 reference keys per 256-bit block: 34 MB for the 3*10^7 keys of a 100 k-protein block, ~2 % false positives)
// instead of touching the bucket directory and the key array in HBM.  At 6.8 keys per block (134 MB) the filter itself
// missed L2 and the probe ran at DRAM random-sector speed (ncu: 62 B of DRAM traffic per query position).
__device__ __forceinline__ void bloom_slots(uint64_t key, uint32_t block_mask, uint32_t& block, uint32_t& bits) {
	const uint64_t h = key * 0xD6E8FEB86659FD93ull;
	block = (uint32_t)(h >> 40) & block_mask;
	bits = (uint32_t)(h >> 8);  // four 8-bit positions inside the 256-bit block
}
// First-level filter in front of the Bloom blocks: ONE bit per key in a table of >= 4 bits per reference key (16 MB for the 3*10^7
// keys of a 100 k-protein block).  The 34 MB of Bloom blocks do not stay L2 resident beside the streamed query letters on a B200
// (lines homed in the far L2 partition are cached twice: ncu showed 23 % of the Bloom sector reads going to DRAM, 4.3 GB per 10^6
// queries for 0.25 GB of letters); the bitmap does, and it answers ~80 % of the query positions, so only one position in five
// goes on to a Bloom block at all.
__device__ __forceinline__ uint32_t bitmap_slot(uint64_t key, uint32_t mask) { return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 33) & mask; }
__device__ __forceinline__ bool bitmap_test(const uint32_t* __restrict__ bitmap, uint32_t mask, uint64_t key) {
	const uint32_t b = bitmap_slot(key, mask);
	return (bitmap[b >> 5] >> (b & 31)) & 1u;
}
__global__ void bloom_build_kernel(const uint64_t* __restrict__ keys, size_t n, uint32_t* bloom, uint32_t block_mask, uint32_t* bitmap, uint32_t bitmap_mask) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (i > 0 && keys[i] == keys[i - 1]) return;  // sorted: one insert per distinct key
	{ const uint32_t b = bitmap_slot(keys[i], bitmap_mask); atomicOr(&bitmap[b >> 5], 1u << (b & 31)); }
	uint32_t block, bits;
	bloom_slots(keys[i], block_mask, block, bits);
	uint32_t* w = bloom + (size_t)block * 8;
#pragma unroll
	for (int k = 0; k < 4; ++k) { const uint32_t p = (bits >> (8 * k)) & 255u; atomicOr(&w[p >> 5], 1u << (p & 31)); }
}
__device__ __forceinline__ bool bloom_test(const uint32_t* __restrict__ bloom, uint32_t block_mask, uint64_t key) {
	uint32_t block, bits;
	bloom_slots(key, block_mask, block, bits);
	const uint4* w4 = reinterpret_cast<const uint4*>(bloom + (size_t)block * 8);
	const uint4 a = w4[0], b = w4[1];
	const uint32_t w[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
	bool ok = true;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const uint32_t p = (bits >> (8 * k)) & 255u;
		uint32_t word = 0;
#pragma unroll
		for (int x = 0; x < 8; ++x) if ((p >> 5) == (uint32_t)x) word = w[x];
		ok &= (word >> (p & 31)) & 1u;
	}
	return ok;
}

__global__ void bucket_hist_kernel(const uint64_t* __restrict__ keys, size_t n, int shift, uint32_t* hist) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) atomicAdd(&hist[(uint32_t)(keys[i] >> shift)], 1u);
}

// Shape of the seed, passed by value (kernel parameters live in the constant bank: uniform, no memory traffic).
struct ShapeArg {
	int8_t pos[DMND_MAX_WEIGHT];
	int weight, span, rsize, seedp_bits;
};

// Tile loader shared by the enumeration kernels: TILE consecutive letters (+ 32 halo) become "codes" in shared memory:
// reduced class 0..9, 0x40 for MASK/STOP (reduction 23), 0x80 for the delimiter.  One global byte per letter, coalesced.
#define SEED_TILE 1024
#define SEED_RAW_BYTES (SEED_TILE + 32 + 32)  // the tile's letters + the longest seed window, widened to 16-byte boundaries on both sides
// The tile's letters travel global -> shared as ONE bulk asynchronous copy (cp.async.bulk, the 1-D form of TMA; SASS UBLKCP): a
// single thread arms an mbarrier with the byte count and issues the copy, the copy engine completes the barrier, every thread waits
// on its phase.  Source and destination must be 16-byte aligned: the copy starts at p0 rounded down and ends at min(tile end,
// p_end + 32) rounded up (positions >= p_end are never evaluated, so what lies behind that is not needed -- and not read: the last
// tile of a block would otherwise reach past the block's padding).  Device builds only; the CPU emulation of this file
// (tests/emu_seed.cpp) takes the plain loop below.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void bulk_load_letters(const int8_t* __restrict__ src16, unsigned bytes, uint8_t* s_raw, unsigned long long* mbar) {
	const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar), dst = (unsigned)__cvta_generic_to_shared(s_raw);
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src16), "r"(bytes), "r"(mb) : "memory");
	}
	unsigned done = 0;
	while (!done)  // phase 0 of a barrier used once per CTA
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb) : "memory");
}
#endif
__device__ __forceinline__ void load_code_tile(const int8_t* __restrict__ letters, const uint32_t* __restrict__ soft, size_t p0, size_t p_end, const DevParams* __restrict__ P, uint8_t* s_code, uint8_t* s_lut,
                                               uint8_t* s_raw, unsigned long long* mbar, bool bulk) {
	if (threadIdx.x < 32) {
		const unsigned r = P->reduction[threadIdx.x];
		s_lut[threadIdx.x] = threadIdx.x == DMND_DELIMITER ? 0x80 : (r == 23 ? 0x40 : (uint8_t)r);
	}
#if defined(__CUDA_ARCH__)
	if (bulk) {
		const size_t a0 = p0 & ~(size_t)15, need_end = min(p0 + SEED_TILE + 32, p_end + 32), a1 = (need_end + 15) & ~(size_t)15;
		bulk_load_letters(letters + a0, (unsigned)(a1 - a0), s_raw, mbar);  // (contains the barrier that also publishes s_lut)
		const int off = (int)(p0 - a0), have = (int)(a1 - p0);
		if (soft) {
			for (int x = threadIdx.x; x < SEED_TILE + 32; x += blockDim.x) {
				const uint8_t c = x < have ? s_lut[s_raw[off + x] & 31] : (uint8_t)0x80;
				s_code[x] = (x < have && soft_bit(soft, p0 + x) && !(c & 0x80)) ? (uint8_t)0x40 : c;
			}
		}
		else
			for (int x = threadIdx.x; x < SEED_TILE + 32; x += blockDim.x) s_code[x] = x < have ? s_lut[s_raw[off + x] & 31] : (uint8_t)0x80;
		__syncthreads();
		return;
	}
#endif
	(void)s_raw; (void)mbar; (void)p_end; (void)bulk;
	__syncthreads();
	if (soft) {  // soft-masked letters read as MASK_LETTER (class flag 0x40); they are never delimiters
		for (int x = threadIdx.x; x < SEED_TILE + 32; x += blockDim.x) {
			const uint8_t c = s_lut[letters[p0 + x] & 31];
			s_code[x] = (soft_bit(soft, p0 + x) && !(c & 0x80)) ? (uint8_t)0x40 : c;
		}
	}
	else
		for (int x = threadIdx.x; x < SEED_TILE + 32; x += blockDim.x) s_code[x] = s_lut[letters[p0 + x] & 31];
	__syncthreads();
}
// Packed seed at tile offset `o` (basic/shape.h:113-171: base-`rsize` number of the reduced classes at the shape's '1'
// positions; invalid if a MASK class is among them or the window runs into a delimiter).
__device__ __forceinline__ bool seed_from_codes(const uint8_t* s_code, int o, const ShapeArg& sh, uint64_t& seed) {
	unsigned flags = 0;
	for (int k = 0; k < sh.span; ++k) flags |= s_code[o + k] & 0x80u;
	uint32_t hi = 0, lo = 0;
	const int wh = sh.weight / 2;
	for (int k = 0; k < wh; ++k) { const unsigned c = s_code[o + sh.pos[k]]; flags |= c & 0x40u; hi = hi * (uint32_t)sh.rsize + (c & 15u); }
	uint32_t pw = 1;
	for (int k = wh; k < sh.weight; ++k) { const unsigned c = s_code[o + sh.pos[k]]; flags |= c & 0x40u; lo = lo * (uint32_t)sh.rsize + (c & 15u); pw *= (uint32_t)sh.rsize; }
	seed = (uint64_t)hi * pw + lo;
	return flags == 0;
}

__global__ void __launch_bounds__(256) probe_kernel(const int8_t* __restrict__ letters, const uint32_t* __restrict__ soft, size_t p_begin, size_t p_end, const DevParams* __restrict__ P, const ShapeArg sh,
                             const uint64_t* __restrict__ keys, const uint32_t* __restrict__ bucket, int shift,
                             const uint32_t* __restrict__ bloom, uint32_t bloom_mask, const uint32_t* __restrict__ bitmap, uint32_t bitmap_mask,
                             Entry* entries, unsigned long long* count, unsigned long long cap, int bulk) {
	__shared__ uint8_t s_code[SEED_TILE + 32];
	__shared__ uint8_t s_lut[32];
	__shared__ __align__(16) uint8_t s_raw[SEED_RAW_BYTES];
	__shared__ __align__(8) unsigned long long s_mbar;
	// matches of the tile are staged in shared memory: ONE pair of global atomics per CTA (returning atomics on a single
	// hot address cost microseconds each and stalled every warp that found a match) and a coalesced copy-out
	__shared__ Entry s_ent[SEED_TILE];
	__shared__ unsigned s_n;
	__shared__ unsigned long long s_pairs, s_base;
	const size_t p0 = p_begin + (size_t)blockIdx.x * SEED_TILE;
	if (threadIdx.x == 0) { s_n = 0; s_pairs = 0; }
	load_code_tile(letters, soft, p0, p_end, P, s_code, s_lut, s_raw, &s_mbar, bulk != 0);
	for (int it = 0; it < SEED_TILE / 256; ++it) {
		const int o = it * 256 + threadIdx.x;
		const size_t p = p0 + o;
		uint64_t seed = 0;
		bool ok = p < p_end && seed_from_codes(s_code, o, sh, seed);
		uint64_t key = 0;
		if (ok) { key = mix40(seed); ok = (bitmap == nullptr || bitmap_test(bitmap, bitmap_mask, key)) && bloom_test(bloom, bloom_mask, key); }
		if (ok) {
			const uint32_t b = (uint32_t)(key >> shift);
			uint32_t i = bucket[b];
			const uint32_t e = bucket[b + 1];
			while (i < e && keys[i] < key) ++i;
			const uint32_t lo = i;
			while (i < e && keys[i] == key) ++i;
			const uint32_t cnt = i - lo;
			if (cnt > 0) {
				s_ent[atomicAdd(&s_n, 1u)] = Entry{ (uint32_t)p, lo, cnt, (uint32_t)(seed & (((uint64_t)1 << sh.seedp_bits) - 1)) };
				atomicAdd(&s_pairs, (unsigned long long)cnt);  // upper bound on (q,s) pairs over all chunks
			}
		}
	}
	__syncthreads();
	const unsigned n = s_n;
	if (n == 0) return;
	if (threadIdx.x == 0) {
		s_base = atomicAdd(count, (unsigned long long)n);
		atomicAdd(count + 1, s_pairs);
	}
	__syncthreads();
	const unsigned long long base = s_base;
	for (unsigned k = threadIdx.x; k < n; k += blockDim.x)
		if (base + k < cap) entries[base + k] = s_ent[k];
}

// search/seed_complexity.cpp:37-51
__device__ __forceinline__ bool seed_is_complex(const DevParams* P, int sid, const int8_t* seq) {
	unsigned count[20];
#pragma unroll
	for (int i = 0; i < 20; ++i) count[i] = 0;
	for (int k = 0; k < P->shape_weight; ++k) {
		const int l = seq[P->shape_pos[sid][k]] & 31;
		if (l >= 20) return false;
		++count[P->reduction[l]];
	}
	double entropy = P->lnfact[P->shape_weight];
	for (int c = 0; c < P->reduction_size; ++c) entropy -= P->lnfact[count[c]];
	return entropy >= P->seed_cut;
}

// Chunk pass 1: entropy masking.  pairs[e] = number of (q,s) pairs entry e contributes to THIS chunk's search.
__global__ void mask_kernel(int8_t* q_letters, const DevParams* __restrict__ P, int sid, Entry* entries, size_t n, uint32_t pb, uint32_t pe,
                            uint64_t* pairs, uint32_t* key_seen /* bitmap over reference run starts */, unsigned long long* counters) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Entry e = entries[i];
	uint64_t np = 0;
	if (e.part >= pb && e.part < pe) {
		const bool first = (atomicOr(&key_seen[e.lo >> 5], 1u << (e.lo & 31)) & (1u << (e.lo & 31))) == 0;  // one per shared key
		if (!seed_is_complex(P, sid, q_letters + e.qloc)) {
			q_letters[e.qloc] = (int8_t)(q_letters[e.qloc] | DMND_SEED_MASK);  // one writer per byte
			entries[i].cnt = 0;
			if (first) atomicAdd(&counters[7], 1ull);
		}
		else { np = e.cnt; if (first) atomicAdd(&counters[0], 1ull); }
	}
	pairs[i] = np;
}

// search/hamming/finger_print.h:180-215: equal letters (bits 0..4) among the 48 positions [-16, 32) around both seeds.  Both windows
// are read as thirteen aligned 32-bit words each and shifted into place (funnel shift by the byte misalignment), compared four
// letters at a time: (x + 0x7f7f7f7f) & 0x80808080 has one bit per UNEQUAL letter (x = masked XOR, every byte <= 0x1f: no carries).
// The aligned reads touch at most 3 bytes before q - 16 and after q + 32: inside the block image's padding / allocation slack.
__device__ __forceinline__ unsigned fingerprint_match(const int8_t* q, const int8_t* s) {
	const uintptr_t qa = (uintptr_t)(q - 16), sa = (uintptr_t)(s - 16);
	const uint32_t* __restrict__ qw = reinterpret_cast<const uint32_t*>(qa & ~(uintptr_t)3);
	const uint32_t* __restrict__ sw = reinterpret_cast<const uint32_t*>(sa & ~(uintptr_t)3);
	const unsigned qsh = (unsigned)(qa & 3) * 8u, ssh = (unsigned)(sa & 3) * 8u;
	uint32_t qp = qw[0], sp = sw[0];
	unsigned diff = 0;
#pragma unroll
	for (int k = 0; k < 12; ++k) {
		const uint32_t qn = qw[k + 1], sn = sw[k + 1];
		const uint32_t x = (__funnelshift_r(qp, qn, qsh) ^ __funnelshift_r(sp, sn, ssh)) & 0x1f1f1f1fu;
		diff += (unsigned)__popc((x + 0x7f7f7f7fu) & 0x80808080u);
		qp = qn; sp = sn;
	}
	return 48u - diff;
}

struct LmCtx {
	const DevParams* P;
	const uint8_t* cur_matcher; uint32_t cur_minlen, cur_suffix;
	const uint8_t* prev_matcher; uint32_t prev_minlen, prev_suffix;
	int sid, chunked;
	uint32_t range_begin, range_end;
};

__device__ __forceinline__ uint32_t matcher_hit(const uint8_t* table, uint32_t minlen, uint32_t suffix, uint32_t h, uint32_t len) {
	if (len < minlen) return 0;
	const uint32_t end = len - minlen + 1;
	uint32_t r = 0;
	for (uint32_t i = 0; i < end; ++i) { r |= (uint32_t)table[h & suffix] << i; h >>= 1; }
	return r;
}

// search/left_most.h:30-49
__device__ bool verify_hit(const LmCtx& x, const int8_t* q, const int8_t* s, bool left, uint32_t match_mask) {
	const DevParams* P = x.P;
	if (x.chunked) {
		const uint32_t sm = P->shape_mask[x.sid];
		if ((sm & match_mask) == sm) {
			uint64_t seed = 0;
			for (int k = 0; k < P->shape_weight; ++k) {  // Shape::set_seed, basic/shape.h:73-96
				const int l = s[P->shape_pos[x.sid][k]] & 31;
				if (l == 23 || l == 31 || l == 24) return false;
				seed = seed * (uint64_t)P->reduction_size + P->reduction[l];
			}
			const uint32_t part = (uint32_t)(seed & (((uint64_t)1 << P->seedp_bits) - 1));
			if (left && !(part < x.range_end)) return false;
			if (!left && !(part < x.range_begin)) return false;
		}
	}
	return fingerprint_match(q, s) >= (unsigned)P->hamming_id;
}
// search/left_most.h:51-60
__device__ bool verify_hits(const LmCtx& x, uint32_t mask, const int8_t* q, const int8_t* s, bool left, uint32_t match_mask) {
	int shift = 0;
	while (mask != 0) {
		const int i = __ffs(mask) - 1;
		if (verify_hit(x, q + i + shift, s + i + shift, left, (i + shift) < 32 ? match_mask >> (i + shift) : 0u)) return true;
		mask = (i + 1) < 32 ? mask >> (i + 1) : 0u;
		shift += i + 1;
	}
	return false;
}
// util/sequence/sequence.h:30-40 on [seq, seq+len) around seq+anchor
__device__ __forceinline__ void clip(const int8_t* seq, int len, int anchor, int& b, int& e) {
	b = 0; e = len;
	for (int k = 0; k < len; ++k)
		if (seq[k] == DMND_DELIMITER) {
			if (k >= anchor) { e = k; return; }
			b = k + 1;
		}
}
// search/left_most.h:62-110
// __noinline__ on purpose.  Inlined into stage2_window_kernel, ptxas 12.9 (sm_100a) produced code whose result depended on how the
// lanes of a warp diverged in the verify loops: on the B200 the filter kept 2-5 of ~3 100 pairs that the same code drops when it
// runs one thread per warp, differently from run to run, only with two or more shapes (candidates that skip the partition test and
// go straight to the fingerprint).  As a real call -- or with fingerprint_match out of line -- the kernel is deterministic and
// equal to the oracle (profiles/lm_variants_r2.txt: variants 1, 5, 6 against 0, 2, 7).
__device__ __noinline__ bool left_most_filter(const LmCtx& x, const int8_t* query, int query_len, const int8_t* subject, int seed_offset, int seed_len) {
	const DevParams* P = x.P;
	int d = max(seed_offset - 16, 0), window_left = min(16, seed_offset);
	const int8_t *q = query + d, *s = subject + d;
	int window = min(query_len - d, window_left + 1 + 32);
	int cb, ce;
	clip(s, window, window_left, cb, ce);
	window = ce;
	d = cb;
	q += d; s += d; window_left -= d; window -= d;
	uint64_t match_mask = 0, seed_bits = 0;
	for (int k = 0; k < window && k < 64; ++k) {
		if (P->map8[q[k] & 31] == P->map8b[s[k] & 31]) match_mask |= (uint64_t)1 << k;
		if (q[k] & DMND_SEED_MASK) seed_bits |= (uint64_t)1 << k;
	}
	const uint64_t query_seed_mask = ~seed_bits;
	const uint32_t len_left = (uint32_t)(window_left + seed_len - 1),
		match_mask_left = (uint32_t)((((uint64_t)1 << len_left) - 1) & match_mask),
		query_mask_left = (uint32_t)((((uint64_t)1 << len_left) - 1) & query_seed_mask);
	const uint32_t left_hit = matcher_hit(x.cur_matcher, x.cur_minlen, x.cur_suffix, match_mask_left, len_left) & query_mask_left;
	if (x.sid == 0 && !x.chunked) return left_hit == 0 || !verify_hits(x, left_hit, q, s, true, match_mask_left);
	const uint32_t len_right = (uint32_t)(window - window_left - 1),
		match_mask_right = (uint32_t)(match_mask >> (window_left + 1)),
		query_mask_right = (uint32_t)(query_seed_mask >> (window_left + 1));
	const uint32_t right_hit = (x.chunked ? matcher_hit(x.cur_matcher, x.cur_minlen, x.cur_suffix, match_mask_right, len_right)
	                                      : matcher_hit(x.prev_matcher, x.prev_minlen, x.prev_suffix, match_mask_right, len_right)) & query_mask_right;
	return (left_hit == 0 || !verify_hits(x, left_hit, q, s, true, match_mask_left))
		&& (right_hit == 0 || !verify_hits(x, right_hit, q + window_left + 1, s + window_left + 1, false, match_mask_right));
}

// What follows the Hamming filter for one (query loc, reference loc) pair (search/stage2.h:73-154): the ungapped window
// filter when the mode has one (batch_size >= 0: number of stage-1 survivors that share this pair's window_ungapped_best call,
// search/stage2.h:114-120; < 0: stage skipped, score 0xFFFF), the left-most filter, the hit.
__device__ __forceinline__ void stage2_tail(const int8_t* __restrict__ q_letters, const int64_t* __restrict__ q_limits, uint32_t nq,
                                            const int8_t* __restrict__ r_letters, const Entry& e, uint32_t sloc, const LmCtx& x, int batch_size,
                                            dmnd_hit* hits, unsigned long long* hit_count, unsigned long long* counters) {
	const int8_t *qp = q_letters + e.qloc, *sp = r_letters + sloc;
	// query id / seed offset (SequenceSet::local_position)
	uint32_t a = 0, b = nq;
	while (b - a > 1) { const uint32_t mid = a + (b - a) / 2; if ((uint64_t)q_limits[mid] <= (uint64_t)e.qloc) a = mid; else b = mid; }
	const int seed_offset = (int)((int64_t)e.qloc - q_limits[a]);
	// search/stage2.h:92-103; ungapped_window(query_len), :58-63: a translated frame of <= 85 letters is scored over its whole length
	const int query_len = (int)(q_limits[a + 1] - q_limits[a] - 1);
	const bool short_frame = x.P->query_contexts > 1 && query_len <= 85;
	const int window = short_frame ? query_len : x.P->ungapped_window;
	int cb, ce;
	clip(qp - window, 2 * window, window, cb, ce);
	const int window_left = window - cb, window_clipped = ce - cb;
	const int8_t* qc = qp - window + cb;
	uint32_t score16 = 0xFFFFu;
	if (batch_size >= 0) {
		// ungapped_cutoff (search/stage2.h:41-57) and the window score: scalar ungapped_window (dp/ungapped_align.cpp:244-257) for
		// calls with < 4 subjects, the int8 kernel (dp/ungapped_simd.cpp:32-88) otherwise -- its biased saturating lanes differ
		// from the scalar loop only by capping the result at 255
		const int cutoff = query_len <= x.P->short_query_max_len ? x.P->short_query_ungapped_cutoff
			: (short_frame ? x.P->ungapped_cutoff_short : x.P->ungapped_cutoff)[32 - __clz((unsigned)query_len)];
		const int8_t* sw = sp - window_left;
		int st = 0, best = 0;
		for (int t = 0; t < window_clipped; ++t) {
			st += (int)x.P->score[((qc[t] & 31) << 5) | (sw[t] & 31)];
			st = max(st, 0);
			best = max(best, st);
		}
		const int score = (batch_size >= 4 && best > 255) ? 255 : best;
		if (!(score > cutoff)) return;
		atomicAdd(&counters[8], 1ull);
		score16 = (uint32_t)score & 0xFFFFu;
	}
	const int interval_mod = x.P->left_most_interval > 0 ? seed_offset % x.P->left_most_interval : window_left;
	const int overhang = max(window_left - interval_mod, 0);
	if (!left_most_filter(x, qc + overhang, window_clipped - overhang, sp - window_left + overhang, window_left - overhang, x.P->shape_len[x.sid])) return;
	const unsigned long long idx = atomicAdd(hit_count, 1ull);
	dmnd_hit h;
	h.query = a; h.seed_offset = seed_offset; h.subject_score = (uint64_t)sloc | ((uint64_t)score16 << 48);
	hits[idx] = h;
}

// ---- which (entry, k) is pair `pid` of the chunk?  The chunk's ACTIVE entries (those that contribute pairs: right partition range,
// seed not masked) are listed in act[] with the running pair count act_off[] (act_off[nact] = pairs of the chunk).  One binary search
// per CTA finds the entry of the CTA's first pair; the CTA's 128 pairs lie in at most 128 further active entries (each has >= 1 pair),
// whose offsets go to shared memory, and every thread finishes with a 7-step search there.  (A per-pair binary search over ALL
// entries -- 28 dependent L2 reads where every query position has a partner, as in the sensitive modes -- was most of those modes' time.)
struct PairLookup { const uint32_t* act; const uint64_t* act_off; const uint32_t* nact; };
constexpr int STAGE_CTA = 128;
__device__ __forceinline__ bool locate_pair(const PairLookup& L, uint64_t* s_off, uint32_t* s_first, uint64_t& pid, size_t& entry, uint32_t& k) {
	const uint32_t nact = *L.nact;
	const uint64_t total = L.act_off[nact], pid0 = (uint64_t)blockIdx.x * STAGE_CTA;
	pid = pid0 + threadIdx.x;
	if (pid0 >= total) return false;  // the whole CTA is beyond the chunk's pairs (the grid covers a bound over all chunks)
	if (threadIdx.x == 0) {
		uint32_t lo = 0, hi = nact;
		while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (L.act_off[mid] <= pid0) lo = mid; else hi = mid; }
		*s_first = lo;
	}
	__syncthreads();
	const uint32_t first = *s_first;
	for (int t = threadIdx.x; t <= STAGE_CTA; t += STAGE_CTA) s_off[t] = L.act_off[min(first + (uint32_t)t, nact)];
	__syncthreads();
	if (pid >= total) return false;
	int lo = 0, hi = STAGE_CTA + 1;
	while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= pid) lo = mid; else hi = mid; }
	entry = (size_t)L.act[first + (uint32_t)lo];
	k = (uint32_t)(pid - s_off[lo]);
	return true;
}
// rank[] = exclusive count of active entries before i (scan of pairs[i] > 0): the scatter builds act / act_off from it
__global__ void active_scatter_kernel(const uint64_t* __restrict__ pairs, const uint64_t* __restrict__ pair_off, const uint32_t* __restrict__ rank, size_t n,
                                      uint32_t* act, uint64_t* act_off, uint32_t* nact) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n) return;
	if (i == n) { const uint32_t na = rank[n]; *nact = na; act_off[na] = pair_off[n]; return; }
	if (pairs[i] > 0) { act[rank[i]] = (uint32_t)i; act_off[rank[i]] = pair_off[i]; }
}

// Chunk pass 2 (modes without the ungapped window filter, --fast): one thread per (query loc, reference loc) pair of the
// chunk's surviving keys.
__global__ void __launch_bounds__(STAGE_CTA) stage12_kernel(const int8_t* __restrict__ q_letters, const int64_t* __restrict__ q_limits, uint32_t nq,
                               const int8_t* __restrict__ r_letters, const Entry* __restrict__ entries, PairLookup L,
                               const uint32_t* __restrict__ ref_locs, LmCtx x, dmnd_hit* hits, unsigned long long* hit_count,
                               unsigned long long* counters) {
	__shared__ uint64_t s_off[STAGE_CTA + 1];
	__shared__ uint32_t s_first;
	uint64_t pid; size_t lo; uint32_t k;
	if (!locate_pair(L, s_off, &s_first, pid, lo, k)) return;
	const Entry e = entries[lo];
	const uint32_t sloc = ref_locs[e.lo + k];
	if (fingerprint_match(q_letters + e.qloc, r_letters + sloc) < (unsigned)x.P->hamming_id) return;
	atomicAdd(&counters[2], 1ull);
	stage2_tail(q_letters, q_limits, nq, r_letters, e, sloc, x, -1, hits, hit_count, counters);
}

// Modes WITH the ungapped window filter: the reference scores the stage-1 survivors of one query location in calls of up to 32
// subjects (per 1024-subject tile of the key, ascending subject order: search/hamming/kernel.h:61-74, hit_field.h:44-57), and
// the size of a survivor's call decides which window kernel scores it.  Pass A writes one survivor bit per pair (one ballot
// word per warp, the grid covers the bound so every word is written) AND appends the survivors (4-5 % of the pairs with the
// low-weight shapes of the sensitive modes) to a compact list; pass B runs over that list only -- a grid over all pairs spent 40 %
// of the --sensitive seed stage finding out that a thread's pair had not survived -- and counts the survivors of the pair's tile
// in the bitmap to find the size of its call.
struct Survivor { uint64_t pid; uint32_t entry, k; };
__global__ void __launch_bounds__(STAGE_CTA) stage1_flags_kernel(const int8_t* __restrict__ q_letters, const int8_t* __restrict__ r_letters, const Entry* __restrict__ entries,
                                    PairLookup L, const uint32_t* __restrict__ ref_locs, unsigned hamming_id, uint32_t* flags,
                                    Survivor* surv, unsigned long long* surv_count, unsigned long long surv_cap, unsigned long long* counters) {
	__shared__ uint64_t s_off[STAGE_CTA + 1];
	__shared__ uint32_t s_first;
	uint64_t pid; size_t lo; uint32_t k;
	bool pass = false;
	if (locate_pair(L, s_off, &s_first, pid, lo, k)) {
		const Entry e = entries[lo];
		const uint32_t sloc = ref_locs[e.lo + k];
		pass = fingerprint_match(q_letters + e.qloc, r_letters + sloc) >= hamming_id;
	}
	const unsigned word = __ballot_sync(0xffffffffu, pass);
	const unsigned lane = threadIdx.x & 31;
	unsigned long long base = 0;
	if (lane == 0) {
		flags[pid >> 5] = word;
		if (word) { atomicAdd(&counters[2], (unsigned long long)__popc(word)); base = atomicAdd(surv_count, (unsigned long long)__popc(word)); }
	}
	if (word) {
		base = __shfl_sync(0xffffffffu, base, 0);
		if (pass && base + 32 <= surv_cap) surv[base + __popc(word & ((1u << lane) - 1u))] = Survivor{ pid, (uint32_t)lo, k };  // (a full list is detected by the count: the window pass then takes the grid over all pairs)
	}
}
__device__ __forceinline__ unsigned count_bits(const uint32_t* __restrict__ bits, uint64_t a, uint64_t b) {  // set bits in [a, b)
	unsigned n = 0;
	while (a < b) {
		const unsigned sh = (unsigned)(a & 31);
		const uint64_t take = min((uint64_t)(32 - sh), b - a);
		const uint32_t m = (take == 32 ? 0xffffffffu : ((1u << take) - 1u)) << sh;
		n += __popc(bits[a >> 5] & m);
		a += take;
	}
	return n;
}
__global__ void __launch_bounds__(STAGE_CTA) stage2_window_kernel(const int8_t* __restrict__ q_letters, const int64_t* __restrict__ q_limits, uint32_t nq,
                                     const int8_t* __restrict__ r_letters, const Entry* __restrict__ entries, PairLookup L,
                                     const uint32_t* __restrict__ ref_locs, const uint32_t* __restrict__ flags, LmCtx x,
                                     const Survivor* __restrict__ surv, const unsigned long long* __restrict__ surv_count, unsigned long long surv_cap,
                                     dmnd_hit* hits, unsigned long long* hit_count, unsigned long long* counters) {
	(void)L;
	const unsigned long long n = *surv_count;  // known on the device only: a fixed grid strides over the list
	if (n + 32 > surv_cap) return;             // the list overflowed (sized for 1/8 of the pairs): stage2_window_full_kernel does this chunk
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
		const Survivor s = surv[i];
		const uint64_t pid = s.pid;
		const uint32_t k = s.k;
		const Entry e = entries[s.entry];
		const uint64_t first = pid - k;
		const uint32_t tile_begin = k & ~1023u, tile_end = min(tile_begin + 1024u, e.cnt);
		const unsigned rank = count_bits(flags, first + tile_begin, pid), total = count_bits(flags, first + tile_begin, first + tile_end);
		const int batch_size = (int)min(32u, total - (rank & ~31u));
		stage2_tail(q_letters, q_limits, nq, r_letters, e, ref_locs[e.lo + k], x, batch_size, hits, hit_count, counters);
	}
}

// the same pass as a grid over all pairs of the chunk: only when the survivor list overflowed
__global__ void __launch_bounds__(STAGE_CTA) stage2_window_full_kernel(const int8_t* __restrict__ q_letters, const int64_t* __restrict__ q_limits, uint32_t nq,
                                     const int8_t* __restrict__ r_letters, const Entry* __restrict__ entries, PairLookup L,
                                     const uint32_t* __restrict__ ref_locs, const uint32_t* __restrict__ flags, LmCtx x,
                                     const unsigned long long* __restrict__ surv_count, unsigned long long surv_cap,
                                     dmnd_hit* hits, unsigned long long* hit_count, unsigned long long* counters) {
	__shared__ uint64_t s_off[STAGE_CTA + 1];
	__shared__ uint32_t s_first;
	if (*surv_count + 32 <= surv_cap) return;
	uint64_t pid; size_t lo; uint32_t k;
	if (!locate_pair(L, s_off, &s_first, pid, lo, k) || !((flags[pid >> 5] >> (pid & 31)) & 1u)) return;
	const Entry e = entries[lo];
	const uint64_t first = pid - k;
	const uint32_t tile_begin = k & ~1023u, tile_end = min(tile_begin + 1024u, e.cnt);
	const unsigned rank = count_bits(flags, first + tile_begin, pid), total = count_bits(flags, first + tile_begin, first + tile_end);
	const int batch_size = (int)min(32u, total - (rank & ~31u));
	stage2_tail(q_letters, q_limits, nq, r_letters, e, ref_locs[e.lo + k], x, batch_size, hits, hit_count, counters);
}

__global__ void extract_query_kernel(const dmnd_hit* __restrict__ h, size_t n, uint32_t* keys) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) keys[i] = h[i].query;
}

// xdrop_ungapped (dp/ungapped_align.cpp:150-214, ScoreOnly + bias): one thread per hit
__global__ void xdrop_kernel(const int8_t* __restrict__ q_letters, const int8_t* __restrict__ q_bias, const int64_t* __restrict__ q_limits,
                             const int8_t* __restrict__ r_letters, const int64_t* __restrict__ r_limits, uint32_t nr,
                             const dmnd_hit* __restrict__ hits, size_t n, const DevParams* __restrict__ P, int xdrop, dmnd_segment* out, dmnd_hit_site* sites) {
	__shared__ int8_t s_score[1024];
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_score[i] = P->score[i];
	__syncthreads();
	const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const dmnd_hit hit = hits[k];
	const uint64_t sloc = hit.subject_score & 0xFFFFFFFFFFFFull;
	uint32_t a = 0, b = nr;
	while (b - a > 1) { const uint32_t mid = a + (b - a) / 2; if ((uint64_t)r_limits[mid] <= sloc) a = mid; else b = mid; }
	const int64_t qo = q_limits[hit.query];
	const int8_t *qs = q_letters + qo, *cb = q_bias + qo, *ss = r_letters + r_limits[a];
	const int qa = hit.seed_offset, sa = (int)((int64_t)sloc - r_limits[a]);
	int score = 0, st = 0, n1 = 1, delta = 0, len = 0;
	int q = qa - 1, s = sa - 1;
	for (;;) {
		if (!(score - st < xdrop)) break;
		const int ql = qs[q] & 31, sl = ss[s] & 31;
		if (ql == DMND_DELIMITER || sl == DMND_DELIMITER) break;
		st += (int)s_score[(ql << 5) | sl] + (int)cb[q];
		if (st > score) { score = st; delta = n1; }
		--q; --s; ++n1;
	}
	q = qa; s = sa; st = score; n1 = 1;
	for (;;) {
		if (!(score - st < xdrop)) break;
		const int ql = qs[q] & 31, sl = ss[s] & 31;
		if (ql == DMND_DELIMITER || sl == DMND_DELIMITER) break;
		st += (int)s_score[(ql << 5) | sl] + (int)cb[q];
		if (st > score) { score = st; len = n1; }
		++q; ++s; ++n1;
	}
	out[k] = dmnd_segment{ qa - delta, sa - delta, len + delta, score };
	if (sites) sites[k] = dmnd_hit_site{ a, sa };
}

}  // namespace dmnd_cuda
