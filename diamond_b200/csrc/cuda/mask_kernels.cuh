// mask_kernels.cuh -- device code of dmnd_block_mask (sequence masking of a resident block) for sm_100a.  Kernels only: the
// launch sequence is in mask.cu; tests/emu_mask.cpp compiles THIS file for the CPU behind a lock-step emulation of the CUDA
// built-ins it uses and checks every kernel against the oracle.
//
//   tantan hard masking   masking/tantan.cpp:113-214 (Util::tantan::mask, mask_mode 1), called for every sequence of a loaded
//                         query / reference block (masking/masking.cpp:155-166, run/double_indexed.cpp:122-127, :737-741)
//   motif soft masking    masking/masking.cpp:110-131 (mask_motifs), MaskingTable (:76-107), Block::soft_mask (data/block/block.cpp:162-178)
//
// tantan is a forward-backward pass over a 50-state repeat HMM, sequential along the sequence and 50-wide across it.  The
// reference's AVX2 object evaluates it in fp32 with 8-lane registers and no FMA (util/simd/vector8_avx2.h:124-139 without
// __FMA__), so the result of every letter is fixed by an evaluation ORDER: products and sums are separate roundings, a
// register's horizontal sum is ((a0+a4)+(a1+a5))+((a2+a6)+(a3+a7)), the six register sums and the two tail states are added
// left to right.  Here EIGHT LANES own one sequence (lane k holds states 8r+k, r = 0..5, every lane carries the tail states 48
// and 49 redundantly), four sequences per warp: the xor-shuffle tree 4,1,2 reproduces hsum's association exactly (fp32 add
// is commutative, so all eight lanes end with the same bits), everything else is lane-local.  All arithmetic goes through
// __fmul_rn / __fadd_rn / __fdiv_rn: never contracted, never reassociated.  Bound: shuffle + LSU issue (18 shuffles, 8 byte
// loads and 8 shared-memory ratio look-ups per 4 letters), not HBM -- the pass reads each letter ~100 times from L1 and
// writes 4 B (the scaled background probability pb[i], tantan.cpp:186) per letter.
//
// Motifs: one thread per letter position looks its 8-mer up (base-20 code, util/kmer/kmer.h:38-47) in the 1000-entry table
// held in shared memory behind a 64 Kbit prefilter; hits set bits in a coverage bitmap; the (rare) sequences with a hit are
// finished by one thread each: the >= 50 % rule, merged ranges, ranges of <= max_motif_len letters become `soft` bits.
#pragma once
#include "dev_params.h"
#include "../host/motif_table.h"

namespace dmnd_cuda {

// One bit per letter: inside a MaskingTable entry (abundant motif, dmnd_block_mask).  While seeds are enumerated such a
// letter reads as MASK_LETTER (Block::soft_mask, data/block/block.cpp:162-171, search/seed_array/enum_seeds.h:262-270).
__device__ __forceinline__ bool soft_bit(const uint32_t* __restrict__ soft, size_t p) { return (soft[p >> 5] >> (p & 31)) & 1u; }

// hsum (util/simd/vector8_avx2.h:132-139) over the 8 lanes of a group; every lane returns the same value.  All four groups
// of a warp run in lock step (see tantan_kernel), so the shuffles use the full mask: xor 4, 1, 2 never leaves a group.
__device__ __forceinline__ float hsum8(float v) {
	const float x = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
	const float y = __fadd_rn(x, __shfl_xor_sync(0xffffffffu, x, 1));
	return __fadd_rn(y, __shfl_xor_sync(0xffffffffu, y, 2));
}

// (sequence length, sequence id) of the range, input of the longest-first order the tantan kernel walks
static __global__ void seq_len_kernel(const int64_t* __restrict__ limits, uint32_t s_begin, uint32_t n, uint32_t* len, uint32_t* id) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	const uint32_t sid = s_begin + t;
	len[t] = (uint32_t)(limits[sid + 1] - limits[sid] - 1);
	id[t] = sid;
}

// Eight lanes per sequence, four sequences per warp, `order` = sequence ids sorted by length (longest first), so that the
// four sequences of a warp have (almost) the same length and the warp can run ONE loop of maxlen steps in lock step:
//  * forward pass: a shorter sequence starts late (i = t - (maxlen - len)); before its start it is fed delimiters, whose
//    likelihood ratios are 0, which keeps f == 0 exactly, and b is set to 1 at i == 0 -- all four end together;
//  * backward pass: all start together at their own last letter (i = len - 1 - t); a sequence that is finished keeps
//    stepping on delimiters and stores nothing.
// Lock step makes every shuffle a full-mask shuffle (no per-group convergence barriers) and lets the letters travel in
// registers: lane k keeps the six letters its states look back at (i-1-8r-k, r = 0..5) packed 5 bits each in W, every lane
// keeps the two tail letters (i-49, i-50); one step shifts the window by two shuffles instead of eight byte loads.
// e_seg[off] (tantan.cpp:163-171,181) = ratio(letter i, letter i-1-off) with 0 before the sequence start == ratio row of the
// delimiter code 31, which dmnd_params_init leaves at 0.
//
// The two passes are two kernels, because most sequences never need the second one.  pf(i) = 1 - P(background at i) and
// every path weight of the HMM is non-negative, so P(background at i) >= W_bg / Z for every i, where Z is the total weight
// (tantan.cpp:189, what the forward pass ends with) and W_bg = b2b^(len+1) the weight of the all-background path.  A
// sequence with Z < 8 W_bg therefore has pf(i) <= 0.875 everywhere: nothing reaches p_mask = 0.9 and the backward pass can
// decide nothing.  (The margin to 0.9, a factor 1.25 in Z, is far beyond the rounding of the fp32 recurrences -- sums of
// positive terms -- at the lengths the shortcut is allowed for; longer sequences always take the full computation.)  ~98 %
// of random 300-letter sequences and ~93 % of a synthetic 100 k-protein block are decided by the forward kernel alone; the
// masked letters are identical either way (tests: emulation + device vs the oracle, which knows no shortcut).
#define TANTAN_CERT_MAX_LEN 4096
#define TANTAN_CERT_LOG2_RATIO 3.0f

struct TantanGroup {  // what both passes derive from the slot of an 8-lane group
	uint32_t sid;
	int64_t beg;
	int len, maxlen;
};
__device__ __forceinline__ TantanGroup tantan_group(const int64_t* __restrict__ limits, const uint32_t* __restrict__ order, uint32_t n_seq) {
	TantanGroup g;
	const uint32_t slot = (uint32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 3);
	g.sid = 0; g.beg = 0; g.len = 0;
	if (slot < n_seq) { g.sid = order[slot]; g.beg = limits[g.sid]; g.len = (int)(limits[g.sid + 1] - g.beg - 1); }
	int m = max(g.len, __shfl_xor_sync(0xffffffffu, g.len, 8));
	g.maxlen = max(m, __shfl_xor_sync(0xffffffffu, m, 16));
	return g;
}

// Forward pass of every sequence of the range: stores pb[i] (tantan.cpp:186), the scale factors (:181-185), 1/z per sequence
// and whether the backward pass is needed at all.
static __global__ void __launch_bounds__(128) tantan_forward_kernel(const int8_t* __restrict__ letters, const int64_t* __restrict__ limits, const uint32_t* __restrict__ order,
                                                                    uint32_t n_seq, const DevParams* __restrict__ P, float* pb, float* scale, int64_t base, uint32_t s_begin,
                                                                    float* zinv_out /* by sid - s_begin */, uint8_t* need /* by slot */) {
	__shared__ float s_lr[1024];
	for (int x = threadIdx.x; x < 1024; x += blockDim.x) s_lr[x] = P->tantan_lr[x];
	__syncthreads();
	const int lane = threadIdx.x & 31, k = lane & 7, lead = lane & 24;
	const uint32_t slot = (uint32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 3);
	const TantanGroup g = tantan_group(limits, order, n_seq);
	const int len = g.len, maxlen = g.maxlen;
	if (maxlen <= 0) {  // warp-uniform: nothing but empty slots / empty sequences
		if (k == 0 && slot < n_seq) need[slot] = 0;
		return;
	}
	const int8_t* seq = letters + g.beg;
	float* pbs = pb + (g.beg - base);
	float* scs = scale + ((g.beg - base) >> 4) + (g.sid - s_begin);
	float d[6];
#pragma unroll
	for (int r = 0; r < 6; ++r) d[r] = P->tantan_d[8 * r + k];
	const float d48 = P->tantan_d[48], d49 = P->tantan_d[49];
	const float b2b = P->tantan_b2b, f2f = P->tantan_f2f, p_repeat_end = P->tantan_p_repeat_end;
	float f[6], f48 = 0.0f, f49 = 0.0f;
#pragma unroll
	for (int r = 0; r < 6; ++r) f[r] = 0.0f;
	float b = 1.0f, f_sum = 0.0f;
	float log2_unscale = 0.0f;  // log2 of the product of the b's the rescaling divided by: Z = z * 2^log2_unscale
	uint32_t W = 0x3fffffffu, T48 = 31u, T49 = 31u;  // nothing but delimiters before the sequence
	// tantan.cpp:173-187 with forward_step :43-76
	const int shift = maxlen - len;
	for (int t = 0; t < maxlen; ++t) {
		const int i = t - shift;
		const uint32_t s_cur = i >= 0 ? (uint32_t)(seq[i] & 31) : 31u;
		if (i == 0) b = 1.0f;  // (f and f_sum are still exactly 0)
		const float* row = s_lr + s_cur * 32;
		const float b_old = b;
		float f_sum_new = 0.0f;
#pragma unroll
		for (int r = 0; r < 6; ++r) {
			const float tmp = __fadd_rn(__fmul_rn(f[r], f2f), __fmul_rn(b_old, d[r]));
			f[r] = __fmul_rn(tmp, row[(W >> (5 * r)) & 31u]);
			f_sum_new = __fadd_rn(f_sum_new, hsum8(f[r]));
		}
		f48 = __fmul_rn(__fadd_rn(__fmul_rn(f48, f2f), __fmul_rn(b_old, d48)), row[T48]);
		f_sum_new = __fadd_rn(f_sum_new, f48);
		f49 = __fmul_rn(__fadd_rn(__fmul_rn(f49, f2f), __fmul_rn(b_old, d49)), row[T49]);
		f_sum_new = __fadd_rn(f_sum_new, f49);
		b = __fadd_rn(__fmul_rn(b_old, b2b), __fmul_rn(f_sum, p_repeat_end));
		f_sum = f_sum_new;
		if ((i & 15) == 15) {  // (also at negative i: there it only renormalises b, f is 0)
			const float s = __fdiv_rn(1.0f, b);
			if (i >= 0) { log2_unscale += __log2f(b); if (k == 0) scs[i >> 4] = s; }
			b = __fmul_rn(b, s);
#pragma unroll
			for (int r = 0; r < 6; ++r) f[r] = __fmul_rn(f[r], s);
			f48 = __fmul_rn(f48, s); f49 = __fmul_rn(f49, s);
			f_sum = __fmul_rn(f_sum, s);
		}
		if (k == 0 && i >= 0) pbs[i] = b;
		// window of letter i + 1: lane k takes lane k-1's, lane 0 takes lane 7's shifted up with letter i at the bottom
		const uint32_t Wp = __shfl_sync(0xffffffffu, W, lead + ((k + 7) & 7));
		const uint32_t W7 = __shfl_sync(0xffffffffu, W, lead + 7);
		T49 = T48;
		T48 = (W7 >> 25) & 31u;  // letter i-48
		W = k ? Wp : (((Wp << 5) | s_cur) & 0x3fffffffu);
	}
	// z = b * b2b + sum(f, 50) * p_repeat_end, util/simd/vector.h:37-48 for the sum
	float acc = 0.0f;
#pragma unroll
	for (int r = 0; r < 6; ++r) acc = __fadd_rn(acc, f[r]);
	float fs = hsum8(acc);
	fs = __fadd_rn(fs, f48);
	fs = __fadd_rn(fs, f49);
	const float z = __fadd_rn(__fmul_rn(b, b2b), __fmul_rn(fs, p_repeat_end));
	if (k == 0 && slot < n_seq) {
		zinv_out[g.sid - s_begin] = __fdiv_rn(1.0f, z);
		// log2(Z / W_bg) < 3: no letter of this sequence can reach the masking threshold (see above); NaN compares false
		const float log2_ratio = __log2f(z) + log2_unscale - (float)(len + 1) * __log2f(b2b);
		need[slot] = (len > 0 && !(len <= TANTAN_CERT_MAX_LEN && log2_ratio < TANTAN_CERT_LOG2_RATIO)) ? 1 : 0;
	}
}

// Backward pass (tantan.cpp:192-212 with backward_step :78-111) of the sequences the forward kernel could not decide:
// `order` is the compacted longest-first list, *n_need its length (the grid covers the worst case, surplus warps leave).
static __global__ void __launch_bounds__(128) tantan_backward_kernel(int8_t* letters, const int64_t* __restrict__ limits, const uint32_t* __restrict__ order,
                                                                     const int* __restrict__ n_need, const DevParams* __restrict__ P, const float* __restrict__ pb,
                                                                     const float* __restrict__ scale, int64_t base, uint32_t s_begin, const float* __restrict__ zinv_in,
                                                                     uint32_t* hard /* bit per letter offset */) {
	__shared__ float s_lr[1024];
	for (int x = threadIdx.x; x < 1024; x += blockDim.x) s_lr[x] = P->tantan_lr[x];
	__syncthreads();
	const int lane = threadIdx.x & 31, k = lane & 7, lead = lane & 24;
	const TantanGroup g = tantan_group(limits, order, (uint32_t)*n_need);
	const int len = g.len, maxlen = g.maxlen;
	if (maxlen <= 0) return;  // warp-uniform
	const int8_t* seq = letters + g.beg;
	const float* pbs = pb + (g.beg - base);
	const float* scs = scale + ((g.beg - base) >> 4) + (g.sid - s_begin);
	float d[6];
#pragma unroll
	for (int r = 0; r < 6; ++r) d[r] = P->tantan_d[8 * r + k];
	const float d48 = P->tantan_d[48], d49 = P->tantan_d[49];
	const float b2b = P->tantan_b2b, f2f = P->tantan_f2f, p_repeat_end = P->tantan_p_repeat_end, p_mask = P->tantan_p_mask;
	const float zinv = len > 0 ? zinv_in[g.sid - s_begin] : 0.0f;
	// window of the last letter: the letters len-2-8r-k (delimiter code before the start), tails len-50 and len-51
	uint32_t W = 0;
#pragma unroll
	for (int r = 0; r < 6; ++r) { const int j = len - 2 - 8 * r - k; W |= (j >= 0 ? (uint32_t)(seq[j] & 31) : 31u) << (5 * r); }
	uint32_t T48 = len >= 50 ? (uint32_t)(seq[len - 50] & 31) : 31u, T49 = len >= 51 ? (uint32_t)(seq[len - 51] & 31) : 31u;
	uint32_t s_cur = len >= 1 ? (uint32_t)(seq[len - 1] & 31) : 31u;
	float f[6], f48 = p_repeat_end, f49 = p_repeat_end;
#pragma unroll
	for (int r = 0; r < 6; ++r) f[r] = p_repeat_end;
	float b = b2b;
	for (int t = 0; t < maxlen; ++t) {
		const int i = len - 1 - t;  // < 0: this sequence is done, the group idles on delimiters
		bool mask_it = false;
		if (k == 0 && i >= 0) {
			const float pf = __fsub_rn(1.0f, __fmul_rn(__fmul_rn(pbs[i], b), zinv));
			mask_it = pf >= p_mask;
		}
		if (i >= 0 && (i & 15) == 15) {
			const float s = scs[i >> 4];
			b = __fmul_rn(b, s);
#pragma unroll
			for (int r = 0; r < 6; ++r) f[r] = __fmul_rn(f[r], s);
			f48 = __fmul_rn(f48, s); f49 = __fmul_rn(f49, s);
		}
		const float* row = s_lr + s_cur * 32;
		const float vC = __fmul_rn(p_repeat_end, b);
		float tsum = 0.0f;
#pragma unroll
		for (int r = 0; r < 6; ++r) {
			const float vf = __fmul_rn(f[r], row[(W >> (5 * r)) & 31u]);
			const float vt = __fmul_rn(vf, d[r]);
			f[r] = __fadd_rn(__fmul_rn(vf, f2f), vC);
			tsum = __fadd_rn(tsum, hsum8(vt));
		}
		{
			const float vf = __fmul_rn(f48, row[T48]);
			tsum = __fadd_rn(tsum, __fmul_rn(vf, d48));
			f48 = __fadd_rn(__fmul_rn(vf, f2f), vC);
		}
		{
			const float vf = __fmul_rn(f49, row[T49]);
			tsum = __fadd_rn(tsum, __fmul_rn(vf, d49));
			f49 = __fadd_rn(__fmul_rn(vf, f2f), vC);
		}
		b = __fadd_rn(__fmul_rn(b2b, b), tsum);
		if (mask_it) {  // position i is never looked at again (its letter already sits in the registers of the steps that need it)
			letters[g.beg + i] = 23;  // value_traits.mask_char = MASK_LETTER
			const uint64_t p = (uint64_t)(g.beg + i);
			atomicOr(&hard[p >> 5], 1u << (p & 31));
		}
		// window of letter i - 1: lane k takes lane k+1's, lane 7 takes lane 0's shifted down with letter i-49 on top
		const uint32_t Wn = __shfl_sync(0xffffffffu, W, lead + ((k + 1) & 7));
		const uint32_t W0 = __shfl_sync(0xffffffffu, W, lead);
		W = k < 7 ? Wn : ((W0 >> 5) | (T48 << 25));
		s_cur = i >= 1 ? (W0 & 31u) : 31u;  // letter i-1 (lane 0's most recent one)
		T48 = T49;
		T49 = i >= 51 ? (uint32_t)(seq[i - 51] & 31) : 31u;
	}
}

static __global__ void popc_kernel(const uint32_t* __restrict__ bits, size_t w_begin, size_t w_end, unsigned long long* total) {
	const size_t w = w_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned n = w < w_end ? __popc(bits[w]) : 0u;
	for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
	if ((threadIdx.x & 31) == 0 && n) atomicAdd(total, (unsigned long long)n);
}

struct BitSet {
	const uint32_t* bits;
	__host__ __device__ bool operator()(const uint64_t& p) const { return (bits[p >> 5] >> (p & 31)) & 1u; }
};

// ---- motifs -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t motif_hash(uint64_t code) { return (uint32_t)((code * 0x9E3779B97F4A7C15ull) >> 48); }

#define MOTIF_TILE 2048
static __global__ void __launch_bounds__(256) motif_hit_kernel(const int8_t* __restrict__ letters, size_t raw_len, size_t p_begin, size_t p_end, const uint64_t* __restrict__ table,
                                                        uint32_t* cov, const int64_t* __restrict__ limits, uint32_t s_begin, uint32_t s_end,
                                                        uint32_t* seq_flag, uint32_t* seq_list, unsigned int* n_list) {
	__shared__ uint64_t s_tab[DMND_MOTIF_COUNT];
	__shared__ uint32_t s_pre[2048];  // 64 Kbit prefilter over motif_hash
	__shared__ uint8_t s_let[MOTIF_TILE + DMND_MOTIF_LEN];
	for (int x = threadIdx.x; x < 2048; x += blockDim.x) s_pre[x] = 0;
	for (int x = threadIdx.x; x < DMND_MOTIF_COUNT; x += blockDim.x) s_tab[x] = table[x];
	const size_t p0 = p_begin + (size_t)blockIdx.x * MOTIF_TILE;
	for (int x = threadIdx.x; x < MOTIF_TILE + DMND_MOTIF_LEN; x += blockDim.x) s_let[x] = p0 + x < raw_len ? (uint8_t)(letters[p0 + x] & 31) : (uint8_t)DMND_DELIMITER;
	__syncthreads();
	for (int x = threadIdx.x; x < DMND_MOTIF_COUNT; x += blockDim.x) { const uint32_t h = motif_hash(s_tab[x]); atomicOr(&s_pre[h >> 5], 1u << (h & 31)); }
	__syncthreads();
	for (int it = 0; it < MOTIF_TILE / 256; ++it) {
		const int o = it * 256 + threadIdx.x;
		const size_t p = p0 + o;
		if (p >= p_end) continue;
		// KmerIterator<8> (util/kmer/kmer.h:62-117): a k-mer exists where 8 consecutive letters are < TRUE_AA (the delimiter is 31)
		uint64_t code = 0;
		unsigned bad = 0;
#pragma unroll
		for (int q = 0; q < DMND_MOTIF_LEN; ++q) { const unsigned l = s_let[o + q]; bad |= l >= 20u; code = code * 20u + l; }
		if (bad) continue;
		const uint32_t h = motif_hash(code);
		if (!((s_pre[h >> 5] >> (h & 31)) & 1u)) continue;
		int lo = 0, hi = DMND_MOTIF_COUNT;
		while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_tab[mid] < code) lo = mid + 1; else hi = mid; }
		if (lo == DMND_MOTIF_COUNT || s_tab[lo] != code) continue;
		// pos.push_back(p, p + 8), masking.cpp:116-119: coverage bits
		const unsigned sh = (unsigned)(p & 31);
		atomicOr(&cov[p >> 5], 0xffu << sh);
		if (sh > 24) atomicOr(&cov[(p >> 5) + 1], 0xffu >> (32 - sh));
		uint32_t a = s_begin, b = s_end;  // sequence holding p
		while (b - a > 1) { const uint32_t mid = a + (b - a) / 2; if ((uint64_t)limits[mid] <= (uint64_t)p) a = mid; else b = mid; }
		const uint32_t bit = 1u << (a & 31);
		if (!(atomicOr(&seq_flag[a >> 5], bit) & bit)) seq_list[atomicAdd(n_list, 1u)] = a;
	}
}

__device__ __forceinline__ bool bit_at(const uint32_t* bits, size_t p) { return (bits[p >> 5] >> (p & 31)) & 1u; }
__device__ __forceinline__ void set_bits(uint32_t* bits, size_t b, size_t e) {  // [b, e)
	for (size_t p = b; p < e;) {
		const unsigned sh = (unsigned)(p & 31);
		const size_t n = (size_t)(32 - sh) < e - p ? (size_t)(32 - sh) : e - p;
		const uint32_t m = (n == 32 ? 0xffffffffu : ((1u << n) - 1u)) << sh;
		atomicOr(&bits[p >> 5], m);
		p += n;
	}
}

// One thread per sequence with at least one table hit: masking.cpp:120-129
static __global__ void motif_apply_kernel(const int64_t* __restrict__ limits, const uint32_t* __restrict__ seq_list, const unsigned int* __restrict__ n_list,
                                   const uint32_t* __restrict__ cov, uint32_t* soft, int max_motif_len) {
	const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= *n_list) return;
	const uint32_t sid = seq_list[t];
	const size_t beg = (size_t)limits[sid], end = (size_t)limits[sid + 1] - 1;
	const size_t len = end - beg;
	size_t n = 0;
	for (size_t p = beg; p < end; ++p) n += bit_at(cov, p);
	if ((double)(ptrdiff_t)n / (double)len >= 0.5) return;
	for (size_t p = beg; p < end;) {
		if (!bit_at(cov, p)) { ++p; continue; }
		size_t e = p;
		while (e < end && bit_at(cov, e)) ++e;
		if (e - p <= (size_t)max_motif_len) set_bits(soft, p, e);
		p = e;
	}
}

// clears the bits of [p_begin, p_end) in a bitmap other lanes share at the edges
static __global__ void clear_bits_kernel(uint32_t* bits, size_t p_begin, size_t p_end) {
	const size_t w = (p_begin >> 5) + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (w > ((p_end - 1) >> 5)) return;
	uint32_t m = 0xffffffffu;
	if (w == (p_begin >> 5)) m &= 0xffffffffu << (p_begin & 31);
	if (w == ((p_end - 1) >> 5)) m &= 0xffffffffu >> (31 - ((p_end - 1) & 31));
	if (m == 0xffffffffu) bits[w] = 0; else atomicAnd(&bits[w], ~m);
}

// MaskingTable::remove(template_len, add_bit_mask = true) after the query enumeration (masking/masking.cpp:96-107 through
// search/seed_array/enum_seeds.h:255-260, EnumCfg::mask_seeds of the query side, search/stage0.cpp:139-142): an entry [b, e)
// leaves SEED_MASK on [max(b - span + 1, 0), e) of its sequence, i.e. on every position j that sees a soft-masked letter of
// its own sequence within [j, j + span).  One thread per position; each writes only its own byte.
static __global__ void motif_seedmask_kernel(int8_t* q_letters, const uint32_t* __restrict__ soft, size_t p_begin, size_t p_end, int span) {
	const size_t j = p_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= p_end) return;
	const uint32_t w0 = soft[j >> 5], w1 = soft[(j >> 5) + 1];
	const uint32_t v = __funnelshift_r(w0, w1, (unsigned)(j & 31)) & ((1u << span) - 1u);
	if (v == 0) return;
	for (int k = 0; k < span; ++k) {
		if ((q_letters[j + k] & 31) == DMND_DELIMITER) return;
		if ((v >> k) & 1u) { q_letters[j] = (int8_t)(q_letters[j] | DMND_SEED_MASK); return; }
	}
}

}  // namespace dmnd_cuda
