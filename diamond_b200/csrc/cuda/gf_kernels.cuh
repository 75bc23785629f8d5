// gf_kernels.cuh -- device code of dmnd_hits_gapped_filter (modes from --sensitive upwards) for sm_100a.  Kernel only: the
// launch is in seed.cu; tests/emu_gf.cpp compiles THIS file for the CPU behind tests/emu_cuda.h and checks it against the oracle.
//
//   Extension::gapped_filter     align/gapped_filter.cpp:33-63     per hit: 64-diagonal scan (window 100) > cutoff 1, then
//                                                                  128-diagonal scan (window gapped_filter_window) > cutoff 2
//   DP::scan_diags64/128         dp/scan_diags.cpp:30-275          running score per diagonal, floored at 0, saturating at 255
//   DP::make_profile8            dp/score_profile.cpp:32-65        sat8(matrix8[l][query[i]] + bias[i]) for l < 20, -1 in the padding
//   DP::diag_alignment           dp/scan_diags.cpp:277-297         gapped combination of the per-diagonal maxima
//
// One warp per hit; lane k owns the diagonals d + k, d + k + 32, ... of the band (2 or 4 running scores in registers).  All
// lanes walk the same target columns (one broadcast byte load per column), each reads its own query letter + bias (consecutive
// lanes -> consecutive bytes).  Integer work on L1-resident bytes: 6 ops per cell, no HBM traffic to speak of.
#pragma once
#include "dev_params.h"

namespace dmnd_cuda {

template<int BAND>
__device__ __forceinline__ int gf_scan_align(const int8_t* __restrict__ s_score, int* s_diag, const DevParams* __restrict__ P, const int8_t* __restrict__ qs,
                                             const int8_t* __restrict__ cb, int qlen, const int8_t* __restrict__ ss, int slen, int hi, int hj, int window, int lane) {
	constexpr int PER = BAND / 32;
	const int diag = hi - hj;
	const int d = max(diag - BAND / 2, -(slen - 1));
	const int jw0 = max(hj - window, 0), jw1 = min(hj + window, slen);
	const int j0 = max(jw0, -(d + BAND - 1)), j1 = min(qlen - d, jw1);
	int v[PER], mx[PER];
#pragma unroll
	for (int u = 0; u < PER; ++u) { v[u] = 0; mx[u] = 0; }
	for (int j = j0; j < j1; ++j) {
		const int l = ss[j] & 31;
		const int i = d + j;
#pragma unroll
		for (int u = 0; u < PER; ++u) {
			const int qi = i + lane + 32 * u;
			int pv = -1;  // LongScoreProfile padding
			if (qi >= 0 && qi < qlen) {
				pv = (int)s_score[(l << 5) | (qs[qi] & 31)];
				if (l < 20) pv = min(max(pv + (int)cb[qi], -128), 127);  // ScoreVector<int8_t, 0> += bias: saturating int8 add
			}
			const int x = min(max(v[u] + pv, 0), 255);
			v[u] = x;
			mx[u] = max(mx[u], x);
		}
	}
#pragma unroll
	for (int u = 0; u < PER; ++u) s_diag[lane + 32 * u] = mx[u];
	__syncwarp();
	int best = 0;
	if (lane == 0) {  // DP::diag_alignment
		int best_gap = -P->gap_open, dd = -1;
		for (int i = 0; i < BAND; ++i) {
			const int s = s_diag[i];
			if (s < P->gapped_filter_diag_score) continue;
			const int gap_score = -P->gap_extend * (i - dd) + best_gap;
			int n = s;
			if (gap_score + s > best) best = n = gap_score + s;
			if (s > best) best = n = s;
			const int open_score = -P->gap_open + n;
			if (open_score > gap_score) { best_gap = open_score; dd = i; }
		}
	}
	__syncwarp();
	return __shfl_sync(0xffffffffu, best, 0);
}

static __global__ void __launch_bounds__(128) gapped_filter_kernel(const int8_t* __restrict__ q_letters, const int8_t* __restrict__ q_bias, const int64_t* __restrict__ q_limits,
                                                                   const int8_t* __restrict__ r_letters, const int64_t* __restrict__ r_limits, uint32_t nr,
                                                                   const dmnd_hit* __restrict__ hits, size_t n, const DevParams* __restrict__ P, uint8_t* pass) {
	__shared__ int8_t s_score[1024];
	__shared__ int s_diag[4][128];
	for (int x = threadIdx.x; x < 1024; x += blockDim.x) s_score[x] = P->score[x];
	__syncthreads();
	const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (w >= n) return;  // warp-uniform
	const dmnd_hit hit = hits[w];
	const uint64_t sloc = hit.subject_score & 0xFFFFFFFFFFFFull;
	uint32_t a = 0, b = nr;
	while (b - a > 1) { const uint32_t mid = a + (b - a) / 2; if ((uint64_t)r_limits[mid] <= sloc) a = mid; else b = mid; }
	const int64_t qo = q_limits[hit.query], ro = r_limits[a];
	const int qlen = (int)(q_limits[hit.query + 1] - qo - 1), slen = (int)(r_limits[a + 1] - ro - 1);
	const int8_t *qs = q_letters + qo, *cb = q_bias + qo, *ss = r_letters + ro;
	const int hi = hit.seed_offset, hj = (int)((int64_t)sloc - ro);
	int* diag = s_diag[(threadIdx.x >> 5) & 3];
	// the cutoffs are read at query_profile->length() = the length of the query's FIRST context, whatever frame the hit lies in;
	// translated queries shorter than MIN_STAGE2_QLEN = 100 pass after the first scan (align/gapped_filter.cpp:44-55)
	const bool translated = P->query_contexts > 1;
	const uint32_t q0 = translated ? hit.query / (uint32_t)P->query_contexts * (uint32_t)P->query_contexts : hit.query;
	const int qlen0 = translated ? (int)(q_limits[q0 + 1] - q_limits[q0] - 1) : qlen;
	const int bq = 32 - __clz((unsigned)qlen0), bs = 32 - __clz((unsigned)slen);
	bool ok = false;
	const int f1 = gf_scan_align<64>(s_score, diag, P, qs, cb, qlen, ss, slen, hi, hj, 100, lane);
	if (qlen0 > 0 && f1 > (int)P->gapped_cutoff1[bq][bs]) {  // warp-uniform (f1 is broadcast)
		if (translated && qlen0 < 100) ok = true;
		else {
			const int f2 = gf_scan_align<128>(s_score, diag, P, qs, cb, qlen, ss, slen, hi, hj, P->gapped_filter_window, lane);
			ok = f2 > (int)P->gapped_cutoff2[bq][bs];
		}
	}
	if (lane == 0) pass[w] = ok ? 1 : 0;
}

}  // namespace dmnd_cuda
