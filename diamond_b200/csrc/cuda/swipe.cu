// swipe.cu -- banded affine-gap local alignment (DP::BandedSwipe::swipe, dp/swipe/banded_swipe.h:189-351) for sm_100a.
//
// Mapping.  The reference fills SIMD lanes with <= 32 targets of ONE query and walks band rows serially.  Here every
// DP problem (one DpTarget) gets one warp, and the lanes split the BAND: lane t owns R consecutive diagonals
// (band rows r = t*R .. t*R+R-1, diagonal d = d_begin + r) and keeps their H / hgap / vgap state in registers.
// Cell (r, c) [c = column index into the banded target range] depends on
//     (r, c-1) diagonal predecessor, (r+1, c-1) horizontal gap source, (r-1, c) vertical gap source,
// so with the wavefront time s = 2c + r all three are exactly one or two steps old: at step s the rows with
// r == s (mod 2) are active, every lane updates R/2 cells per step and exchanges ONE boundary value with a
// neighbour lane by warp shuffle (vgap from the lane above on even steps, hgap from the lane below on odd steps).
// The recurrence is (max,+): DPX three-input instructions (VIADDMNMX / VIMNMX3 with fused relu) carry it; it is
// not a contraction, so tensor cores do not apply.
//
// Semantics are the reference's int8/int16 lane semantics (every H, hgap, vgap floored at 0; cell_update.h:103-141),
// evaluated in exact int32, so no overflow cascade is needed.  Traceback mode stores the reference's two mask pairs
// per cell (gap: cur==vgap / cur==hgap, open: vgap'==open / hgap'==open) and a second kernel walks them
// (banded_swipe.h:127-187, banded_matrix.h:357-402, basic/hssp.cpp:260-290), one problem per thread.
#include "ctx.cuh"
#include "swipe16.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cstring>
#include <numeric>
#include <cstdlib>
#include <chrono>
#include <mutex>

namespace dmnd_cuda {

// The four reference masks of one cell, written as ORDER tests on the inputs instead of equality tests on the max
// results (all of e_in, f_in, open are >= 0 by the floor semantics, so the two forms are equivalent):
//   cur == vgap  <=>  f_in >= max(hd, e_in)            cur == hgap  <=>  e_in >= max(hd, f_in)
//   vgap' == open <=> open >= f_in - ge                 hgap' == open <=> open >= e_in - ge
// (ptxas 12.9 folds `max.s32.relu(x, y) == y` into the VIMNMX.RELU predicate output with the wrong polarity; the
// order form compiles to plain ISETP.GE and was verified in SASS.)
__device__ __forceinline__ uint8_t trace_flags(int hd, int e_in, int f_in, int open, int ge) {
	const int b0 = (f_in >= hd) & (f_in >= e_in), b1 = (e_in >= hd) & (e_in >= f_in);
	const int b2 = open >= f_in - ge, b3 = open >= e_in - ge;
	return (uint8_t)(b0 | (b1 << 1) | (b2 << 2) | (b3 << 3));
}
// Same four masks from differences instead of compares (h >= e_in, f_in and E', F' >= open always hold, so each
// difference is >= 0 and min(diff, 1) is the negated mask bit): 4 VIADDMNMX + 3 IMAD per cell.  Returns the nibble.
__device__ __forceinline__ unsigned trace_flags_arith(int h, int e_in, int f_in, int e_new, int f_new, int open) {
	const int n0 = __viaddmin_s32(h, -f_in, 1), n1 = __viaddmin_s32(h, -e_in, 1);
	const int n2 = __viaddmin_s32(f_new, -open, 1), n3 = __viaddmin_s32(e_new, -open, 1);
	return 15u ^ (unsigned)((n0 + 2 * n1) + 4 * (n2 + 2 * n3));
}

template<int R>
__device__ __forceinline__ void trace_store(uint8_t* p, const uint32_t* pk) {
	if (R == 2) *p = (uint8_t)pk[0];
	else if (R == 4) *reinterpret_cast<uint16_t*>(p) = (uint16_t)pk[0];
	else if (R == 8) *reinterpret_cast<uint32_t*>(p) = pk[0];
	else if (R == 16) *reinterpret_cast<uint2*>(p) = make_uint2(pk[0], pk[1]);
	else *reinterpret_cast<uint4*>(p) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}
template<int R, bool TRACE>
__global__ void __launch_bounds__(128) swipe_kernel(const SwipeArgs a, const DevParams* __restrict__ P) {
	__shared__ int8_t s_score[1024];
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_score[i] = P->score[i];
	__syncthreads();
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int go = P->gap_open + P->gap_extend, ge = P->gap_extend;
	const int r0 = lane * R;

	for (;;) {
		unsigned int w = 0;
		if (lane == 0) w = atomicAdd(a.work, 1u);
		w = __shfl_sync(FULL, w, 0);
		if (w >= a.n) break;
		const uint32_t pi = a.order[w];
		const dmnd_dp_problem pr = a.probs[pi];
		const ProbGeom g = geom(a, pr);
		int H[R], E[R], F[R], bestv[R], bestc[R];
#pragma unroll
		for (int k = 0; k < R; ++k) { H[k] = 0; E[k] = 0; F[k] = 0; bestv[k] = 0; bestc[k] = 0; }
		int best = 0;
		// lane t stores its R nibbles of one macro step as R/2 consecutive bytes: the warp writes 16*R contiguous bytes
		uint8_t* tr = TRACE ? a.trace + (a.trace_excl[a.order_pos0 + w] - a.trace_base) + (size_t)lane * (R / 2) : nullptr;
		constexpr int PKW = (R + 7) / 8;
		if (g.B > 0 && g.cols > 0) {
			const int ibase = g.j0 + g.d_begin;  // i = ibase + c + r
			const int nsteps = 2 * (g.cols - 1) + g.B;
			const int nmacro = (nsteps + 1) >> 1;
			for (int m = 0; m < nmacro; ++m) {
				uint32_t pk[PKW];
#pragma unroll
				for (int x = 0; x < PKW; ++x) pk[x] = 0;
				// ---- even step s = 2m : rows k = 0,2,.. ; column c = m - (r0 + k)/2
				{
					int f_up = __shfl_up_sync(FULL, F[R - 1], 1);
					if (lane == 0) f_up = 0;
#pragma unroll
					for (int k = 0; k < R; k += 2) {
						const int r = r0 + k, c = m - ((r0 + k) >> 1), i = ibase + c + r;
						if (r < g.B && (unsigned)c < (unsigned)g.cols && (unsigned)i < (unsigned)g.qlen) {
							const int sc = (int)s_score[((g.q[i] & 31) << 5) | (g.t[g.j0 + c] & 31)] + (int)g.cb[i];
							const int e_in = E[k + 1], f_in = k > 0 ? F[k > 0 ? k - 1 : 0] : f_up;
							const int hd = H[k] + sc;
							const int h = __vimax3_s32_relu(hd, e_in, f_in);
							const int open = max(h - go, 0);
							const int e = max(max(e_in - ge, 0), open), f = max(max(f_in - ge, 0), open);
							if (TRACE) {
								pk[k >> 3] |= (uint32_t)trace_flags(hd, e_in, f_in, open, ge) << ((k & 7) * 4);
								if (h > bestv[k]) { bestv[k] = h; bestc[k] = c; }
							}
							else best = max(best, h);
							H[k] = h; E[k] = e; F[k] = f;
						}
					}
				}
				// ---- odd step s = 2m + 1 : rows k = 1,3,.. ; column c = m - (r0 + k - 1)/2
				{
					int e_dn = __shfl_down_sync(FULL, E[0], 1);
					if (lane == 31) e_dn = 0;
#pragma unroll
					for (int k = 1; k < R; k += 2) {
						const int r = r0 + k, c = m - ((r0 + k - 1) >> 1), i = ibase + c + r;
						if (r < g.B && (unsigned)c < (unsigned)g.cols && (unsigned)i < (unsigned)g.qlen) {
							const int sc = (int)s_score[((g.q[i] & 31) << 5) | (g.t[g.j0 + c] & 31)] + (int)g.cb[i];
							const int e_in = k + 1 < R ? E[k + 1 < R ? k + 1 : 0] : e_dn, f_in = F[k - 1];
							const int hd = H[k] + sc;
							const int h = __vimax3_s32_relu(hd, e_in, f_in);
							const int open = max(h - go, 0);
							const int e = max(max(e_in - ge, 0), open), f = max(max(f_in - ge, 0), open);
							if (TRACE) {
								pk[k >> 3] |= (uint32_t)trace_flags(hd, e_in, f_in, open, ge) << ((k & 7) * 4);
								if (h > bestv[k]) { bestv[k] = h; bestc[k] = c; }
							}
							else best = max(best, h);
							H[k] = h; E[k] = e; F[k] = f;
						}
					}
				}
				if (TRACE) trace_store<R>(tr + (size_t)m * (16 * R), pk);
			}
		}
		if (TRACE) {
			// end cell = maximal H; ties: smallest column, then largest band row (banded_swipe.h:312-328, cell_update.h:43-46)
			int bv = 0, bc = 0, br = 0;
#pragma unroll
			for (int k = 0; k < R; ++k) {
				const int r = r0 + k;
				if (bestv[k] > bv || (bestv[k] == bv && bv > 0 && (bestc[k] < bc || (bestc[k] == bc && r > br)))) { bv = bestv[k]; bc = bestc[k]; br = r; }
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				const int ov = __shfl_xor_sync(FULL, bv, o), oc = __shfl_xor_sync(FULL, bc, o), orr = __shfl_xor_sync(FULL, br, o);
				if (ov > bv || (ov == bv && ov > 0 && (oc < bc || (oc == bc && orr > br)))) { bv = ov; bc = oc; br = orr; }
			}
			if (lane == 0) { a.score[pi] = bv; a.end_cell[2 * (size_t)pi] = bc; a.end_cell[2 * (size_t)pi + 1] = br; }
		}
		else {
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
			if (lane == 0) a.score[pi] = best;
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------------
// v2: query-profile kernel.  Same wavefront mapping and the same results as swipe_kernel, with the per-cell work cut to
// the recurrence itself:
//  * the warp first builds the problem's score profile in shared memory, prof[a][i + HALO] = S[a][q_i] + bias_i for the
//    27 target-letter rows a (row 26 = delimiter / out-of-alphabet) -- one LDS per cell replaces three global byte loads,
//    the 32x32 table lookup and the bias add;
//  * validity tests disappear: query positions outside [0, qlen) read -128 from the profile halo (HALO = 16 R columns on
//    each side covers every lane of every executed step), target positions are clamped to the delimiters at -1 / tlen,
//    and cells before the first column are zero by induction.  Cells past the query or target end can become positive
//    but only feed other out-of-matrix cells and stay strictly below a real cell's score, so neither the recurrence nor
//    the end-cell choice sees them (CPU emulation against the oracle: tests/test_swipe_emulation.py);
//  * each lane keeps the R/2 target letters it needs as premultiplied profile-row offsets in registers and shifts the
//    window by one letter per macro step; the band's lower edge is a per-lane row count kb (cells with k > kb are
//    predicated off, so their registers stay 0 exactly like the reference's hgap_[band] sentinel).
struct ProfArgs {
	int w4;              // profile row stride in bytes (multiple of 4), >= max(qlen) + 32 R + 4 of the launch
	unsigned int* overflow;  // set when S + bias does not fit int8: the host re-runs the call on the generic kernel
};

template<int R, bool TRACE>
__global__ void __launch_bounds__(128) swipe_prof_kernel(const SwipeArgs a, const DevParams* __restrict__ P, const ProfArgs pa) {
	extern __shared__ int8_t smem[];
	__shared__ int8_t s_score[1024];
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_score[i] = P->score[i];
	__syncthreads();
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int go = P->gap_open + P->gap_extend, ge = P->gap_extend;
	const int W4 = pa.w4;
	int8_t* prof = smem + (size_t)warp * 27 * W4;
	constexpr int HALO = 16 * R, U = R / 2;
	constexpr int PKW = (R + 7) / 8;

	for (;;) {
		unsigned int w = 0;
		if (lane == 0) w = atomicAdd(a.work, 1u);
		w = __shfl_sync(FULL, w, 0);
		if (w >= a.n) break;
		const uint32_t pi = a.order[w];
		const dmnd_dp_problem pr = a.probs[pi];
		const ProbGeom g = geom(a, pr);
		// per row: best cell as ONE key = score * 2^15 + (31743 - macro step + lane offset): an unsigned max keeps the highest
		// score and, among equals, the first column (1 IMAD on the FMA pipe + 1 VIMNMX instead of compare + select + max).
		// Fits: the profile lives in <= 200 KB of shared memory, so qlen < 7.6 k, macro steps < 31743 and scores < 2^17
		// (checked below; an out-of-range score raises the overflow flag and the call is redone on the generic kernel).
		int H[R], E[R], F[R];
		unsigned bestk[R];
#pragma unroll
		for (int k = 0; k < R; ++k) { H[k] = 0; E[k] = 0; F[k] = 0; bestk[k] = 0; }
		int best = 0;
		if (g.B > 0 && g.cols > 0) {
			// ---- profile
			const int W = g.qlen + 32 * R + 4;
			bool ok = true;
			__syncwarp();
			for (int idx = lane; idx < W; idx += 32) {
				const int i = idx - HALO;
				if (i >= 0 && i < g.qlen) {
					const int ql = g.q[i] & 31, cb = g.cb[i];
#pragma unroll 9
					for (int al = 0; al < 27; ++al) {
						int v = al < 26 ? (int)s_score[(al << 5) | ql] + cb : -128;
						if (al < 26 && (v > 127 || v < -127)) { ok = false; v = max(min(v, 127), -127); }
						prof[al * W4 + idx] = (int8_t)v;
					}
				}
				else {
#pragma unroll 9
					for (int al = 0; al < 27; ++al) prof[al * W4 + idx] = (int8_t)-128;
				}
			}
			if (!__all_sync(FULL, ok) && lane == 0) atomicExch(pa.overflow, 1u);
			__syncwarp();
			// ---- wavefront
			const int ibase = g.j0 + g.d_begin;
			const int nsteps = 2 * (g.cols - 1) + g.B, nmacro = (nsteps + 1) >> 1;
			const int m_lo = max(0, -ibase - 16 * R), m_hi = min(nmacro, g.qlen - ibase);
			const int kb = min(R - 1, g.B - 1 - lane * R);
			const int lofs = lane * U;
			auto trow_of = [&](int j) { const int jj = min(max(j, -1), g.tlen); return min((int)(g.t[jj] & 31), 26) * W4; };
			int trow[U];
#pragma unroll
			for (int u = 0; u < U; ++u) trow[u] = trow_of(g.j0 + m_lo - lofs - u);
			int I0 = ibase + m_lo + lofs + HALO;  // profile column of (even row u = 0); +u, +u+1 for the others
			uint8_t* tr = TRACE ? a.trace + (a.trace_excl[a.order_pos0 + w] - a.trace_base) + (size_t)lane * U : nullptr;
			for (int m = m_lo; m < m_hi; ++m) {
				const int tnext = trow_of(g.j0 + m + 1 - lofs);  // issued early: consumed after the two half steps
				const unsigned ckey = (unsigned)(31743 - m + lofs);
				uint32_t pk[PKW];
#pragma unroll
				for (int x = 0; x < PKW; ++x) pk[x] = 0;
				{
					int f_up = __shfl_up_sync(FULL, F[R - 1], 1);
					if (lane == 0) f_up = 0;
#pragma unroll
					for (int k = 0; k < R; k += 2) {
						{
							// rows past the band's lower edge (k > kb) are kept at H = E = 0 by a select on h: their F is never
							// read by a live row, so the reference's hgap_[band] = 0 sentinel holds without a branch
							const int u = k >> 1;
							const int sc = (int)prof[trow[u] + I0 + u];
							const int e_in = E[k + 1], f_in = k > 0 ? F[k > 0 ? k - 1 : 0] : f_up;
							const int hd = H[k] + sc;
							const int h = k <= kb ? __vimax3_s32_relu(hd, e_in, f_in) : 0;
							const int open = __viaddmax_s32_relu(h, -go, 0);  // relu(h - go), cell_update.h:129-131
							const int e_new = __viaddmax_s32_relu(e_in, -ge, open), f_new = __viaddmax_s32_relu(f_in, -ge, open);
							if (TRACE) {
								pk[k >> 3] |= trace_flags_arith(h, e_in, f_in, e_new, f_new, open) << ((k & 7) * 4);
								bestk[k] = max(bestk[k], (unsigned)h * 32768u + ckey);
							}
							else best = max(best, h);
							H[k] = h;
							E[k] = e_new;
							F[k] = f_new;
						}
					}
				}
				{
					int e_dn = __shfl_down_sync(FULL, E[0], 1);
					if (lane == 31) e_dn = 0;
#pragma unroll
					for (int k = 1; k < R; k += 2) {
						{
							const int u = k >> 1;
							const int sc = (int)prof[trow[u] + I0 + u + 1];
							const int e_in = k + 1 < R ? E[k + 1 < R ? k + 1 : 0] : e_dn, f_in = F[k - 1];
							const int hd = H[k] + sc;
							const int h = k <= kb ? __vimax3_s32_relu(hd, e_in, f_in) : 0;
							const int open = __viaddmax_s32_relu(h, -go, 0);  // relu(h - go), cell_update.h:129-131
							const int e_new = __viaddmax_s32_relu(e_in, -ge, open), f_new = __viaddmax_s32_relu(f_in, -ge, open);
							if (TRACE) {
								pk[k >> 3] |= trace_flags_arith(h, e_in, f_in, e_new, f_new, open) << ((k & 7) * 4);
								bestk[k] = max(bestk[k], (unsigned)h * 32768u + ckey);
							}
							else best = max(best, h);
							H[k] = h;
							E[k] = e_new;
							F[k] = f_new;
						}
					}
				}
				if (TRACE) trace_store<R>(tr + (size_t)m * (16 * R), pk);
#pragma unroll
				for (int u = U - 1; u > 0; --u) trow[u] = trow[u - 1];
				trow[0] = tnext;
				++I0;
			}
		}
		if (TRACE) {
			int bv = 0, bc = 0, br = 0;
			const int lofs = lane * U;
			bool wide = false;
#pragma unroll
			for (int k = 0; k < R; ++k) {
				const int r = lane * R + k;
				const int v = (int)(bestk[k] >> 15), c = 31743 + lofs - (int)(bestk[k] & 32767u) - lofs - (k >> 1);
				wide |= v >= 131072 - 256;
				if (v > bv || (v == bv && bv > 0 && (c < bc || (c == bc && r > br)))) { bv = v; bc = c; br = r; }
			}
			if (__any_sync(FULL, wide) && lane == 0) atomicExch(pa.overflow, 1u);
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				const int ov = __shfl_xor_sync(FULL, bv, o), oc = __shfl_xor_sync(FULL, bc, o), orr = __shfl_xor_sync(FULL, br, o);
				if (ov > bv || (ov == bv && ov > 0 && (oc < bc || (oc == bc && orr > br)))) { bv = ov; bc = oc; br = orr; }
			}
			if (lane == 0) { a.score[pi] = bv; a.end_cell[2 * (size_t)pi] = bc; a.end_cell[2 * (size_t)pi + 1] = br; }
		}
		else {
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
			if (lane == 0) a.score[pi] = best;
		}
	}
}

static __global__ void fill_score_results(const int32_t* score, dmnd_dp_result* res, uint32_t n) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	dmnd_dp_result r;
	r.score = score[i];
	r.q_begin = r.q_end = r.t_begin = r.t_end = 0;
	r.identities = r.mismatches = r.gap_openings = r.length = r.gaps = r.positives = 0;
	r.transcript_off = r.transcript_len = 0; r.status = 0;
	res[i] = r;
}

template<bool TRACE>
static void launch_bin(int R, const SwipeArgs& a, const DevParams* P, int grid, cudaStream_t st) {
	switch (R) {
	case 2: swipe_kernel<2, TRACE><<<grid, 128, 0, st>>>(a, P); break;
	case 4: swipe_kernel<4, TRACE><<<grid, 128, 0, st>>>(a, P); break;
	case 8: swipe_kernel<8, TRACE><<<grid, 128, 0, st>>>(a, P); break;
	case 16: swipe_kernel<16, TRACE><<<grid, 128, 0, st>>>(a, P); break;
	default: swipe_kernel<32, TRACE><<<grid, 128, 0, st>>>(a, P); break;
	}
}

// The kernels with dynamic shared memory above the 48 KB default get the opt-in limit ONCE per kernel (the attribute is a property of
// the function, shared by every lane thread of the process: setting it to each launch's own size let one lane lower it between another
// lane's set and launch -> "invalid argument" when the lanes' longest queries differed)
constexpr int DMND_SMEM_OPTIN = 227 * 1024;
template<typename K> static cudaError_t smem_optin(K kernel) {
	static std::mutex mtx;
	static std::vector<const void*> done;  // (kernels of one signature share this instantiation: keyed by the function itself)
	std::lock_guard<std::mutex> g(mtx);
	const void* key = reinterpret_cast<const void*>(kernel);
	if (std::find(done.begin(), done.end(), key) != done.end()) return cudaSuccess;
	cudaFuncAttributes fa;
	cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DMND_SMEM_OPTIN - (int)fa.sharedSizeBytes);  // the limit covers static + dynamic
	if (e == cudaSuccess) done.push_back(key);
	return e;
}

template<bool TRACE>
static int launch_prof_bin(int R, const SwipeArgs& a, const DevParams* P, const ProfArgs& pa, int grid, int threads, size_t smem, cudaStream_t st) {
#define DMND_LAUNCH_PROF(RR)                                                                                            \
	do {                                                                                                                 \
		if (smem > 40 * 1024) DMND_CUDA_CHECK(smem_optin(swipe_prof_kernel<RR, TRACE>)); /* the 48 KB default covers static + dynamic */ \
		swipe_prof_kernel<RR, TRACE><<<grid, threads, smem, st>>>(a, P, pa);                                             \
	} while (0)
	switch (R) {
	case 2: DMND_LAUNCH_PROF(2); break;
	case 4: DMND_LAUNCH_PROF(4); break;
	case 8: DMND_LAUNCH_PROF(8); break;
	case 16: DMND_LAUNCH_PROF(16); break;
	default: DMND_LAUNCH_PROF(32); break;
	}
#undef DMND_LAUNCH_PROF
	return 0;
}

// ---- statistics passes for problems above max_swipe_dp (no transcript requested) ---------------------------------------
// Reference: swipe_wrapper.cpp:89-96 bins such targets into the statistics kernels, dispatch_swipe :177-199 runs
// ForwardCell (ident, len) forwards and BackwardCell (mismatch, gapopen) over the reversed query x reversed target prefix
// (recompute_reversed :364-444); cell rules stat_cell.h:225-272 + cell_update.h:100-139.  These problems are rare (a band
// of some hundred diagonals over thousands of columns), so the kernel favours clarity: one CTA per problem, one thread
// per band row, wavefront time s = 2c + r with one barrier per step; a cell is {v, a, b} and ties copy the statistics of
// the max() argument exactly like set_max does.
struct SCell { int v, a, b; };
__device__ __forceinline__ void scell_max(SCell& x, const SCell& y) {
	x.v = max(x.v, y.v);
	if (x.v == y.v) { x.a = y.a; x.b = y.b; }
}
struct StatsArgs {
	const int8_t *q_letters, *q_bias, *r_letters;
	const int64_t *q_limits, *r_limits;
	const dmnd_dp_problem* probs;
	const uint32_t* order;
	uint32_t n;
	dmnd_dp_result* res;
};
struct StatsOut { int best, col, row, a, b; };

constexpr int WIDE_THREADS = 1024, WIDE_RPT = DMND_MAX_BAND / WIDE_THREADS;  // band rows per thread: r = tid + x * 1024

template<bool BACKWARD>
__device__ void stats_pass(const int8_t* __restrict__ q, const int8_t* __restrict__ cbs, int qlen, const int8_t* __restrict__ t, int tlen,
                           int d_begin, int d_end, const int8_t* __restrict__ score, int go, int ge, int* sm, StatsOut& out) {
	// BACKWARD reads q, cbs and t mirrored: q'[i] = q[qlen-1-i], t'[j] = t[tlen-1-j]
	const int band = d_end - d_begin, tid = (int)threadIdx.x;
	const int i1 = max(d_end - 1, 0), i0 = i1 + 1 - band, j0 = i1 - (d_end - 1);
	const int cols = min(qlen - 1 - d_begin, tlen - 1) + 1 - j0;
	constexpr int S = DMND_MAX_BAND + 1;
	int *hg_v = sm, *hg_a = sm + S, *hg_b = sm + 2 * S, *vg_v = sm + 3 * S, *vg_a = sm + 4 * S, *vg_b = sm + 5 * S;
	for (int k = tid; k <= band; k += WIDE_THREADS) { hg_v[k] = hg_a[k] = hg_b[k] = 0; vg_v[k] = vg_a[k] = vg_b[k] = 0; }
	__syncthreads();
	SCell h[WIDE_RPT];
	int my_best[WIDE_RPT], my_col[WIDE_RPT], my_a[WIDE_RPT], my_b[WIDE_RPT];
#pragma unroll
	for (int x = 0; x < WIDE_RPT; ++x) { h[x] = SCell{ 0, 0, 0 }; my_best[x] = my_col[x] = my_a[x] = my_b[x] = 0; }
	const int steps = cols > 0 ? 2 * (cols - 1) + band : 0;
	for (int s = 0; s < steps; ++s) {
#pragma unroll
		for (int x = 0; x < WIDE_RPT; ++x) {
			const int r = tid + x * WIDE_THREADS, c2 = s - r;
			if (r < band && c2 >= 0 && !(c2 & 1) && (c2 >> 1) < cols) {
				const int c = c2 >> 1, i = i0 + c + r, j = j0 + c;
				if (i >= 0 && i < qlen) {
					const int qi = BACKWARD ? qlen - 1 - i : i, tj = BACKWARD ? tlen - 1 - j : j;
					const int ql = q[qi] & DMND_LETTER_MASK, tl = t[tj] & DMND_LETTER_MASK;
					SCell hg{ hg_v[r + 1], hg_a[r + 1], hg_b[r + 1] };
					SCell vg{ 0, 0, 0 };
					if (r > 0 && i > 0) vg = SCell{ vg_v[r], vg_a[r], vg_b[r] };
					SCell cur = h[x];
					cur.v += (int)score[ql * 32 + tl] + (int)cbs[qi];
					const int id = ql == tl;
					if (!BACKWARD) { cur.a += id; cur.b += 1; hg.b += 1; vg.b += 1; }
					else cur.a += 1 - id;
					scell_max(cur, hg); scell_max(cur, vg);
					cur.v = max(cur.v, 0);
					if (cur.v > my_best[x]) { my_best[x] = cur.v; my_col[x] = c; my_a[x] = cur.a; my_b[x] = cur.b; }
					vg.v = max(vg.v - ge, 0); hg.v = max(hg.v - ge, 0);
					SCell open = cur;
					open.v = max(cur.v - go, 0);
					if (BACKWARD) open.b += 1;
					if (cur.v == 0) { cur.a = 0; cur.b = 0; }
					scell_max(hg, open); scell_max(vg, open);
					hg_v[r] = hg.v; hg_a[r] = hg.a; hg_b[r] = hg.b;
					vg_v[r + 1] = vg.v; vg_a[r + 1] = vg.a; vg_b[r + 1] = vg.b;
					h[x] = cur;
				}
			}
		}
		__syncthreads();  // rows r and r +- 1 never act in the same step (parity of s - r), so one barrier per step orders all exchanges
	}
	// end cell: highest value, then the first column, then the last band row (banded_swipe.h:321-326, VectorRowCounter)
	__syncthreads();
#pragma unroll
	for (int x = 0; x < WIDE_RPT; ++x) {
		const int r = tid + x * WIDE_THREADS;
		if (r < band) { hg_v[r] = my_best[x]; hg_a[r] = my_col[x]; hg_b[r] = my_a[x]; vg_v[r] = my_b[x]; }
	}
	__syncthreads();
	if (tid == 0) {
		StatsOut o{ 0, 0, 0, 0, 0 };
		for (int k = 0; k < band; ++k) {
			const int v = hg_v[k], c = hg_a[k];
			if (v > o.best || (v == o.best && v > 0 && c <= o.col)) { o.best = v; o.col = c; o.row = k; o.a = hg_b[k]; o.b = vg_v[k]; }
		}
		vg_a[0] = o.best; vg_a[1] = o.col; vg_a[2] = o.row; vg_a[3] = o.a; vg_a[4] = o.b;
	}
	__syncthreads();
	out.best = vg_a[0]; out.col = vg_a[1]; out.row = vg_a[2]; out.a = vg_a[3]; out.b = vg_a[4];
	__syncthreads();
}

#define DMND_STATS_SMEM (6 * (DMND_MAX_BAND + 1) * sizeof(int))
__global__ void __launch_bounds__(WIDE_THREADS) swipe_stats_kernel(const StatsArgs a, const DevParams* __restrict__ P) {
	extern __shared__ int sm[];
	const uint32_t pi = a.order[blockIdx.x];
	const dmnd_dp_problem pr = a.probs[pi];
	const int64_t qo = a.q_limits[pr.query], to = a.r_limits[pr.target];
	const int qlen = (int)(a.q_limits[pr.query + 1] - qo - 1), tlen = (int)(a.r_limits[pr.target + 1] - to - 1);
	const int8_t *q = a.q_letters + qo, *cbs = a.q_bias + qo, *t = a.r_letters + to;
	const int go = P->gap_open + P->gap_extend, ge = P->gap_extend;
	const int band = pr.d_end - pr.d_begin;
	dmnd_dp_result res;
	res.score = 0; res.q_begin = res.q_end = res.t_begin = res.t_end = 0;
	res.identities = res.mismatches = res.gap_openings = res.length = res.gaps = res.positives = 0;
	res.transcript_off = 0; res.transcript_len = 0; res.status = 0;
	StatsOut f, b;
	stats_pass<false>(q, cbs, qlen, t, tlen, pr.d_begin, pr.d_end, P->score, go, ge, sm, f);
	if (f.best > 0) {
		const int i1 = max(pr.d_end - 1, 0), i0 = i1 + 1 - band, j0 = i1 - (pr.d_end - 1);
		const int q_end = i0 + f.col + f.row + 1, t_end = j0 + f.col + 1;
		const int rd0 = -(pr.d_end - 1) + qlen - t_end, rd1 = -pr.d_begin + qlen - t_end + 1;  // Geo::rev_diag
		stats_pass<true>(q, cbs, qlen, t, t_end, rd0, rd1, P->score, go, ge, sm, b);
		if (b.best > 0) {
			const int ri1 = max(rd1 - 1, 0), ri0 = ri1 + 1 - band, rj0 = ri1 - (rd1 - 1);
			res.score = b.best;
			res.q_end = q_end; res.t_end = t_end;
			res.q_begin = qlen - (ri0 + b.col + b.row + 1);
			res.t_begin = t_end - (rj0 + b.col + 1);
			res.identities = f.a; res.length = f.b;
			res.mismatches = b.a; res.gap_openings = b.b;
			res.gaps = res.length - res.identities - res.mismatches;  // assign_stats, stat_cell.h:215-219
		}
	}
	if (threadIdx.x == 0) a.res[pi] = res;
}

// ---- bands wider than 1024 diagonals (chains with far-apart diagonals on very long sequences; rare) ---------------------
// Same recurrence, masks, end-cell rule and trace layout as the warp kernels (tile_rows() = 64 / 128 fixes the nibble
// addresses the walk kernel reads), evaluated like the statistics passes: one CTA per problem, thread t owns the band rows
// t, t + 1024, ..., wavefront time s = 2c + r, one barrier per step, H in registers, E / F exchanged through shared memory.
// The two nibbles of a trace byte belong to rows 2x and 2x + 1, which act in consecutive steps: the slice is zeroed
// beforehand and each cell ORs its nibble in (the barrier orders the two read-modify-writes).
#define DMND_WIDE_SMEM (2 * (DMND_MAX_BAND + 1) * sizeof(int))
template<bool TRACE>
__global__ void __launch_bounds__(WIDE_THREADS) swipe_wide_kernel(const SwipeArgs a, const DevParams* __restrict__ P) {
	extern __shared__ int sm[];
	const uint32_t w = blockIdx.x;
	const uint32_t pi = a.order[w];
	const dmnd_dp_problem pr = a.probs[pi];
	const ProbGeom g = geom(a, pr);
	const int tid = (int)threadIdx.x;
	const int go = P->gap_open + P->gap_extend, ge = P->gap_extend;
	constexpr int S = DMND_MAX_BAND + 1;
	int *hg = sm, *vg = sm + S;
	const int band = g.B, R = tile_rows(band);
	for (int k = tid; k <= band; k += WIDE_THREADS) { hg[k] = 0; vg[k] = 0; }
	__syncthreads();
	int h[WIDE_RPT], my_best[WIDE_RPT], my_col[WIDE_RPT];
#pragma unroll
	for (int x = 0; x < WIDE_RPT; ++x) { h[x] = 0; my_best[x] = 0; my_col[x] = 0; }
	uint8_t* tr = TRACE ? a.trace + (a.trace_excl[a.order_pos0 + w] - a.trace_base) : nullptr;
	const int ibase = g.j0 + g.d_begin;
	const int steps = (band > 0 && g.cols > 0) ? 2 * (g.cols - 1) + band : 0;
	for (int s = 0; s < steps; ++s) {
#pragma unroll
		for (int x = 0; x < WIDE_RPT; ++x) {
			const int r = tid + x * WIDE_THREADS, c2 = s - r;
			if (r < band && c2 >= 0 && !(c2 & 1) && (c2 >> 1) < g.cols) {
				const int c = c2 >> 1, i = ibase + c + r;
				if (i >= 0 && i < g.qlen) {
					const int sc = (int)P->score[((g.q[i] & 31) << 5) | (g.t[g.j0 + c] & 31)] + (int)g.cb[i];
					const int e_in = hg[r + 1], f_in = (r > 0 && i > 0) ? vg[r] : 0;
					const int hd = h[x] + sc;
					const int cur = max(max(max(hd, e_in), f_in), 0);
					const int open = max(cur - go, 0);
					const int e = max(max(e_in - ge, 0), open), f = max(max(f_in - ge, 0), open);
					if (TRACE) {
						const unsigned nib = trace_flags(hd, e_in, f_in, open, ge);
						const int m = c + (r >> 1), lane = r / R, k = r - lane * R;
						uint8_t* p = tr + ((size_t)m * 32 + lane) * (size_t)(R >> 1) + (k >> 1);
						*p = (uint8_t)(*p | (nib << ((k & 1) * 4)));
					}
					if (cur > my_best[x]) { my_best[x] = cur; my_col[x] = c; }
					hg[r] = e; vg[r + 1] = f;
					h[x] = cur;
				}
			}
		}
		__syncthreads();
	}
	__syncthreads();
#pragma unroll
	for (int x = 0; x < WIDE_RPT; ++x) {
		const int r = tid + x * WIDE_THREADS;
		if (r < band) { hg[r] = my_best[x]; vg[r] = my_col[x]; }
	}
	__syncthreads();
	if (tid == 0) {
		int bv = 0, bc = 0, br = 0;
		for (int k = 0; k < band; ++k) {
			const int v = hg[k], c = vg[k];
			if (v > bv || (v == bv && v > 0 && c <= bc)) { bv = v; bc = c; br = k; }
		}
		a.score[pi] = bv;
		if (TRACE) { a.end_cell[2 * (size_t)pi] = bc; a.end_cell[2 * (size_t)pi + 1] = br; }
	}
}

// ---- device-side preparation: geometry, launch group, order key, trace cost ---------------------------------------------
// Launch groups: 0..7 = packed 16-bit kernel (swipe16.cuh), register tile R = 4, 8, 12, 16 x {short, long query};
// 8..21 = int32 kernels, tile_rows() = 2..128 x {short, long query}; 22 = statistics passes.  Inside a group the problems are
// ordered by macro steps, longest first: the four problems of a quarter-warp-per-problem warp run in lock step, so they should
// be of one size, and the heaviest work starts first.
constexpr int NG = 23, G_STATS = 22, G_LEGACY = 8;
struct PrepOut {
	uint32_t* key;       // [n] group << 20 | (0xFFFFF - macro steps)
	uint32_t* idx;       // [n] problem index (sort payload)
	uint64_t* cost;      // [n] trace bytes (traceback) or cells (score only)
	uint64_t* tslen;     // [n] qlen + tlen (transcript capacity)
	unsigned int* hist;  // [NG] problems per group
	unsigned int* maxq;  // [NG] longest query per group (sizes the shared memory of the launch)
	unsigned long long* cost_hist;  // [NG] trace bytes per group, [NG] = sum of tslen: the host derives group bases and totals without a second sync;
	                                // [NG + 1] algorithmic cells of the call, [NG + 2] cells the kernels evaluate incl. the padding of the register tiles
	unsigned int* flag;  // error flag
};
__global__ void __launch_bounds__(256) prep_kernel(const dmnd_dp_problem* __restrict__ probs, uint32_t n, const int64_t* __restrict__ ql, uint32_t nq,
                                                    const int64_t* __restrict__ rl, uint32_t nr, int trace, int s16, PrepOut o) {
	__shared__ unsigned int s_hist[NG], s_maxq[NG];
	__shared__ unsigned long long s_cost[NG + 3];
	for (int x = threadIdx.x; x < NG + 3; x += blockDim.x) { if (x < NG) { s_hist[x] = 0; s_maxq[x] = 0; } s_cost[x] = 0; }
	__syncthreads();
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) {
		const dmnd_dp_problem p = probs[k];
		uint32_t key = 0; uint64_t cost = 0, tslen = 0;
		if (p.query >= nq || p.target >= nr) atomicMax(o.flag, 1u);
		else {
			const int qlen = (int)(ql[p.query + 1] - ql[p.query] - 1), tlen = (int)(rl[p.target + 1] - rl[p.target] - 1);
			const int B = p.d_end - p.d_begin;
			const int i1 = max(p.d_end - 1, 0), j0 = i1 - (p.d_end - 1);
			const int cols = min(qlen - 1 - p.d_begin, tlen - 1) + 1 - j0;
			if (B > DMND_MAX_BAND) atomicMax(o.flag, 2u);
			else {
				const bool live = B > 0 && cols > 0;
				const unsigned long long cells = live ? (unsigned long long)B * (unsigned long long)cols : 0ull;
				const unsigned long long nmacro = live ? (unsigned long long)((2 * (cols - 1) + B + 1) >> 1) : 0ull;
				// trace == 2: no transcript is wanted, so problems above max_swipe_dp take the statistics passes (no trace bytes)
				const bool stats = trace == 2 && cells > (unsigned long long)DMND_MAX_SWIPE_DP;
				int g; unsigned long long step_bytes, rows = (unsigned long long)B;
				if (stats) { g = G_STATS; step_bytes = 0; }
				else if (s16 && B <= S16_MAX_BAND && nmacro <= (unsigned long long)S16_MAX_MACRO && qlen <= S16_MAX_QLEN) {
					const int R = s16_rows(B);
					g = (R / 4 - 1) * 2 + ((qlen + 8 * R + 4) > 768 ? 1 : 0);
					step_bytes = (unsigned long long)s16_step_bytes(R); rows = (unsigned long long)(S16_LANES * R);
				}
				else {
					const int R = tile_rows(B);
					const int b = R == 2 ? 0 : R == 4 ? 1 : R == 8 ? 2 : R == 16 ? 3 : R == 32 ? 4 : R == 64 ? 5 : 6;
					g = G_LEGACY + b * 2 + ((qlen + 32 * R + 4) > 768 ? 1 : 0);
					step_bytes = 16ull * (unsigned long long)R; rows = 32ull * (unsigned long long)R;
				}
				key = ((uint32_t)g << 20) | (0xFFFFFu - (uint32_t)min(nmacro, 0xFFFFFull));
				cost = stats ? 0ull : trace ? (g < G_LEGACY ? s16_trace_bytes((int)(step_bytes / 4), nmacro) : nmacro * step_bytes) : cells;
				tslen = (uint64_t)qlen + (uint64_t)tlen;
				atomicAdd(&s_hist[g], 1u);
				atomicMax(&s_maxq[g], (unsigned)qlen);
				atomicAdd(&s_cost[g], cost);
				atomicAdd(&s_cost[NG], tslen);
				atomicAdd(&s_cost[NG + 1], cells);
				atomicAdd(&s_cost[NG + 2], nmacro * rows);
			}
		}
		o.key[k] = key; o.idx[k] = k; o.cost[k] = cost; o.tslen[k] = tslen;
	}
	__syncthreads();
	for (int x = threadIdx.x; x < NG + 3; x += blockDim.x) {
		if (x < NG && s_hist[x]) { atomicAdd(&o.hist[x], s_hist[x]); atomicMax(&o.maxq[x], s_maxq[x]); }
		if (s_cost[x]) atomicAdd(&o.cost_hist[x], s_cost[x]);
	}
}
__global__ void gather_cost_kernel(const uint32_t* __restrict__ order, const uint64_t* __restrict__ cost, uint32_t n, uint64_t* out) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) out[k] = cost[order[k]];
}

template<bool TRACE>
static int launch_s16_bin(int R, const SwipeArgs& a, const DevParams* P, const S16Args& sa, int grid, int threads, size_t smem, cudaStream_t st) {
#define DMND_LAUNCH_S16(RR)                                                                                             \
	do {                                                                                                                 \
		if (smem > 40 * 1024) DMND_CUDA_CHECK(smem_optin(swipe16_kernel<RR, TRACE>));                                     \
		swipe16_kernel<RR, TRACE><<<grid, threads, smem, st>>>(a, P, sa);                                                \
	} while (0)
	switch (R) {
	case 4: DMND_LAUNCH_S16(4); break;
	case 8: DMND_LAUNCH_S16(8); break;
	case 12: DMND_LAUNCH_S16(12); break;
	default: DMND_LAUNCH_S16(16); break;
	}
#undef DMND_LAUNCH_S16
	return 0;
}

int s16_table_build(dmnd_ctx* ctx) {
	DMND_CUDA_CHECK(cudaMalloc(&ctx->d_s16_table, (size_t)S16_TABLE_ENTRIES + 16));
	unsigned* d_bad = nullptr;
	DMND_CUDA_CHECK(cudaMalloc(&d_bad, sizeof(unsigned)));
	DMND_CUDA_CHECK(cudaMemset(d_bad, 0, sizeof(unsigned)));
	s16_table_kernel<<<(S16_TABLE_ENTRIES + 255) / 256, 256>>>(ctx->d_params, ctx->d_s16_table, d_bad);
	unsigned bad = 1;
	DMND_CUDA_CHECK(cudaMemcpy(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost));
	cudaFree(d_bad);
	ctx->s16_ok = bad == 0;  // a matrix with S + bias outside int8 never takes the packed kernel
	return 0;
}

// `d_problems` != nullptr: the problem list is already on the device (dmnd_hits_chain) and `problems` is ignored
int banded_swipe_impl(dmnd_ctx* ctx, const dmnd_block* query, const dmnd_block* ref, const dmnd_dp_problem* problems, const dmnd_dp_problem* d_problems, size_t n, int mode,
                      dmnd_dp_result* results, uint8_t* transcripts, size_t transcript_cap) {
	if (n == 0) return 0;
	if (n > 0xfffffff0ull) { set_error("dmnd_banded_swipe: too many problems in one call"); return 1; }
	const bool trace = mode == DMND_DP_TRACEBACK;
	cudaStream_t st = ctx->stream;
	const bool prof = getenv("DMND_PROFILE") != nullptr;
	auto tp = std::chrono::steady_clock::now();
	auto lap = [&](const char* what) {
		if (!prof) return;
		auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "[dmnd profile]     swipe %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tp).count());
		tp = now;
	};
	static const int RS[7] = { 2, 4, 8, 16, 32, 64, 128 };
	const unsigned nb = (unsigned)((n + 255) / 256);
	const bool force_generic = ctx->force_generic_dp || getenv("DMND_GENERIC_DP") != nullptr;
	const bool use_s16 = ctx->s16_ok && !ctx->force_int32_dp && !force_generic && getenv("DMND_INT32_DP") == nullptr;
	// ---- device buffers
	size_t sort_tmp = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 25, st);
	if (ctx->b_probs.ensure(n * sizeof(dmnd_dp_problem)) || ctx->b_order.ensure(n * sizeof(uint32_t)) || ctx->b_results.ensure(n * sizeof(dmnd_dp_result))
	    || ctx->b_work.ensure(n * sizeof(int32_t) * 3 + 256 * sizeof(unsigned int) + 8 + 64 * sizeof(unsigned long long))
	    || ctx->b_prep.ensure(n * (8 + 8 + 8 + 8 + 4 + 4 + 4) + 64) || ctx->b_cub.ensure(sort_tmp))
		return 1;
	int32_t* d_score = ctx->b_work.as<int32_t>();
	int32_t* d_end = d_score + n;
	unsigned int* d_counters = (unsigned int*)(d_end + 2 * n);  // [0..63] work counters, [64] error flag, [65] overflow of the int8 profile, [66] overflow of the packed kernel, [96..] hist, [128..] max qlen
	uint64_t* d_cost = ctx->b_prep.as<uint64_t>();
	uint64_t* d_tslen = d_cost + n;
	uint64_t* d_cum = d_tslen + n;    // exclusive prefix of cost in order sequence (n entries) -- reused as gather buffer
	uint64_t* d_tsoff = d_cum + n;    // exclusive prefix of tslen in problem order
	uint32_t* d_key = (uint32_t*)(d_tsoff + n);
	uint32_t* d_key2 = d_key + n;
	uint32_t* d_idx = d_key2 + n;
	if (!d_problems) {
		PhaseTimer t(ctx, PH_H2D);
		DMND_CUDA_CHECK(cudaMemcpyAsync(ctx->b_probs.p, problems, n * sizeof(dmnd_dp_problem), cudaMemcpyHostToDevice, st));
		t.stop();
		ctx->h2d_bytes += n * sizeof(dmnd_dp_problem);
	}
	const dmnd_dp_problem* dev_probs = d_problems ? d_problems : ctx->b_probs.as<dmnd_dp_problem>();
	lap("upload problems");
	PhaseTimer t_dp(ctx, trace ? PH_DP_TRACE : PH_DP_SCORE);
	DMND_CUDA_CHECK(cudaMemsetAsync(d_counters, 0, 256 * sizeof(unsigned int), st));
	unsigned long long* d_cost_hist = (unsigned long long*)(((uintptr_t)(d_counters + 256) + 7) & ~(uintptr_t)7);  // [0..NG-1] cost per group, [NG] sum of tslen
	DMND_CUDA_CHECK(cudaMemsetAsync(d_cost_hist, 0, 64 * sizeof(unsigned long long), st));
	PrepOut po{ d_key, d_idx, d_cost, d_tslen, d_counters + 96, d_counters + 128, d_cost_hist, d_counters + 64 };
	prep_kernel<<<nb, 256, 0, st>>>(dev_probs, (uint32_t)n, query->limits, query->nseq, ref->limits, ref->nseq, trace ? (transcripts ? 1 : 2) : 0, use_s16 ? 1 : 0, po);
	DMND_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ctx->b_cub.p, sort_tmp, d_key, d_key2, d_idx, ctx->b_order.as<uint32_t>(), n, 0, 25, st));
	unsigned int* hp = (unsigned int*)ctx->h_pinned;  // [0..31] hist, [32..63] max qlen, [64] flag, [65] profile overflow, [66] packed-kernel overflow
	DMND_CUDA_CHECK(cudaMemcpyAsync(hp, d_counters + 96, 32 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
	DMND_CUDA_CHECK(cudaMemcpyAsync(hp + 32, d_counters + 128, 32 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
	DMND_CUDA_CHECK(cudaMemcpyAsync(hp + 64, d_counters + 64, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
	uint64_t* hq = (uint64_t*)(hp + 128);  // [0..NG-1] cost per group, [NG] sum of tslen, then [32..] group bases
	DMND_CUDA_CHECK(cudaMemcpyAsync(hq, d_cost_hist, (NG + 3) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
	ctx->launches += 4;
	t_dp.mark_end();  // device time of the preparation; the host's decisions below are not device time
	DMND_CUDA_CHECK(stream_wait(ctx, st));
	t_dp.collect();
	if (hp[64] == 1) { set_error("dmnd_banded_swipe: sequence index out of range"); return 1; }
	if (hp[64] == 2) { set_error("dmnd_banded_swipe: band wider than 4096 diagonals is not supported by this build"); return 1; }
	(trace ? ctx->dp_cells_trace : ctx->dp_cells_score) += hq[NG + 1];
	ctx->dp_cells_padded += hq[NG + 2];
	size_t grp_begin[NG + 1];
	grp_begin[0] = 0;
	for (int g = 0; g < NG; ++g) grp_begin[g + 1] = grp_begin[g] + hp[g];
	t_dp.restart();
	lap("device prep");

	SwipeArgs a;
	a.q_letters = query->letters; a.q_bias = query->bias; a.r_letters = ref->letters; a.q_limits = query->limits; a.r_limits = ref->limits;
	a.probs = dev_probs;
	a.score = d_score; a.end_cell = d_end; a.trace = nullptr; a.trace_excl = nullptr; a.trace_base = 0; a.order_pos0 = 0;
	uint64_t ts_total = 0;
	// one DP launch over order[pos, e) of group g
	auto launch_dp = [&](int g, size_t pos, size_t e, bool tr_mode) -> int {
		const unsigned maxq = hp[32 + g];
		a.order = ctx->b_order.as<uint32_t>() + pos; a.n = (uint32_t)(e - pos); a.order_pos0 = (uint32_t)pos;
		int R = 0, warps = 0;
		if (g < G_LEGACY) {  // packed 16-bit kernel: shared score table + the queries of 4 problems per warp as 16-bit codes
			R = 4 * (g / 2 + 1);
			const int qstride = ((int)maxq + 8 * R + 4 + S16_TILE + 7) & ~7;
			const size_t tab = (size_t)((S16_TABLE_BYTES + 15) & ~15), per_warp = (size_t)4 * (size_t)qstride * 4;  // per problem: bias pairs (2 bytes) + codes (16 bit)
			warps = (int)std::min<size_t>(4, (((size_t)200 << 10) - tab) / per_warp);
			if (warps < 1) { set_error("dmnd_banded_swipe: query too long for the packed kernel"); return 1; }  // (prep_kernel routes those to the int32 kernels)
			const size_t smem = tab + per_warp * (size_t)warps;
			// 227 KB of shared memory per SM, 1 KB reserved per CTA: the 27 KB table + 16 queries of <= 300 letters make 4 CTAs = 16 warps per SM
			// (the register file holds exactly 16 warps of the 128-register kernels)
			const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(16 / warps, ((size_t)227 << 10) / (smem + 1024)));
			const int grid = (int)std::min<size_t>((e - pos + 4 * warps - 1) / (4 * warps), (size_t)ctx->sm_count * ctas_per_sm);
			S16Args sa{ ctx->d_s16_table, qstride, d_counters + 66 };
			if (tr_mode ? launch_s16_bin<true>(R, a, ctx->d_params, sa, grid, warps * 32, smem, st) : launch_s16_bin<false>(R, a, ctx->d_params, sa, grid, warps * 32, smem, st)) return 1;
		}
		else {
			R = RS[(g - G_LEGACY) >> 1];
			const int w4 = ((int)maxq + 32 * R + 4 + 3) & ~3;
			const size_t per_warp = (size_t)27 * (size_t)w4;
			warps = (int)std::min<size_t>(4, ((size_t)200 << 10) / per_warp);
			if (R > 32) {  // bands beyond 1024 diagonals: one CTA per problem
				if (tr_mode) swipe_wide_kernel<true><<<(unsigned)(e - pos), WIDE_THREADS, DMND_WIDE_SMEM, st>>>(a, ctx->d_params);
				else swipe_wide_kernel<false><<<(unsigned)(e - pos), WIDE_THREADS, DMND_WIDE_SMEM, st>>>(a, ctx->d_params);
			}
			else if (force_generic || warps == 0) {
				const int grid = (int)std::min<size_t>((e - pos + 3) / 4, (size_t)ctx->sm_count * 8);
				if (tr_mode) launch_bin<true>(R, a, ctx->d_params, grid, st); else launch_bin<false>(R, a, ctx->d_params, grid, st);
			}
			else {
				const size_t smem = per_warp * (size_t)warps;
				const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(16 / warps, ((size_t)220 << 10) / (smem + 2048)));
				const int grid = (int)std::min<size_t>((e - pos + warps - 1) / warps, (size_t)ctx->sm_count * ctas_per_sm);
				ProfArgs pa{ w4, d_counters + 65 };
				if (tr_mode ? launch_prof_bin<true>(R, a, ctx->d_params, pa, grid, warps * 32, smem, st) : launch_prof_bin<false>(R, a, ctx->d_params, pa, grid, warps * 32, smem, st)) return 1;
			}
		}
		++ctx->launches;
		if (cudaError_t le = cudaGetLastError()) {
			char msg[256];
			snprintf(msg, sizeof msg, "dmnd_banded_swipe: DP launch failed (%s): group %d, %zu problems, R %d, max query %u, %d warps", cudaGetErrorString(le), g, e - pos, R, maxq, warps);
			set_error(msg);
			return 1;
		}
		return 0;
	};

	if (!trace) {
		for (int g = 0; g < G_STATS; ++g) {
			const size_t pos = grp_begin[g], e = grp_begin[g + 1];
			if (e > pos) { a.work = d_counters + g; if (launch_dp(g, pos, e, false)) return 1; }
		}
		fill_score_results<<<nb, 256, 0, st>>>(d_score, ctx->b_results.as<dmnd_dp_result>(), (uint32_t)n);
		++ctx->launches;
		DMND_CUDA_CHECK(cudaGetLastError());
		t_dp.stop();
	}
	else {
		// exclusive prefix of the trace bytes along the order sequence; its host copy drives the slicing
		size_t tmp1 = 0, tmp2 = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, tmp1, d_cum, d_cum, n, st);
		cub::DeviceScan::ExclusiveSum(nullptr, tmp2, d_tslen, d_tsoff, n, st);
		if (ctx->b_cub.ensure(std::max(tmp1, tmp2)) || ctx->b_trace_off.ensure((n + 1) * sizeof(uint64_t))) return 1;
		uint64_t* d_excl = ctx->b_trace_off.as<uint64_t>();
		gather_cost_kernel<<<nb, 256, 0, st>>>(ctx->b_order.as<uint32_t>(), d_cost, (uint32_t)n, d_cum);
		DMND_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ctx->b_cub.p, tmp1, d_cum, d_excl, n, st));
		ctx->launches += 2;
		if (transcripts) {
			DMND_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ctx->b_cub.p, tmp2, d_tslen, d_tsoff, n, st));
			++ctx->launches;
		}
		// group bases and totals come from the per-group cost sums of the prep kernel (already on the host): the scans
		// above stay on the device, no second synchronisation
		uint64_t* gbase = hq + 32;  // [0..NG] trace bytes before group g
		gbase[0] = 0;
		for (int g = 0; g < NG; ++g) gbase[g + 1] = gbase[g] + hq[g];
		const uint64_t trace_total = gbase[NG];  // (the statistics group carries no trace)
		ts_total = hq[NG];
		if (transcripts && ts_total > transcript_cap) { set_error("dmnd_banded_swipe: transcript buffer too small (need sum(qlen+tlen))"); return 1; }
		if (ts_total > 0xffffffffull) { set_error("dmnd_banded_swipe: transcript buffer exceeds 4 GiB in one call"); return 1; }
		if (transcripts && ctx->b_tr.ensure(ts_total + 16)) return 1;
		size_t free_b = 0, total_b = 0;
		DMND_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
		uint64_t budget = std::max<uint64_t>((uint64_t)1 << 30, (uint64_t)((free_b + ctx->b_trace.cap) / 5) * 2);  // <= 40 % of what is free
		if (const char* ev = getenv("DMND_TRACE_BUDGET")) budget = std::max<uint64_t>(1024, strtoull(ev, nullptr, 10));  // tests: force slicing
		if (ctx->b_trace.ensure((size_t)std::min<uint64_t>(trace_total, budget) + 64)) return 1;
		// the whole prefix is only needed on the host when the trace has to be cut into slices
		const bool sliced = trace_total > budget;
		std::vector<uint64_t>& excl = ctx->h_excl;
		if (sliced) {
			excl.resize(n + 1);
			DMND_CUDA_CHECK(cudaMemcpyAsync(excl.data(), d_excl, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
			DMND_CUDA_CHECK(stream_wait(ctx, st));
			excl[n] = trace_total;
		}
		lap("trace prefix");
		for (int g = 0; g < G_STATS; ++g) {
			size_t pos = grp_begin[g];
			const size_t gend = grp_begin[g + 1];
			while (pos < gend) {
				// slice [pos, e) of this group whose trace fits the budget (at least one problem); launches are stream-ordered,
				// so the arena can be reused by the next slice without a host synchronisation
				size_t e = gend;
				uint64_t base = gbase[g], bytes = gbase[g + 1] - gbase[g];
				if (sliced) {
					e = (size_t)(std::upper_bound(excl.begin() + pos + 1, excl.begin() + gend + 1, excl[pos] + budget) - excl.begin()) - 1;
					e = std::max(e, pos + 1);
					base = excl[pos]; bytes = excl[e] - excl[pos];
				}
				if (ctx->b_trace.ensure((size_t)bytes + 64)) return 1;
				DMND_CUDA_CHECK(cudaMemsetAsync(d_counters, 0, 64 * sizeof(unsigned int), st));
				a.work = d_counters;
				a.trace = ctx->b_trace.as<uint8_t>(); a.trace_excl = d_excl; a.trace_base = base;
				if (g >= G_LEGACY && RS[(g - G_LEGACY) >> 1] > 32) DMND_CUDA_CHECK(cudaMemsetAsync(ctx->b_trace.p, 0, (size_t)bytes, st));  // the wide kernel ORs nibbles in
				if (launch_dp(g, pos, e, true)) return 1;
				WalkArgs wa;
				wa.q_letters = a.q_letters; wa.q_bias = a.q_bias; wa.r_letters = a.r_letters; wa.q_limits = a.q_limits; wa.r_limits = a.r_limits;
				wa.probs = a.probs; wa.order = a.order; wa.n = a.n; wa.score = d_score; wa.end_cell = d_end;
				wa.trace = a.trace; wa.trace_excl = d_excl; wa.trace_base = a.trace_base; wa.order_pos0 = a.order_pos0;
				wa.res = ctx->b_results.as<dmnd_dp_result>();
				wa.s16 = g < G_LEGACY ? 1 : 0;
				wa.transcripts = transcripts ? ctx->b_tr.as<uint8_t>() : nullptr;
				wa.transcript_off = transcripts ? d_tsoff : nullptr;
				{
					const unsigned wg = (unsigned)((a.n + 127) / 128);
					switch (g < G_LEGACY ? 4 * (g / 2 + 1) : 0) {
					case 4: walk16_kernel<4><<<wg, 128, 0, st>>>(wa, ctx->d_params); break;
					case 8: walk16_kernel<8><<<wg, 128, 0, st>>>(wa, ctx->d_params); break;
					case 12: walk16_kernel<12><<<wg, 128, 0, st>>>(wa, ctx->d_params); break;
					case 16: walk16_kernel<16><<<wg, 128, 0, st>>>(wa, ctx->d_params); break;
					default: walk_kernel<<<wg, 128, 0, st>>>(wa, ctx->d_params); break;
					}
				}
				++ctx->launches;
				DMND_CUDA_CHECK(cudaGetLastError());
				pos = e;
			}
		}
		if (hp[G_STATS]) {  // statistics passes, one CTA per problem
			StatsArgs sa;
			sa.q_letters = a.q_letters; sa.q_bias = a.q_bias; sa.r_letters = a.r_letters; sa.q_limits = a.q_limits; sa.r_limits = a.r_limits;
			sa.probs = a.probs; sa.order = ctx->b_order.as<uint32_t>() + grp_begin[G_STATS]; sa.n = hp[G_STATS];
			sa.res = ctx->b_results.as<dmnd_dp_result>();
			DMND_CUDA_CHECK(cudaFuncSetAttribute(swipe_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DMND_STATS_SMEM));
			swipe_stats_kernel<<<hp[G_STATS], WIDE_THREADS, DMND_STATS_SMEM, st>>>(sa, ctx->d_params);
			++ctx->launches;
			DMND_CUDA_CHECK(cudaGetLastError());
		}
		t_dp.stop();
	}
	lap("kernels");
	{
		PhaseTimer t(ctx, PH_D2H);
		DMND_CUDA_CHECK(cudaMemcpyAsync(results, ctx->b_results.p, n * sizeof(dmnd_dp_result), cudaMemcpyDeviceToHost, st));
		if (trace && transcripts && ts_total) DMND_CUDA_CHECK(cudaMemcpyAsync(transcripts, ctx->b_tr.p, ts_total, cudaMemcpyDeviceToHost, st));
		t.stop();
		ctx->d2h_bytes += n * sizeof(dmnd_dp_result) + ((trace && transcripts) ? ts_total : 0);
	}
	DMND_CUDA_CHECK(cudaMemcpyAsync(hp + 65, d_counters + 65, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
	DMND_CUDA_CHECK(stream_wait(ctx, st));
	lap("download results");
	if (hp[66] != 0 && use_s16) {
		// a score near the int16 range or a Hauser bias outside the table: the reference's int16 -> int32 cascade
		// (banded_swipe.h:337,347), here for the whole call on the exact int32 kernels
		++ctx->dp_overflows;
		ctx->force_int32_dp = true;
		const int rc = banded_swipe_impl(ctx, query, ref, problems, d_problems, n, mode, results, transcripts, transcript_cap);
		ctx->force_int32_dp = false;
		return rc;
	}
	if (hp[65] != 0 && !force_generic) {
		// S + bias left the int8 range of the shared-memory profile somewhere: redo the whole call on the generic kernel
		ctx->force_generic_dp = true;
		const int rc = banded_swipe_impl(ctx, query, ref, problems, d_problems, n, mode, results, transcripts, transcript_cap);
		ctx->force_generic_dp = false;
		return rc;
	}
	if (trace && !getenv("DMND_NO_TRACE_CHECK"))
		for (size_t k = 0; k < n; ++k)
			if (results[k].status == 2) { set_error("dmnd_banded_swipe: Traceback error."); return 1; }
	return 0;
}

}  // namespace dmnd_cuda
