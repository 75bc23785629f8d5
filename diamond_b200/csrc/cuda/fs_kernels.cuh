// fs_kernels.cuh -- 3-frame banded DP with frameshifts (banded_3frame_swipe, dp/swipe/banded_3frame_swipe.cpp:392-520; cell update
// dp/swipe/swipe.h:57-83) and its traceback (:338-390, TracebackIterator :152-250) for sm_100a.  Device code only;
// tests/emu_fs.cpp compiles THIS file for the CPU behind tests/emu_cuda.h and checks it against the oracle.
//
// Mapping.  One problem (one DpTarget of one strand of a DNA query) per warp, as in swipe.cu: lane t owns the R consecutive band
// rows r = t*R .. t*R+R-1 (band row r at column c is codon i = i0 + c + r of the strand); a "row" now carries THREE cells, one
// per reading frame f (codon start x = 3i + f on the strand), with their H / hgap / vgap in registers.  In nucleotide terms
//     H(x, j) = max(0, H(x-3, j-1) + s, H(x-4, j-1) + s - F, H(x-2, j-1) + s - F, hgap(x, j-1), vgap(x-3, j))
// so cell (r, f) at column c reads, from column c - 1: its own row (diagonal, and the neighbouring frames f -+ 1 for the two
// shifts), frame 2 of row r - 1 (the x - 4 of frame 0) and frame 0 of row r + 1 (the x - 2 of frame 2) -- and the gaps exactly as in
// the one-frame recurrence, per frame.  With the wavefront time s = 2c + r every operand is at least one step old; row r - 1
// has by then moved on to column c, so every row keeps its previous frame-2 value one step longer (H2old).  Lanes exchange four
// values per half step by shuffle (even: vgap x 3 + H2old from the lane above; odd: hgap x 3 + H(frame 0) from the lane below).
//
// Semantics: the reference's int32 instantiation (scores floored at 0, gaps not), which has the same H values as its int16 SIMD
// pass; the caller applies that pass's batch-wide band (host/legacy.inc).  Traceback keeps the whole score matrix in the
// reference's layout -- one column of W = 3*band + 1 ints per target letter after a zero column, band index p = 3r + f, the
// last entry of a column stays 0 -- and a second kernel walks it with the reference's comparisons in the reference's order
// (diagonal, forward shift, reverse shift, then the gap search: horizontal before vertical at equal length).
#pragma once
#include "dev_params.h"

namespace dmnd_cuda {

struct FsArgs {
	const int8_t *q_letters, *r_letters;
	const int64_t *q_limits, *r_limits;
	const dmnd_dp_problem* probs;
	const uint32_t* order;        // the problems of this launch (one register-tile class of one memory slice)
	uint32_t n;
	int frame_shift;
	int32_t* score;               // [problem]
	int32_t* max_col;             // [problem] first column that reaches the score (traceback only)
	int32_t* matrix;              // score matrices of the launch (traceback only), zero-filled by the caller
	const uint64_t* matrix_off;   // [problem] offset (ints) of the problem's matrix from the start of the CALL's matrices
	uint64_t matrix_base;         // offset of the first problem of this launch
	unsigned int* work;
};

struct FsGeom {
	const int8_t* q[3];
	int ql[3];
	const int8_t* t;
	int tlen, d_begin, d_end, B, i0, pos0, ncol, dna_len;
};

__device__ __forceinline__ FsGeom fs_geom(const int8_t* q_letters, const int64_t* q_limits, const int8_t* r_letters, const int64_t* r_limits, const dmnd_dp_problem& pr) {
	FsGeom g;
#pragma unroll
	for (int f = 0; f < 3; ++f) {
		const int64_t o = q_limits[pr.query + f];
		g.q[f] = q_letters + o; g.ql[f] = (int)(q_limits[pr.query + f + 1] - o - 1);
	}
	g.dna_len = g.ql[0] + g.ql[1] + g.ql[2] + 2;  // frame f holds (dna_len - f) / 3 codons
	const int64_t to = r_limits[pr.target];
	g.t = r_letters + to; g.tlen = (int)(r_limits[pr.target + 1] - to - 1);
	g.d_begin = pr.d_begin; g.d_end = pr.d_end; g.B = pr.d_end - pr.d_begin;
	const int i1 = max(pr.d_end - 1, 0);
	g.i0 = i1 + 1 - g.B; g.pos0 = i1 - (pr.d_end - 1);  // banded_3frame_swipe.cpp:410-421, target_iterator.h:68-92
	// columns: until the target ends or the band's first row passes the query end (:447-449)
	g.ncol = (g.B > 0 && g.ql[0] > 0) ? max(min(g.tlen - g.pos0, g.ql[0] - g.i0), 0) : 0;
	return g;
}
// ints of one problem's score matrix: (columns + 1) x (3 * band + 1)
__host__ __device__ __forceinline__ unsigned long long fs_matrix_ints(int B, int ncol) { return (unsigned long long)(ncol + 1) * (unsigned long long)(3 * B + 1); }
__host__ __device__ __forceinline__ int fs_tile_rows(int B) { return B <= 64 ? 2 : B <= 128 ? 4 : B <= 256 ? 8 : B <= 512 ? 16 : 32; }
#define DMND_FS_MAX_BAND 1024

template<int R, bool TRACE>
__global__ void __launch_bounds__(128) fs_swipe_kernel(const FsArgs a, const DevParams* __restrict__ P) {
	__shared__ int8_t s_score[1024];
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_score[i] = P->score[i];
	__syncthreads();
	const unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int go = P->gap_open + P->gap_extend, ge = P->gap_extend, FS = a.frame_shift;
	const int r0 = lane * R;

	for (;;) {
		unsigned int w = 0;
		if (lane == 0) w = atomicAdd(a.work, 1u);
		w = __shfl_sync(FULL, w, 0);
		if (w >= a.n) break;
		const uint32_t pi = a.order[w];
		const dmnd_dp_problem pr = a.probs[pi];
		const FsGeom g = fs_geom(a.q_letters, a.q_limits, a.r_letters, a.r_limits, pr);
		int H[R][3], E[R][3], F[R][3], H2old[R], bestv[R], bestc[R];
#pragma unroll
		for (int k = 0; k < R; ++k) {
#pragma unroll
			for (int f = 0; f < 3; ++f) { H[k][f] = 0; E[k][f] = 0; F[k][f] = 0; }
			H2old[k] = 0; bestv[k] = 0; bestc[k] = 0;
		}
		int32_t* S = TRACE ? a.matrix + (a.matrix_off[pi] - a.matrix_base) : nullptr;
		const int W = 3 * g.B + 1;
		// one cell triple: row k of this lane at column c; up = (vgap[3], H2old) of row r - 1, dn = (hgap[3], H frame 0) of row r + 1
		auto row_update = [&](int k, int c, const int* upF, int upH2, const int* dnE, int dnH0) {
			const int r = r0 + k, i = g.i0 + c + r;
			if (!(r < g.B && (unsigned)c < (unsigned)g.ncol && (unsigned)i < (unsigned)g.ql[0])) return;
			const int tl = g.t[g.pos0 + c] & 31;
			const int hd0 = H[k][0], hd1 = H[k][1], hd2 = H[k][2];  // column c - 1, row i - 1
			H2old[k] = hd2;
			int rowbest = 0;
#pragma unroll
			for (int f = 0; f < 3; ++f) {
				if (f > 0 && i >= g.ql[f]) break;  // :469,477
				const int sc = (int)s_score[((g.q[f][i] & 31) << 5) | tl];
				const int sm3 = f == 0 ? hd0 : f == 1 ? hd1 : hd2;
				const int sm4 = f == 0 ? upH2 : f == 1 ? hd0 : hd1;
				const int sm2 = f == 0 ? hd1 : f == 1 ? hd2 : dnH0;
				int cur = max(sm3 + sc, max(sm4, sm2) + sc - FS);
				cur = max(max(cur, upF[f]), max(dnE[f], 0));
				const int open = cur - go;
				F[k][f] = max(upF[f] - ge, open);
				E[k][f] = max(dnE[f] - ge, open);
				H[k][f] = cur;
				rowbest = max(rowbest, cur);
				if (TRACE) S[(size_t)(c + 1) * W + 3 * r + f] = cur;
			}
			if (rowbest > bestv[k]) { bestv[k] = rowbest; bestc[k] = c; }
		};
		if (g.ncol > 0) {
			const int nsteps = 2 * (g.ncol - 1) + g.B;
			const int nmacro = (nsteps + 1) >> 1;
			for (int m = 0; m < nmacro; ++m) {
				// ---- even step s = 2m: rows k = 0, 2, ..; column c = m - (r0 + k) / 2
				{
					int upF[3], upH2;
#pragma unroll
					for (int f = 0; f < 3; ++f) { upF[f] = __shfl_up_sync(FULL, F[R - 1][f], 1); if (lane == 0) upF[f] = 0; }
					upH2 = __shfl_up_sync(FULL, H2old[R - 1], 1);
					if (lane == 0) upH2 = 0;
#pragma unroll
					for (int k = 0; k < R; k += 2) {
						const int c = m - ((r0 + k) >> 1);
						if (k == 0) row_update(0, c, upF, upH2, E[1], H[1][0]);
						else row_update(k, c, F[k > 0 ? k - 1 : 0], H2old[k > 0 ? k - 1 : 0], E[k + 1], H[k + 1][0]);
					}
				}
				// ---- odd step s = 2m + 1: rows k = 1, 3, ..; column c = m - (r0 + k - 1) / 2
				{
					int dnE[3], dnH0;
#pragma unroll
					for (int f = 0; f < 3; ++f) { dnE[f] = __shfl_down_sync(FULL, E[0][f], 1); if (lane == 31) dnE[f] = 0; }
					dnH0 = __shfl_down_sync(FULL, H[0][0], 1);
					if (lane == 31) dnH0 = 0;
#pragma unroll
					for (int k = 1; k < R; k += 2) {
						const int c = m - ((r0 + k - 1) >> 1);
						if (k == R - 1) row_update(k, c, F[k - 1], H2old[k - 1], dnE, dnH0);
						else row_update(k, c, F[k - 1], H2old[k - 1], E[k + 1 < R ? k + 1 : 0], H[k + 1 < R ? k + 1 : 0][0]);
					}
				}
			}
		}
		// best score; first column that reaches it (the reference updates max_col only on a strictly larger column best, :495-498)
		int bv = 0, bc = 0;
#pragma unroll
		for (int k = 0; k < R; ++k)
			if (bestv[k] > bv || (bestv[k] == bv && bv > 0 && bestc[k] < bc)) { bv = bestv[k]; bc = bestc[k]; }
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const int ov = __shfl_xor_sync(FULL, bv, o), oc = __shfl_xor_sync(FULL, bc, o);
			if (ov > bv || (ov == bv && ov > 0 && oc < bc)) { bv = ov; bc = oc; }
		}
		if (lane == 0) { a.score[pi] = bv; a.max_col[pi] = bc; }
	}
}

struct FsWalkArgs {
	const int8_t *q_letters, *r_letters;
	const int64_t *q_limits, *r_limits;
	const dmnd_dp_problem* probs;
	uint32_t n, pos0;
	int frame_shift;
	const int32_t* score;
	const int32_t* max_col;
	const int32_t* matrix;
	const uint64_t* matrix_off;
	uint64_t matrix_base;
	dmnd_fs_result* res;
	uint8_t* transcripts;            // may be null
	const uint64_t* transcript_off;  // [problem]
	const uint32_t* transcript_cap;  // [problem]
};

// one problem per thread: traceback<_sv>() of banded_3frame_swipe.cpp:338-390
__device__ __forceinline__ void fs_walk_body(const FsWalkArgs& a, const DevParams* __restrict__ P) {
	const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= a.n) return;
	const uint32_t pi = a.pos0 + w;
	const dmnd_dp_problem pr = a.probs[pi];
	const FsGeom g = fs_geom(a.q_letters, a.q_limits, a.r_letters, a.r_limits, pr);
	dmnd_fs_result res;
	res.score = a.score[pi];
	res.q_begin = res.q_end = res.frame_begin = res.frame_end = res.t_begin = res.t_end = 0;
	res.identities = res.mismatches = res.gap_openings = res.length = res.gaps = res.positives = 0;
	res.transcript_off = 0; res.transcript_len = 0; res.status = 0;
	const int best = res.score;
	if (best > 0) {
		const int B3 = 3 * g.B, W = B3 + 1, FS = a.frame_shift;
		const int32_t* S = a.matrix + (a.matrix_off[pi] - a.matrix_base);
		const long long total = (long long)W * (long long)(g.ncol + 1);
		auto SV = [&](long long x) -> int { return (x >= 0 && x < total) ? S[x] : 0; };
		const int max_col = a.max_col[pi], col = max_col + 1, i0c = g.i0 + max_col;
		long long idx = -1;
		for (int x = max(-i0c, 0) * 3, xe = min(B3, g.dna_len - 2 - i0c * 3); x < xe; ++x)  // dp.traceback, :257-267
			if (S[(size_t)col * W + x] == best) { idx = (long long)col * W + x; break; }
		bool err = idx < 0;
		int fr = 0, i = 0, j = 0;
		uint32_t n = 0;
		uint8_t* out = a.transcripts ? a.transcripts + a.transcript_off[pi] : nullptr;
		const uint32_t cap = a.transcripts ? a.transcript_cap[pi] : 0;
		const int go = P->gap_open + P->gap_extend, ge = P->gap_extend;
		if (!err) {
			const int p0 = (int)(idx - (long long)col * W);
			fr = p0 % 3; i = i0c + p0 / 3; j = g.pos0 + max_col;
			res.q_end = i + 1; res.t_end = j + 1; res.frame_end = fr;
		}
		auto push = [&](unsigned b) { if (out && n < cap) out[n] = (uint8_t)b; ++n; };
		while (!err && SV(idx) > 0) {
			if (i < 0 || j < 0 || i >= g.ql[fr]) { err = true; break; }
			const int qa = g.q[fr][i] & 31, sa = g.t[j] & 31;
			const int m = P->score[(qa << 5) | sa], sc = SV(idx);
			int step = 0;  // 1 diagonal, 2 forward shift, 3 reverse shift
			if (sc == SV(idx - W) + m) step = 1;
			else if (sc == SV(idx - (W + 1)) + m - FS) step = 2;
			else if (sc == SV(idx - (W - 1)) + m - FS) step = 3;
			if (step) {
				if (qa == sa) { push(DMND_OP_MATCH << 6); ++res.identities; ++res.positives; }  // Hsp::push_match
				else { push((DMND_OP_SUBSTITUTION << 6) | sa); ++res.mismatches; if (m > 0) ++res.positives; }
				++res.length;
				if (step == 1) { idx -= W; --i; --j; }
				else if (step == 2) { push(DMND_TR_FRAMESHIFT_FWD); idx -= W + 1; --i; --j; if (--fr == -1) { fr = 2; --i; } }  // walk_forward_shift
				else { push(DMND_TR_FRAMESHIFT_REV); idx -= W - 1; --i; --j; if (++fr == 3) { fr = 0; ++i; } }               // walk_reverse_shift
				continue;
			}
			// walk_gap(d_begin, d_end), :204-244
			const int i0g = max(g.d_begin + j, 0), j0g = max(i - g.d_end, -1);
			const long long hstep = B3 - 2;
			long long h = idx - hstep, v = idx - 3;
			const long long h0 = idx - (long long)(j - j0g) * hstep, v0 = idx - (long long)(i - i0g + 1) * 3;
			int gp = go, l = 1, found = 0;
			while (v > v0 && h > h0) {
				if (sc + gp == SV(h)) { found = 2; break; }
				else if (sc + gp == SV(v)) { found = 1; break; }
				h -= hstep; v -= 3; ++l; gp += ge;
			}
			if (!found) while (v > v0) { if (sc + gp == SV(v)) { found = 1; break; } v -= 3; ++l; gp += ge; }
			if (!found) while (h > h0) { if (sc + gp == SV(h)) { found = 2; break; } h -= hstep; ++l; gp += ge; }
			if (!found) { err = true; break; }
			++res.gap_openings; res.length += l; res.gaps += l;  // Hsp::push_gap
			if (found == 1) { idx = v; i -= l; for (int k = 0; k < l; ++k) push(DMND_OP_INSERTION << 6); }
			else { idx = h; j -= l; for (int k = 0; k < l; ++k) push((DMND_OP_DELETION << 6) | (g.t[j + l - k] & 31)); }
		}
		if (err) res.status = 2;  // "Traceback error."
		else {
			res.q_begin = i + 1; res.t_begin = j + 1; res.frame_begin = fr;
			if (out) {
				if (n > cap) res.status = 1;
				else {
					for (uint32_t x = 0, y = n; x + 1 < y; ++x, --y) { const uint8_t tmp = out[x]; out[x] = out[y - 1]; out[y - 1] = tmp; }
					res.transcript_off = (uint32_t)a.transcript_off[pi];
					res.transcript_len = n;
				}
			}
		}
	}
	a.res[pi] = res;
}
__global__ void __launch_bounds__(128) fs_walk_kernel(const FsWalkArgs a, const DevParams* __restrict__ P) { fs_walk_body(a, P); }

}  // namespace dmnd_cuda
