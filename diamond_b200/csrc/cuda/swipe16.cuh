// swipe16.cuh -- banded SWIPE in packed 16-bit lanes (DP::BandedSwipe::swipe<ScoreVector<int16_t>>, dp/swipe/banded_swipe.h:189-351,
// cell_update.h:103-141) for sm_100a.  Device code only; tests/emu_swipe16.cpp compiles THIS file for the CPU behind
// tests/emu_cuda.h and checks scores, end cells and transcripts against the oracle.
//
// Mapping.  One problem (DpTarget) per QUARTER warp: 8 lanes x R band rows (R = 4, 8, 12, 16 => bands of up to 32 / 64 / 96 /
// 128 diagonals; a warp holds four problems of similar size in lock step).  Lane t owns the R consecutive diagonals
// r = t*R .. t*R+R-1 and keeps them as R/2 REGISTER PAIRS: pair j = (row j | row j + R/2 << 16), both halves signed 16 bit.
// The wavefront is the one of swipe.cu (time s = 2c + r: at macro step m the rows with even k update column m - (r >> 1),
// then the rows with odd k), and R/2 is even, so both rows of a pair are always active in the same half step: every DPX
// instruction (VIADDMNMX.S16x2, VIMNMX3.S16x2.RELU) updates two cells.  Neighbour values are whole pairs: the horizontal gap
// of pair j comes from pair j + 1, the vertical gap from pair j - 1; only the two pairs at the ends of a lane need one PRMT
// with the shuffled pair of the adjacent lane.
//
// Scores.  Instead of a per-problem score profile (27 x query length bytes, which limited residency to ~20 problems per SM)
// the CTA holds ONE table in shared memory: the 27 x 32 matrix S[t][q] + 128 as bytes (row 26 and the query codes >= 26 are 0 =
// score -128: delimiter / outside the sequence), REPLICATED PER LANE -- entry e = t * 32 + q of lane l lives in byte e & 3 of
// word (e >> 2) * 32 + l, i.e. in bank l -- so the 32 lanes of a warp, which look up unrelated (t, q) pairs, never meet in a bank
// (the 55 KB table with the bias folded in that this replaces ran at 3.8 wavefronts per request and kept the shared-memory pipe
// 87 % busy: profiles/ncu_swipe16_r2b.txt).  A problem keeps its query as byte offsets into a table row ("codes", 16 bit) and its
// composition bias as packed pairs (bias[i] - 128 | bias[i + R/4] - 128 << 16: the two rows of a register pair are R/4 query
// positions apart; stored as two bytes, expanded when a lane loads one), so one LDS.U8 per cell + one IMAD packs the two scores of a pair and one VIADDMNMX.S16x2 adds both biases and
// removes both offsets.  A lane keeps the R/2 + 1 codes, R/4 + 1 bias pairs and R/2 target rows it needs in registers and shifts
// them by one per macro step.
//
// Results are the reference's: H, hgap, vgap floored at 0 (saturating lanes, score_vector_int16.h), trace masks per cell
// (banded_matrix.h:313-445), end cell = first column of the maximum, then the last row (banded_swipe.h:312-328).  A score
// that reaches 32767 - 255 raises the overflow flag and the caller repeats the call on the exact int32 kernels of swipe.cu --
// the reference's int16 -> int32 cascade (banded_swipe.h:337).
#pragma once
#include "dev_params.h"

// dynamic shared memory of the kernel (tests/emu_cuda.h defines DMND_DYN_SMEM as a static arena for the CPU emulation)
#ifndef DMND_DYN_SMEM
#define DMND_DYN_SMEM(name) extern __shared__ __align__(16) int8_t name[]
#endif

namespace dmnd_cuda {

struct ProbGeom {  // derived on the device from the problem + block limits
	const int8_t *q, *cb, *t;
	int qlen, tlen, d_begin, B, j0, cols;
};

struct SwipeArgs {
	const int8_t *q_letters, *q_bias, *r_letters;
	const int64_t *q_limits, *r_limits;
	const dmnd_dp_problem* probs;
	const uint32_t* order;  // problem indices of this launch, heaviest first
	uint32_t n;
	int32_t* score;         // [problem]
	int32_t* end_cell;      // [problem][2] = (column, band row) of the end cell, traceback only
	uint8_t* trace;         // traceback masks, see trace_store() / s16_trace_nibble()
	const uint64_t* trace_excl; // exclusive prefix of trace bytes over the ORDER sequence (global order position)
	uint64_t trace_base;    // prefix value at the first position of the slice in flight
	uint32_t order_pos0;    // global order position of order[0]
	unsigned int* work;     // atomic work counter
};

__device__ __forceinline__ ProbGeom geom(const SwipeArgs& a, const dmnd_dp_problem& pr) {
	ProbGeom g;
	const int64_t qo = a.q_limits[pr.query], to = a.r_limits[pr.target];
	g.qlen = (int)(a.q_limits[pr.query + 1] - qo - 1);
	g.tlen = (int)(a.r_limits[pr.target + 1] - to - 1);
	g.q = a.q_letters + qo; g.cb = a.q_bias + qo; g.t = a.r_letters + to;
	g.d_begin = pr.d_begin;
	g.B = pr.d_end - pr.d_begin;
	const int i1 = max(pr.d_end - 1, 0);
	g.j0 = i1 - (pr.d_end - 1);
	g.cols = min(g.qlen - 1 - pr.d_begin, g.tlen - 1) + 1 - g.j0;  // dp/dp.h:47-52
	return g;
}

// ---- the shared score table ------------------------------------------------------------------------------------------
constexpr int S16_TABLE_ENTRIES = 27 * 32;             // global image: S[t][q] + 128, one byte per entry
constexpr int S16_TROW = 8 * 32 * 4;                   // bytes of one target-letter row of the lane-replicated table in shared memory
constexpr int S16_TABLE_BYTES = 27 * S16_TROW;
constexpr int S16_HALO_Q = 31;                         // query code of a position outside the query (entry 0 = score -128)
constexpr unsigned S16_HALO_BB = 0xFF80FF80u;          // bias pair of such positions: (0 - 128, 0 - 128)
__host__ __device__ __forceinline__ int s16_qoff(int q) { return (q >> 2) * 128 + (q & 3); }  // byte offset of query code q inside a table row (+ 4 * lane)
constexpr int S16_LANES = 8;                           // lanes per problem
constexpr int S16_MAX_BAND = 128;
constexpr int S16_TILE = 8;                            // macro steps per trace tile
constexpr int S16_MAX_QLEN = 10000;                     // longest query of the packed kernel: one warp's four problems (4 bytes per query position each) beside the table in shared memory
constexpr int S16_MAX_MACRO = 32000;                   // macro steps representable in the 16-bit column key
__host__ __device__ __forceinline__ int s16_rows(int B) { return B <= 32 ? 4 : B <= 64 ? 8 : B <= 96 ? 12 : 16; }
// trace bytes of one macro step of one problem: 8 lanes x R/2 bytes, laid out as R/8 regions of 8 x 4 bytes (the full words
// of the lanes) followed by one region of 8 x 2 bytes when R % 8 == 4
__host__ __device__ __forceinline__ int s16_step_bytes(int R) { return 4 * R; }
// trace bytes of a problem: whole tiles
__host__ __device__ __forceinline__ unsigned long long s16_trace_bytes(int R, unsigned long long nmacro) { return (nmacro + S16_TILE - 1) / S16_TILE * (unsigned long long)(S16_TILE * 4 * R); }

// Fills the global image of the table from the 32 x 32 score matrix (stats/score_matrix.h:35-44 layout): entry t * 32 + q =
// S[t][q] + 128 for letters inside the alphabet, 0 (= score -128) elsewhere.  *bad stays 0 (kept for the caller's check).
__global__ void s16_table_kernel(const DevParams* __restrict__ P, uint8_t* table, unsigned* bad) {
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= S16_TABLE_ENTRIES) return;
	const int t = idx >> 5, ql = idx & 31;
	int v = 0;
	if (t < 26 && ql < 26) v = (int)P->score[(t << 5) | ql] + 128;
	if (v < 0 || v > 255) { atomicExch(bad, 1u); v = 0; }
	table[idx] = (uint8_t)v;
}

struct S16Args {
	const uint8_t* table;    // S16_TABLE_ENTRIES bytes, global
	int qstride;             // uint16 elements per problem slot in shared memory, >= max(qlen) + 8 R + 4 of the launch
	unsigned int* overflow;  // raised when the call has to be repeated on the int32 kernels
};

// nibble (4 trace bits, stored inverted) of cell (column c, band row r): tiles of S16_TILE macro steps; inside a tile every lane's
// S16_TILE steps of one word are adjacent (R / 8 regions of 8 lanes x 32 bytes, then one of 8 x 16 bytes when R % 8 == 4)
__device__ __forceinline__ unsigned s16_trace_nibble(const uint8_t* tr, int R, int c, int r) {
	const int m = c + (r >> 1), lane = r / R, k = r - lane * R, half = R >> 1;
	const int hi = k >= half ? 1 : 0, j = k - hi * half, w = j >> 2, p = j & 3;
	const uint8_t* tb = tr + (size_t)(m / S16_TILE) * (size_t)(S16_TILE * 4 * R);
	const int t = m & (S16_TILE - 1), full = R >> 3;
	unsigned v;
	if (w < full) v = tb[w * (S16_LANES * 4 * S16_TILE) + lane * (4 * S16_TILE) + t * 4 + hi * 2 + (p >> 1)];
	else v = tb[full * (S16_LANES * 4 * S16_TILE) + lane * (2 * S16_TILE) + t * 2 + hi];
	return (~(v >> ((p & 1) * 4))) & 15u;
}

// one table entry (a byte), zero-extended, from a shared-memory byte address (tests/emu_cuda.h maps the address back to its arena)
#ifndef DMND_S16_LDS
__device__ __forceinline__ unsigned s16_lds(unsigned addr) {
	unsigned v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
#endif

template<int R, bool TRACE>
__global__ void __launch_bounds__(128) swipe16_kernel(const SwipeArgs a, const DevParams* __restrict__ P, const S16Args sa) {
	DMND_DYN_SMEM(smem16);
	constexpr int NP = R / 2, U = R / 2, NB = R / 4, HALO = S16_LANES * R / 2 + S16_TILE, NW = (NP + 3) / 4, FULL = R / 8;  // (the halo covers the steps before m_lo of the first trace tile)
	static_assert(R % 4 == 0 && R >= 4 && R <= 16, "rows per lane");
	{  // table: global -> shared, once per CTA; word (e >> 2) * 32 + l holds the entries 4 (e >> 2) .. + 3 for lane l
		const uint32_t* src = reinterpret_cast<const uint32_t*>(sa.table);
		uint32_t* dst = reinterpret_cast<uint32_t*>(smem16);
		for (int i = threadIdx.x; i < S16_TABLE_BYTES / 4; i += blockDim.x) dst[i] = src[i >> 5];
	}
	__syncthreads();
	const unsigned FULLM = 0xffffffffu;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane >> 3, gl = lane & 7;
	const unsigned tab = (unsigned)__cvta_generic_to_shared(smem16);  // byte address of the table in shared memory
	const unsigned one = P->one, k65536 = P->k65536;
	// per problem slot: qstride bias pairs (two bytes: bias[i] + 128, bias[i + R/4] + 128), then qstride codes (16 bit)
	uint16_t* bbs = reinterpret_cast<uint16_t*>(smem16 + ((S16_TABLE_BYTES + 15) & ~15) + (size_t)(warp * 4 + sub) * (size_t)sa.qstride * 4);
	uint16_t* qc = bbs + sa.qstride;
	// a stored pair (u0 | u1 << 8) becomes the packed operand (u0 - 256 | (u1 - 256) << 16) = (bias - 128, bias' - 128): one PRMT spreads the bytes,
	// one OR sets the two high bytes
	auto bb_expand = [](unsigned v) { return __byte_perm(v, 0u, 0x4140u) | 0xFF00FF00u; };
	const unsigned neg2 = P->neg2;  // (-32768, -32768): the no-op third operand of the packed bias add (a run-time value keeps it a register operand)
	const unsigned go2 = (unsigned)(P->gap_open + P->gap_extend) * 0x00010001u, ge = (unsigned)P->gap_extend;
	const unsigned nge2 = (0x10000u - ge) * 0x00010001u & 0xffffffffu;  // (-ge, -ge)
	const unsigned selF = gl == 0 ? 0x54DDu : 0x5432u, selE = gl == S16_LANES - 1 ? 0xBB32u : 0x5432u;
	const int lofs = gl * U;
	const unsigned zero2 = P->zero;

	for (;;) {
		unsigned int w = 0;
		if (lane == 0) w = atomicAdd(a.work, 1u);
		w = __shfl_sync(FULLM, w, 0);
		if (w * 4u >= a.n) break;
		const uint32_t slot = w * 4u + (uint32_t)sub;
		const bool have = slot < a.n;
		const uint32_t pi = a.order[have ? slot : a.n - 1];
		const dmnd_dp_problem pr = a.probs[pi];
		ProbGeom g = geom(a, pr);
		if (!have) { g.B = 0; g.cols = 0; }
		const bool live = g.B > 0 && g.cols > 0;
		// ---- the query as table columns and bias pairs
		__syncwarp();
		if (live) {
			const int W = g.qlen + 8 * R + 4 + S16_TILE;
			for (int idx = gl; idx < W; idx += S16_LANES) {
				const int i = idx - HALO, i2 = i + NB;
				int code = s16_qoff(S16_HALO_Q), b0 = 0, b1 = 0;
				if (i >= 0 && i < g.qlen) { code = s16_qoff(g.q[i] & 31); b0 = (int)g.cb[i]; }
				if (i2 >= 0 && i2 < g.qlen) b1 = (int)g.cb[i2];
				qc[idx] = (uint16_t)code;  // byte offset inside a table row
				bbs[idx] = (uint16_t)((unsigned)(b0 + 128) | ((unsigned)(b1 + 128) << 8));
			}
		}
		__syncwarp();
		// ---- geometry of the wavefront
		const int ibase = g.j0 + g.d_begin;
		const int nsteps = 2 * (g.cols - 1) + g.B, nmacro = live ? (nsteps + 1) >> 1 : 0;
		const int m_lo = live ? max(0, -ibase - (HALO - S16_TILE)) : 0, m_hi = live ? min(nmacro, g.qlen - ibase) : 0;
		const int m0 = m_lo & ~(S16_TILE - 1);  // the wavefront starts at a trace-tile boundary; the steps before m_lo only see the halo (all state stays 0)
		int trip = max((m_hi - m0 + S16_TILE - 1) / S16_TILE, 0);  // tiles
#pragma unroll
		for (int o = 8; o < 32; o <<= 1) trip = max(trip, __shfl_xor_sync(FULLM, trip, o));
		// per pair: gap-open constant (32767 for a row below the band: its H never opens a gap, so E of the first dead row stays
		// 0 = the reference's hgap_[band] sentinel) and the masks that keep dead rows out of the end-cell search
		unsigned GO2[NP], MUL[NP], MSK[NP];
#pragma unroll
		for (int j = 0; j < NP; ++j) {
			const bool lo_live = gl * R + j < g.B, hi_live = gl * R + j + NP < g.B;
			GO2[j] = (0x10000u - (lo_live ? (go2 & 0xffffu) : 32767u)) & 0xffffu;
			GO2[j] |= ((0x10000u - (hi_live ? (go2 & 0xffffu) : 32767u)) & 0xffffu) << 16;  // (-go, -go)
			MUL[j] = lo_live ? 65536u : 0u;
			MSK[j] = hi_live ? 0xffff0000u : 0u;
		}
		unsigned H[NP], E[NP], F[NP], best[NP];
#pragma unroll
		for (int j = 0; j < NP; ++j) { H[j] = 0; E[j] = 0; F[j] = 0; best[j] = 0; }
		auto trow_of = [&](int j) { const int jj = min(max(j, -1), g.tlen); return (int)tab + 4 * lane + min((int)(g.t[jj] & 31), 26) * S16_TROW; };  // byte address of the letter's row in this lane's bank
		int trow[U];
		unsigned qreg[U + 1], bbreg[NB + 1];
		int I0 = ibase + m0 + lofs + HALO;  // code index of row u = 0 (even k); odd k reads one further
		if (live) {
#pragma unroll
			for (int u = 0; u < U; ++u) trow[u] = trow_of(g.j0 + m0 - lofs - u);
#pragma unroll
			for (int v = 0; v <= U; ++v) qreg[v] = qc[I0 + v];
#pragma unroll
			for (int v = 0; v <= NB; ++v) bbreg[v] = bb_expand(bbs[I0 + v]);
		}
		else {
#pragma unroll
			for (int u = 0; u < U; ++u) trow[u] = (int)tab + 4 * lane + 26 * S16_TROW;
#pragma unroll
			for (int v = 0; v <= U; ++v) qreg[v] = (unsigned)s16_qoff(S16_HALO_Q);
#pragma unroll
			for (int v = 0; v <= NB; ++v) bbreg[v] = S16_HALO_BB;
			I0 = 0;
		}
		uint8_t* tr = (TRACE && live) ? a.trace + (a.trace_excl[a.order_pos0 + slot] - a.trace_base) : nullptr;
		// column keys of the macro step: key = score << 16 | ck, ck larger for the earlier column, hi row beats lo row on a tie
		unsigned ckLo = 2u * (unsigned)(S16_MAX_MACRO + 64 - m0), ckHi = ckLo + (unsigned)(R / 2 + 1);
		int m = m0;
		for (int it = 0; it < trip; ++it) {
		const int mt = m;  // first step of the tile
		unsigned buf[S16_TILE][NW];
#pragma unroll
		for (int ts = 0; ts < S16_TILE; ++ts) {
			const bool active = m < m_hi;
			int tnext = (int)tab + 4 * lane + 26 * S16_TROW;
			unsigned qnext = (unsigned)s16_qoff(S16_HALO_Q), bbnext = S16_HALO_BB;
			if (active) { tnext = trow_of(g.j0 + m + 1 - lofs); qnext = qc[I0 + U + 1]; bbnext = bb_expand(bbs[I0 + NB + 1]); }
			unsigned pk[NW];
#pragma unroll
			for (int x = 0; x < NW; ++x) pk[x] = 0;
			// ---- even rows of the lane: pairs j = 0, 2, ..
			{
				const unsigned f_sh = __shfl_up_sync(FULLM, F[NP - 1], 1, S16_LANES);
				const unsigned f_edge = __byte_perm(f_sh, F[NP - 1], selF);
#pragma unroll
				for (int j = 0; j < NP; j += 2) {
					const int ul = j >> 1, uh = ul + R / 4;
					const unsigned sc = __viaddmax_s16x2(s16_lds((unsigned)trow[uh] * one + qreg[uh]) * k65536 + s16_lds((unsigned)trow[ul] * one + qreg[ul]), bbreg[ul], neg2);
					const unsigned e_in = E[j + 1], f_in = j > 0 ? F[j > 0 ? j - 1 : 0] : f_edge;
					const unsigned h = __vimax_s16x2_relu(__viaddmax_s16x2(H[j], sc, e_in), f_in);
					const unsigned open = __viaddmax_s16x2_relu(h, GO2[j], zero2);
					const unsigned e_new = __viaddmax_s16x2_relu(e_in, nge2, open), f_new = __viaddmax_s16x2_relu(f_in, nge2, open);
					if (TRACE) {
						// inverted masks: min(difference, 1) per half (h >= e_in, f_in and e_new, f_new >= open always hold)
						const unsigned n0 = __vminu2(h - f_in, 0x00010001u), n1 = __vminu2(h - e_in, 0x00010001u);
						const unsigned n2 = __vminu2(f_new - open, 0x00010001u), n3 = __vminu2(e_new - open, 0x00010001u);
						pk[j >> 2] += ((n0 + 2u * n1) + 4u * (n2 + 2u * n3)) << (4 * (j & 3));
						best[j] = __vimax3_u32(best[j], h * MUL[j] + ckLo, (h & MSK[j]) | ckHi);
					}
					else best[j] = __vmaxs2(best[j], h);
					H[j] = h; E[j] = e_new; F[j] = f_new;
				}
			}
			// ---- odd rows: pairs j = 1, 3, ..
			{
				const unsigned e_sh = __shfl_down_sync(FULLM, E[0], 1, S16_LANES);
				const unsigned e_edge = __byte_perm(E[0], e_sh, selE);
#pragma unroll
				for (int j = 1; j < NP; j += 2) {
					const int ul = j >> 1, uh = ul + R / 4;
					const unsigned sc = __viaddmax_s16x2(s16_lds((unsigned)trow[uh] * one + qreg[uh + 1]) * k65536 + s16_lds((unsigned)trow[ul] * one + qreg[ul + 1]), bbreg[ul + 1], neg2);
					const unsigned e_in = j + 1 < NP ? E[j + 1 < NP ? j + 1 : 0] : e_edge, f_in = F[j - 1];
					const unsigned h = __vimax_s16x2_relu(__viaddmax_s16x2(H[j], sc, e_in), f_in);
					const unsigned open = __viaddmax_s16x2_relu(h, GO2[j], zero2);
					const unsigned e_new = __viaddmax_s16x2_relu(e_in, nge2, open), f_new = __viaddmax_s16x2_relu(f_in, nge2, open);
					if (TRACE) {
						const unsigned n0 = __vminu2(h - f_in, 0x00010001u), n1 = __vminu2(h - e_in, 0x00010001u);
						const unsigned n2 = __vminu2(f_new - open, 0x00010001u), n3 = __vminu2(e_new - open, 0x00010001u);
						pk[j >> 2] += ((n0 + 2u * n1) + 4u * (n2 + 2u * n3)) << (4 * (j & 3));
						best[j] = __vimax3_u32(best[j], h * MUL[j] + ckLo, (h & MSK[j]) | ckHi);
					}
					else best[j] = __vmaxs2(best[j], h);
					H[j] = h; E[j] = e_new; F[j] = f_new;
				}
			}
			if (TRACE) {
#pragma unroll
				for (int x = 0; x < NW; ++x) buf[ts][x] = pk[x];
			}
#pragma unroll
			for (int u = U - 1; u > 0; --u) trow[u] = trow[u - 1];
			trow[0] = tnext;
#pragma unroll
			for (int v = 0; v < U; ++v) qreg[v] = qreg[v + 1];
			qreg[U] = qnext;
#pragma unroll
			for (int v = 0; v < NB; ++v) bbreg[v] = bbreg[v + 1];
			bbreg[NB] = bbnext;
			if (active) { ++I0; ++m; }
			ckLo -= 2u; ckHi -= 2u;
		}
		// ---- the tile's trace: every lane writes its 8 steps as one 32-byte (full words) / 16-byte (half pieces) run, so that the
		// traceback walk, which mostly moves along a diagonal = one lane word per macro step, finds 8 consecutive steps in one sector
		if (TRACE && mt < m_hi) {
			uint8_t* tb = tr + (size_t)(mt / S16_TILE) * (size_t)(S16_TILE * 4 * R);
#pragma unroll
			for (int x = 0; x < FULL; ++x) {
				uint4* dst = reinterpret_cast<uint4*>(tb + x * (S16_LANES * 4 * S16_TILE) + gl * (4 * S16_TILE));
				dst[0] = make_uint4(buf[0][x], buf[1][x], buf[2][x], buf[3][x]);
				dst[1] = make_uint4(buf[4][x], buf[5][x], buf[6][x], buf[7][x]);
			}
			if (R % 8 == 4) {  // the last word holds two pairs: bytes 0 and 2 -> one 16-bit piece per step
				uint4* dst = reinterpret_cast<uint4*>(tb + FULL * (S16_LANES * 4 * S16_TILE) + gl * (2 * S16_TILE));
				dst[0] = make_uint4(__byte_perm(buf[0][NW - 1], buf[1][NW - 1], 0x6420), __byte_perm(buf[2][NW - 1], buf[3][NW - 1], 0x6420),
				                    __byte_perm(buf[4][NW - 1], buf[5][NW - 1], 0x6420), __byte_perm(buf[6][NW - 1], buf[7][NW - 1], 0x6420));
			}
		}
		}
		// ---- end cell of the problem
		if (TRACE) {
			int bv = 0, bc = 0, br = 0;
			bool wide = false;
#pragma unroll
			for (int j = 0; j < NP; ++j) {
				const int v = (int)(best[j] >> 16);
				if (v > 0) {
					const unsigned ck = best[j] & 0xffffu;
					// lo row: ck = 2 (MAX + 64 - m); hi row: ck = 2 (MAX + 64 - m) + R/2 + 1; R/2 + 1 is odd
					const int hi = (int)(ck & 1u);
					const int mm = S16_MAX_MACRO + 64 - (int)((ck - (hi ? (unsigned)(R / 2 + 1) : 0u)) >> 1);
					const int k = j + hi * NP, r = gl * R + k, c = mm - lofs - (k >> 1);
					wide |= v >= 32767 - 255;
					if (v > bv || (v == bv && (c < bc || (c == bc && r > br)))) { bv = v; bc = c; br = r; }
				}
			}
			if (__any_sync(FULLM, wide) && lane == 0) atomicExch(sa.overflow, 1u);
#pragma unroll
			for (int o = 4; o > 0; o >>= 1) {
				const int ov = __shfl_xor_sync(FULLM, bv, o), oc = __shfl_xor_sync(FULLM, bc, o), orr = __shfl_xor_sync(FULLM, br, o);
				if (ov > bv || (ov == bv && ov > 0 && (oc < bc || (oc == bc && orr > br)))) { bv = ov; bc = oc; br = orr; }
			}
			if (gl == 0 && have) { a.score[pi] = bv; a.end_cell[2 * (size_t)pi] = bc; a.end_cell[2 * (size_t)pi + 1] = br; }
		}
		else {
			int bv = 0;
#pragma unroll
			for (int j = 0; j < NP; ++j) {
				if (gl * R + j < g.B) bv = max(bv, (int)(short)(best[j] & 0xffffu));
				if (gl * R + j + NP < g.B) bv = max(bv, (int)(short)(best[j] >> 16));
			}
			if (__any_sync(FULLM, bv >= 32767 - 255) && lane == 0) atomicExch(sa.overflow, 1u);
#pragma unroll
			for (int o = 4; o > 0; o >>= 1) bv = max(bv, __shfl_xor_sync(FULLM, bv, o));
			if (gl == 0 && have) a.score[pi] = bv;
		}
	}
}

// ---- trace layout of the int32 kernels (swipe.cu) and the traceback walk shared by both kernel families ---------------
// register tile per lane of the warp kernels (32 R >= B); bands beyond 1024 diagonals use the same trace layout with R = 64 / 128
// and are evaluated by swipe_wide_kernel (one CTA per problem)
#define DMND_MAX_BAND 4096
__host__ __device__ __forceinline__ int tile_rows(int B) { return B <= 64 ? 2 : B <= 128 ? 4 : B <= 256 ? 8 : B <= 512 ? 16 : B <= 1024 ? 32 : B <= 2048 ? 64 : 128; }
// nibble of cell (column c, band row r) in the wavefront-major layout
__device__ __forceinline__ unsigned trace_nibble(const uint8_t* tr, int R, int c, int r) {
	const int m = c + (r >> 1), lane = r / R, k = r - lane * R;
	return (tr[((size_t)m * 32 + lane) * (R >> 1) + (k >> 1)] >> ((k & 1) * 4)) & 15u;
}

struct WalkArgs {
	const int8_t *q_letters, *q_bias, *r_letters;
	const int64_t *q_limits, *r_limits;
	const dmnd_dp_problem* probs;
	const uint32_t* order;
	uint32_t n;
	const int32_t* score;
	const int32_t* end_cell;
	const uint8_t* trace;
	const uint64_t* trace_excl;
	uint64_t trace_base;
	uint32_t order_pos0;
	dmnd_dp_result* res;
	int s16;                        // trace layout of this launch: swipe16_kernel (1) or the int32 warp / wide kernels (0)
	uint8_t* transcripts;           // may be null
	const uint64_t* transcript_off; // [problem], capacity qlen + tlen each
};

// S16R > 0: the launch's problems were evaluated by swipe16_kernel<S16R, true> (the register tile is a compile-time constant: the
// nibble address needs no division); S16R = 0: layout chosen per problem at run time (int32 kernels, and the CPU emulation's entry)
template<int S16R>
__device__ __forceinline__ void walk_body(const WalkArgs& a, const DevParams* __restrict__ P) {
	const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= a.n) return;
	const uint32_t pi = a.order[w];
	const dmnd_dp_problem pr = a.probs[pi];
	SwipeArgs sa;
	sa.q_letters = a.q_letters; sa.q_bias = a.q_bias; sa.r_letters = a.r_letters; sa.q_limits = a.q_limits; sa.r_limits = a.r_limits;
	const ProbGeom g = geom(sa, pr);
	dmnd_dp_result res;
	res.score = a.score[pi];
	res.q_begin = res.q_end = res.t_begin = res.t_end = 0;
	res.identities = res.mismatches = res.gap_openings = res.length = res.gaps = res.positives = 0;
	res.transcript_off = 0; res.transcript_len = 0; res.status = 0;
	const int best = res.score;
	if (best > 0) {
		const uint8_t* tr = a.trace + (a.trace_excl[a.order_pos0 + w] - a.trace_base);
		const int R = S16R ? S16R : (a.s16 ? s16_rows(g.B) : tile_rows(g.B));
		int c = a.end_cell[2 * (size_t)pi], r = a.end_cell[2 * (size_t)pi + 1];
		int i = g.j0 + g.d_begin + c + r, j = g.j0 + c;
		res.q_end = i + 1; res.t_end = j + 1;
		uint8_t* out = a.transcripts ? a.transcripts + a.transcript_off[pi] : nullptr;
		const uint32_t cap = (uint32_t)(g.qlen + g.tlen);
		uint32_t n = 0;
		int sc = 0;
		const int gopen = P->gap_open, gext = P->gap_extend;
		bool bad = false;
		while (i >= 0 && j >= 0 && sc < best) {
			if (c < 0 || r < 0 || r >= g.B) { bad = true; break; }
			const unsigned nib = ((S16R || a.s16) ? s16_trace_nibble(tr, R, c, r) : trace_nibble(tr, R, c, r));
			if ((nib & 3) == 0) {
				const int ql = g.q[i] & 31, sl = g.t[j] & 31;
				const int m = P->score[(ql << 5) | sl];
				sc += m + (int)g.cb[i];
				if (ql == sl) { ++res.identities; ++res.positives; if (out && n < cap) out[n] = (uint8_t)(DMND_OP_MATCH << 6); }
				else { ++res.mismatches; if (m > 0) ++res.positives; if (out && n < cap) out[n] = (uint8_t)((DMND_OP_SUBSTITUTION << 6) | sl); }
				++n; ++res.length;
				--i; --j; --c;
			}
			else if (nib & 1) {
				int l = 0;
				do { ++l; --i; --r; } while (r >= 0 && (((S16R || a.s16) ? s16_trace_nibble(tr, R, c, r) : trace_nibble(tr, R, c, r)) & 4) == 0 && i > 0);
				if (r < 0) { bad = true; break; }
				++res.gap_openings; res.length += l; res.gaps += l;
				for (int k = 0; k < l; ++k) { if (out && n < cap) out[n] = (uint8_t)(DMND_OP_INSERTION << 6); ++n; }
				sc -= gopen + l * gext;
			}
			else {
				int l = 0;
				do { ++l; --j; --c; ++r; } while (c >= 0 && r < g.B && (((S16R || a.s16) ? s16_trace_nibble(tr, R, c, r) : trace_nibble(tr, R, c, r)) & 8) == 0 && j > 0);
				if (c < 0 || r >= g.B) { bad = true; break; }
				++res.gap_openings; res.length += l; res.gaps += l;
				for (int k = 0; k < l; ++k) { if (out && n < cap) out[n] = (uint8_t)((DMND_OP_DELETION << 6) | (g.t[j + l - k] & 31)); ++n; }
				sc -= gopen + l * gext;
			}
		}
		if (bad || sc != best) res.status = 2;  // "Traceback error." (banded_swipe.h:176-177)
		res.q_begin = i + 1; res.t_begin = j + 1;
		if (out) {
			if (n > cap) res.status = 1;
			else {
				for (uint32_t x = 0, y = n; x + 1 < y; ++x, --y) { const uint8_t tmp = out[x]; out[x] = out[y - 1]; out[y - 1] = tmp; }
				res.transcript_off = (uint32_t)a.transcript_off[pi];
				res.transcript_len = n;
			}
		}
	}
	a.res[pi] = res;
}
__global__ void __launch_bounds__(128) walk_kernel(const WalkArgs a, const DevParams* __restrict__ P) { walk_body<0>(a, P); }
template<int S16R> __global__ void __launch_bounds__(128) walk16_kernel(const WalkArgs a, const DevParams* __restrict__ P) { walk_body<S16R>(a, P); }


}  // namespace dmnd_cuda
