// emu_cuda.h -- lock-step CPU emulation of the CUDA built-ins used by diamond_b200/csrc/cuda/mask_kernels.cuh, so that the
// kernel SOURCE can be compiled with g++ and run against the oracle without a GPU (tests/emu_mask.cpp).
//
// Every CUDA thread of a block is a ucontext coroutine on one OS thread; blocks run one after the other.  A warp-level
// primitive (__shfl_*_sync, __syncwarp) posts the lane's value, then yields until every lane named in its mask has posted
// for the same round -- lanes really exchange registers, divergent 8-lane groups of one warp progress independently, a
// wrong mask or a missing participant deadlocks (reported) instead of passing silently.  __syncthreads waits for all
// threads of the block that have not returned.  Atomics are plain read-modify-writes (one OS thread).  fp32 intrinsics map
// to the IEEE operators; compile with -ffp-contract=off and without -ffast-math.
#pragma once
#include <ucontext.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static

struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
static EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {

struct Thread {
	ucontext_t ctx;
	std::vector<char> stack;
	bool done = false;
	unsigned tid = 0;
	uint64_t warp_round = 0;   // number of warp-level primitives this thread has entered
	uint64_t posted[2] = { 0, 0 };  // value posted for round parity
	uint64_t block_round = 0;  // __syncthreads count
};
static std::vector<Thread> g_threads;
static ucontext_t g_sched;
static int g_cur = -1;
static std::function<void()> g_body;
static uint64_t g_switches = 0;

static void trampoline() {
	g_body();
	g_threads[(size_t)g_cur].done = true;
	swapcontext(&g_threads[(size_t)g_cur].ctx, &g_sched);
}
static void yield() { swapcontext(&g_threads[(size_t)g_cur].ctx, &g_sched); }

// Runs `body` once per thread of every block of the grid.
static void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
	g_body = body;
	gridDim.x = grid; blockDim.x = block;
	for (unsigned b = 0; b < grid; ++b) {
		blockIdx.x = b;
		g_threads.assign(block, Thread());
		for (unsigned t = 0; t < block; ++t) {
			Thread& th = g_threads[t];
			th.tid = t; th.stack.resize(256 << 10);
			getcontext(&th.ctx);
			th.ctx.uc_stack.ss_sp = th.stack.data(); th.ctx.uc_stack.ss_size = th.stack.size(); th.ctx.uc_link = nullptr;
			makecontext(&th.ctx, trampoline, 0);
		}
		unsigned alive = block;
		uint64_t idle_sweeps = 0;
		while (alive) {
			unsigned progressed = 0;
			for (unsigned t = 0; t < block; ++t) {
				if (g_threads[t].done) continue;
				g_cur = (int)t; threadIdx.x = t;
				const uint64_t before = g_threads[t].warp_round + g_threads[t].block_round;
				swapcontext(&g_sched, &g_threads[t].ctx);
				++g_switches;
				if (g_threads[t].done) { --alive; ++progressed; }
				else if (g_threads[t].warp_round + g_threads[t].block_round != before) ++progressed;
			}
			if (progressed) idle_sweeps = 0;
			else if (++idle_sweeps > 1000) { fprintf(stderr, "emu: deadlock in block %u (a warp primitive waits for a lane that never arrives)\n", b); exit(3); }
		}
	}
}

// Posts `v` for this thread's next warp round and waits until every lane in `mask` (of this thread's warp) has posted the
// same round; returns the round's parity slot to read partners from.
static int warp_rendezvous(unsigned mask, uint64_t v) {
	Thread& me = g_threads[(size_t)g_cur];
	const unsigned lane = me.tid & 31u, wbase = me.tid & ~31u;
	if (!((mask >> lane) & 1u)) { fprintf(stderr, "emu: thread %u calls a warp primitive with a mask (%08x) that excludes it\n", me.tid, mask); exit(3); }
	const uint64_t round = ++me.warp_round;
	const int slot = (int)(round & 1);
	me.posted[slot] = v;
	for (;;) {
		bool all = true;
		for (unsigned l = 0; l < 32 && all; ++l) {
			if (!((mask >> l) & 1u)) continue;
			const unsigned t = wbase + l;
			if (t >= g_threads.size()) { fprintf(stderr, "emu: mask names lane %u beyond the block\n", l); exit(3); }
			// partners of one group always execute the same sequence of primitives, so their round counters march together;
			// lanes of OTHER groups in the warp keep their own counters (they are not in the mask)
			if (g_threads[t].warp_round < round) all = false;  // (a lane that returned without reaching this round never arrives: reported as a deadlock)
		}
		if (all) break;
		yield();
	}
	return slot;
}
static uint64_t read_lane(unsigned src_lane, int slot) {
	const Thread& me = g_threads[(size_t)g_cur];
	return g_threads[(me.tid & ~31u) + src_lane].posted[slot];
}

}  // namespace emu

// NOTE on round counters: a lane's counter only has to agree with the lanes in its own masks.  Groups use disjoint masks and
// the full-warp kernels (popc_kernel) use 0xffffffff from the first primitive on, so counters never mix.
template<typename T> static inline T emu_bits_to(uint64_t u) { T v; static_assert(sizeof(T) == 4 || sizeof(T) == 8, "32- and 64-bit shuffles only"); memcpy(&v, &u, sizeof(T)); return v; }
template<typename T> static inline uint64_t emu_to_bits(T v) { uint64_t u = 0; static_assert(sizeof(T) == 4 || sizeof(T) == 8, "32- and 64-bit shuffles only"); memcpy(&u, &v, sizeof(T)); return u; }

template<typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
	const int slot = emu::warp_rendezvous(mask, emu_to_bits(v));
	const unsigned lane = threadIdx.x & 31u, src = lane ^ (unsigned)lane_mask;
	const T r = emu_bits_to<T>(emu::read_lane(src, slot));
	emu::warp_rendezvous(mask, 0);  // nobody overwrites a slot before every partner has read it
	return r;
}
template<typename T> static inline T __shfl_sync(unsigned mask, T v, int src_lane) {
	const int slot = emu::warp_rendezvous(mask, emu_to_bits(v));
	const T r = emu_bits_to<T>(emu::read_lane((unsigned)src_lane & 31u, slot));
	emu::warp_rendezvous(mask, 0);
	return r;
}
template<typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
	const int slot = emu::warp_rendezvous(mask, emu_to_bits(v));
	const unsigned lane = threadIdx.x & 31u, src = lane + delta;
	const T r = src < 32 ? emu_bits_to<T>(emu::read_lane(src, slot)) : v;
	emu::warp_rendezvous(mask, 0);
	return r;
}
static inline unsigned __ballot_sync(unsigned mask, bool pred) {
	const int slot = emu::warp_rendezvous(mask, pred ? 1u : 0u);
	unsigned r = 0;
	for (unsigned l = 0; l < 32; ++l) if (((mask >> l) & 1u) && emu::read_lane(l, slot)) r |= 1u << l;
	emu::warp_rendezvous(mask, 0);
	return r;
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_rendezvous(mask, 0); }
static inline void __syncthreads() {
	emu::Thread& me = emu::g_threads[(size_t)emu::g_cur];
	const uint64_t round = ++me.block_round;
	for (;;) {
		bool all = true;
		for (const emu::Thread& t : emu::g_threads) if (!t.done && t.block_round < round) { all = false; break; }
		if (all) return;
		emu::yield();
	}
}

static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{ x, y, z, w }; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
#define __log2f(x) log2f(x)
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (shift & 31u)); }
template<typename T, typename U> static inline T atomicAdd(T* p, U v) { const T o = *p; *p = (T)(o + (T)v); return o; }
template<typename T, typename U> static inline T atomicOr(T* p, U v) { const T o = *p; *p = (T)(o | (T)v); return o; }
template<typename T, typename U> static inline T atomicAnd(T* p, U v) { const T o = *p; *p = (T)(o & (T)v); return o; }

// ---- additions for the packed 16-bit banded SWIPE kernel (diamond_b200/csrc/cuda/swipe16.cuh) ----
#define __align__(n)
#define __restrict__
template<typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
	const int slot = emu::warp_rendezvous(mask, emu_to_bits(v));
	const unsigned lane = threadIdx.x & 31u, base = lane & ~(unsigned)(width - 1);
	const T r = lane >= base + delta ? emu_bits_to<T>(emu::read_lane(lane - delta, slot)) : v;
	emu::warp_rendezvous(mask, 0);
	return r;
}
template<typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width) {
	const int slot = emu::warp_rendezvous(mask, emu_to_bits(v));
	const unsigned lane = threadIdx.x & 31u, base = lane & ~(unsigned)(width - 1), src = lane + delta;
	const T r = src < base + (unsigned)width ? emu_bits_to<T>(emu::read_lane(src, slot)) : v;
	emu::warp_rendezvous(mask, 0);
	return r;
}
static inline bool __all_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) == mask; }
static inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0; }
template<typename T, typename U> static inline T atomicExch(T* p, U v) { const T o = *p; *p = (T)v; return o; }
static inline short emu_lo(unsigned x) { return (short)(x & 0xffffu); }
static inline short emu_hi(unsigned x) { return (short)(x >> 16); }
static inline unsigned emu_pack(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }  // wraps like the hardware
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) {
	const int lo = (short)(emu_lo(a) + emu_lo(b)), hi = (short)(emu_hi(a) + emu_hi(b));
	return emu_pack(lo > emu_lo(c) ? lo : emu_lo(c), hi > emu_hi(c) ? hi : emu_hi(c));
}
static inline unsigned __viaddmax_s16x2_relu(unsigned a, unsigned b, unsigned c) {
	const unsigned r = __viaddmax_s16x2(a, b, c);
	return emu_pack(emu_lo(r) > 0 ? emu_lo(r) : 0, emu_hi(r) > 0 ? emu_hi(r) : 0);
}
static inline unsigned __vmaxs2(unsigned a, unsigned b) { return emu_pack(emu_lo(a) > emu_lo(b) ? emu_lo(a) : emu_lo(b), emu_hi(a) > emu_hi(b) ? emu_hi(a) : emu_hi(b)); }
static inline unsigned __vimax_s16x2_relu(unsigned a, unsigned b) {
	const unsigned r = __vmaxs2(a, b);
	return emu_pack(emu_lo(r) > 0 ? emu_lo(r) : 0, emu_hi(r) > 0 ? emu_hi(r) : 0);
}
static inline unsigned __vminu2(unsigned a, unsigned b) {
	const unsigned al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
	return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
}
static inline unsigned __vimax3_u32(unsigned a, unsigned b, unsigned c) { const unsigned m = a > b ? a : b; return m > c ? m : c; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
	const uint64_t v = ((uint64_t)b << 32) | a;
	unsigned r = 0;
	for (int k = 0; k < 4; ++k) {
		const unsigned sel = (s >> (4 * k)) & 0xfu;
		unsigned byte = (unsigned)(v >> (8 * (sel & 7u))) & 0xffu;
		if (sel & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;
		r |= byte << (8 * k);
	}
	return r;
}
