// dpx_rates.cu -- issue-rate micro-benchmarks of the integer instructions the banded SWIPE kernel is built from (sm_100a).
// One result line per instruction (or mix): lane-instructions per clock per SM.  Built by `make ubench`, run on the GPU box:
//     tools/ubench/dpx_rates > gpurun_out/dpx_rates.txt
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
template<int OP>
__global__ void __launch_bounds__(256) rate_kernel(unsigned* out, int iters, unsigned b, unsigned c) {
	unsigned a[CHAINS];
#pragma unroll
	for (int k = 0; k < CHAINS; ++k) a[k] = threadIdx.x * 0x00010001u + k;
	for (int i = 0; i < iters; ++i) {
		b = b * 1664525u + 1013904223u; c ^= b >> 7;  // uniform-datapath work: keeps max / add chains from being folded
#pragma unroll
		for (int k = 0; k < CHAINS; ++k) {
			if (OP == 0) a[k] = (unsigned)__viaddmax_s32((int)a[k], (int)b, (int)c);
			else if (OP == 1) a[k] = __viaddmax_s16x2(a[k], b, c);
			else if (OP == 2) a[k] = __vimax3_s16x2_relu(a[k], b, c);
			else if (OP == 3) a[k] = __vimax_s16x2_relu(a[k], b);
			else if (OP == 4) a[k] = __vadd2(a[k], b);
			else if (OP == 5) a[k] = __byte_perm(a[k], b, c);
			else if (OP == 6) a[k] = (a[k] & b) ^ c;                       // LOP3
			else if (OP == 7) a[k] = a[k] * b + c;                           // IMAD
			else if (OP == 8) a[k] = a[k] + b + c;                           // IADD3
			else if (OP == 9) a[k] = __vimax3_u32(a[k], b, c);
			else if (OP == 10) { a[k] = __viaddmax_s16x2(a[k], b, c); a[k] = a[k] * b + c; }              // DPX + IMAD (2 instr)
			else if (OP == 11) { a[k] = __viaddmax_s16x2(a[k], b, c); a[k] = (a[k] & b) ^ c; }            // DPX + LOP3 (2 instr)
			else if (OP == 12) { a[k] = __vminu2(a[k], b); }
			else if (OP == 13) { a[k] = __viaddmax_s16x2(a[k], b, c); a[k] = a[k] * b + c; a[k] = a[k] * c + b; }  // DPX + 2 IMAD (3 instr)
			else if (OP == 14) { a[k] = __funnelshift_r(a[k], b, c); }       // SHF
		}
	}
	unsigned s = 0;
#pragma unroll
	for (int k = 0; k < CHAINS; ++k) s += a[k];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int OP> static void run(const char* name, int instr_per_iter, unsigned* d, int sms, double clk_hz) {
	const int blocks = sms * 8, threads = 256, iters = 1 << 14;
	rate_kernel<OP><<<blocks, threads>>>(d, 256, 3, 5);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	rate_kernel<OP><<<blocks, threads>>>(d, iters, 3, 5);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
	const double lane_instr = (double)blocks * threads * iters * CHAINS * instr_per_iter;
	printf("%-28s %8.3f ms  %7.2f T lane-instr/s  %6.1f lane-instr/clk/SM\n", name, ms, lane_instr / (ms * 1e-3) * 1e-12, lane_instr / (ms * 1e-3) / clk_hz / sms);
}

int main() {
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
	const double clk = clk_khz * 1e3;
	printf("%s, %d SMs, %.0f MHz (rates per clock assume this clock)\n", p.name, p.multiProcessorCount, clk * 1e-6);
	unsigned* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
	const int sms = p.multiProcessorCount;
	run<0>("VIADDMNMX.S32", 1, d, sms, clk);
	run<1>("VIADDMNMX.S16x2", 1, d, sms, clk);
	run<2>("VIMNMX3.S16x2.RELU", 1, d, sms, clk);
	run<3>("VIMNMX.S16x2.RELU", 1, d, sms, clk);
	run<4>("VIADD.16x2", 1, d, sms, clk);
	run<5>("PRMT", 1, d, sms, clk);
	run<6>("LOP3", 1, d, sms, clk);
	run<7>("IMAD", 1, d, sms, clk);
	run<8>("IADD3", 1, d, sms, clk);
	run<9>("VIMNMX3.U32", 1, d, sms, clk);
	run<12>("VIMNMX.U16x2", 1, d, sms, clk);
	run<14>("SHF", 1, d, sms, clk);
	run<10>("VIADDMNMX.S16x2 + IMAD", 2, d, sms, clk);
	run<11>("VIADDMNMX.S16x2 + LOP3", 2, d, sms, clk);
	run<13>("VIADDMNMX.S16x2 + 2 IMAD", 3, d, sms, clk);
	return cudaGetLastError() != cudaSuccess;
}
