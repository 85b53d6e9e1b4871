// Development aid: per-SM issue throughput of the integer instructions the filter uses (sm_100a).
// Each kernel runs a long chain-free loop of one instruction kind; result = warp-instructions / clk / SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define REPS 4096
template <int OP>
__global__ void __launch_bounds__(1024) K(unsigned *out, unsigned k1, unsigned k2, unsigned k3, long long *clk)
{
	unsigned a[8];
	for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 7 + i * k1;
	long long t0 = clock64();
#pragma unroll 1
	for (int r = 0; r < REPS; r++) {
#pragma unroll
		for (int i = 0; i < 8; i++) {
			unsigned x = a[i];
			if (OP == 0) x = x + k1 + k2;                                  // IADD3
			if (OP == 1) x = (x & k1) ^ k2;                                // LOP3
			if (OP == 2) x = (x >> 3) | (x << 29);                         // SHF
			if (OP == 3) x = __byte_perm(x, k1, 0x2640);                   // PRMT
			if (OP == 4) x = __vadd2(x, k1);                               // VIADD.16x2
			if (OP == 5) x = __vmaxs2(x, k1);                              // VIMNMX.S16x2
			if (OP == 6) x = __viaddmin_s16x2_relu(x, k1, k2);             // VIADDMNMX
			if (OP == 7) x = x * k1 + k2;                                  // IMAD
			if (OP == 8) x = __umulhi(x, k1) + k2;                         // IMAD.HI
			if (OP == 9) x = __dp2a_lo(x, k1, k2);                         // IDP.2A
			if (OP == 10) x = x * k3 + k2;                                 // IMAD by opaque 1 (add on the FMA pipe)
			if (OP == 11) x = (x & 1) ? x + k1 : x + k2;                   // ISETP+SEL-ish
			a[i] = x;
		}
	}
	long long t1 = clock64();
	unsigned s = 0;
	for (int i = 0; i < 8; i++) s += a[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

// mixed: A ALU ops (LOP3) and F FMA-pipe ops (IMAD) per iteration, independent chains
template <int A, int F>
__global__ void __launch_bounds__(1024) Mix(unsigned *out, unsigned k1, unsigned k2, long long *clk)
{
	unsigned a[8], f[8];
	for (int i = 0; i < 8; i++) { a[i] = threadIdx.x + i; f[i] = threadIdx.x * 3 + i; }
	long long t0 = clock64();
#pragma unroll 1
	for (int r = 0; r < REPS; r++) {
#pragma unroll
		for (int i = 0; i < A; i++) a[i] = (a[i] & k1) ^ k2;
#pragma unroll
		for (int i = 0; i < F; i++) f[i] = f[i] * k1 + k2;
	}
	long long t1 = clock64();
	unsigned s = 0;
	for (int i = 0; i < 8; i++) s += a[i] + f[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int OP> void Run(const char *name, unsigned *out, long long *clk, int warps)
{
	K<OP><<<1, warps * 32>>>(out, 0x00ff00ffu, 0x01010101u, 1u, clk);
	cudaDeviceSynchronize();
	K<OP><<<1, warps * 32>>>(out, 0x00ff00ffu, 0x01010101u, 1u, clk);
	cudaError_t e = cudaDeviceSynchronize();
	long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
	double inst = (double)REPS * 8 * warps;   // warp-instructions of the kind under test (plus loop overhead)
	printf("%-12s warps=%2d  %.3f warp-inst/clk/SM  (%s)\n", name, warps, inst / c, cudaGetErrorString(e));
}
template <int A, int F> void RunMix(unsigned *out, long long *clk, int warps)
{
	Mix<A, F><<<1, warps * 32>>>(out, 0x00ff00ffu, 0x01010101u, clk);
	cudaDeviceSynchronize();
	Mix<A, F><<<1, warps * 32>>>(out, 0x00ff00ffu, 0x01010101u, clk);
	cudaDeviceSynchronize();
	long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
	printf("mix LOP3 x%d + IMAD x%d warps=%2d  %.3f warp-inst/clk/SM\n", A, F, warps, (double)REPS * (A + F) * warps / c);
}

int main()
{
	unsigned *out; long long *clk;
	cudaMalloc(&out, 1024 * 4); cudaMalloc(&clk, 8);
	for (int warps : { 16, 32 }) {
		Run<0>("IADD3", out, clk, warps); Run<1>("LOP3", out, clk, warps); Run<2>("SHF", out, clk, warps);
		Run<3>("PRMT", out, clk, warps); Run<4>("VIADD.16x2", out, clk, warps); Run<5>("VIMNMX.16x2", out, clk, warps);
		Run<6>("VIADDMNMX", out, clk, warps); Run<7>("IMAD", out, clk, warps); Run<8>("IMAD.HI", out, clk, warps);
		Run<9>("IDP.2A", out, clk, warps); Run<10>("IMAD(x*1+y)", out, clk, warps); Run<11>("SETP/SEL", out, clk, warps);
	}
	RunMix<8, 0>(out, clk, 32); RunMix<0, 8>(out, clk, 32); RunMix<4, 4>(out, clk, 32); RunMix<6, 2>(out, clk, 32);
	RunMix<5, 3>(out, clk, 32); RunMix<8, 8>(out, clk, 32); RunMix<8, 4>(out, clk, 32);
	return 0;
}
