// Microbenchmark: fp32 issue throughput on sm_100a -- scalar FFMA vs packed FFMA2 (fma.rn.f32x2), MUFU.EX2, FMNMX,
// and an FFMA+ALU mix.  Decides whether the compositing kernels should process two pixels per thread with packed math.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s)
{
	float a[8];
	u64 p[8];
#pragma unroll
	for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; p[i] = pk(a[i], a[i] + 0.5f); }
	const float m = s, c = s * 0.5f;
	const u64 m2 = pk(m, m), c2 = pk(c, c);
	int acc = threadIdx.x;
	for (int it = 0; it < iters; it++) {
		if (MODE == 0) {
#pragma unroll
			for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], m, c);
		} else if (MODE == 1) {
#pragma unroll
			for (int i = 0; i < 8; i++) p[i] = fma2(p[i], m2, c2);
		} else if (MODE == 2) {
#pragma unroll
			for (int i = 0; i < 8; i++) a[i] = exp2f(a[i]) ;   // MUFU.EX2 (+ range handling)
		} else if (MODE == 3) {
#pragma unroll
			for (int i = 0; i < 8; i++) a[i] = fminf(a[i], m) + 0.f * c;
		} else if (MODE == 4) {   // FFMA + integer ALU interleaved
#pragma unroll
			for (int i = 0; i < 8; i++) { a[i] = fmaf(a[i], m, c); acc = (acc ^ (acc >> 3)) + i; }
		} else if (MODE == 5) {   // FFMA2 + integer ALU interleaved
#pragma unroll
			for (int i = 0; i < 8; i++) { p[i] = fma2(p[i], m2, c2); acc = (acc ^ (acc >> 3)) + i; }
		} else if (MODE == 6) {   // ex2.approx only
#pragma unroll
			for (int i = 0; i < 8; i++) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a[i])); a[i] = r; }
		}
	}
	float r = 0;
#pragma unroll
	for (int i = 0; i < 8; i++) r += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
	out[blockIdx.x * blockDim.x + threadIdx.x] = r + acc;
}

template <int MODE>
void run(const char* name, int ops_per_iter_per_thread, int flop_lanes)
{
	const int grid = 148 * 8, iters = 4096;
	float* out;
	cudaMalloc(&out, grid * 256 * 4);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	k<MODE><<<grid, 256>>>(out, iters, 0.999f);
	cudaDeviceSynchronize();
	cudaEventRecord(e0);
	k<MODE><<<grid, 256>>>(out, iters, 0.999f);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double warp_inst = (double)grid * 8 * iters * ops_per_iter_per_thread;
	const double per_sm_clk = warp_inst / (ms * 1e-3) / 148 / 1.965e9;
	printf("%-28s %8.3f ms  %6.2f warp-inst/clk/SM (at 1965 MHz)  %7.1f G lane-results/s x%d\n", name, ms, per_sm_clk,
	       warp_inst * 32 * flop_lanes / (ms * 1e-3) / 1e9, flop_lanes);
	cudaFree(out);
}

int main()
{
	run<0>("FFMA  (8 chains)", 8, 1);
	run<1>("FFMA2 (8 chains x2 lanes)", 8, 2);
	run<2>("exp2f", 8, 1);
	run<6>("ex2.approx", 8, 1);
	run<3>("FMNMX+FFMA", 16, 1);
	run<4>("FFMA + 2 ALU", 24, 1);
	run<5>("FFMA2 + 2 ALU", 24, 1);
	return 0;
}
