// denorm.cu -- are FFMA / FMUL on denormal operands full rate on sm_100a, and exact as integer arithmetic?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void __launch_bounds__(1024) bench(float *out, long long *cyc, float a, float b, int mode)
{
	float f[8];
	for (int c = 0; c < 8; c++) f[c] = (mode == 0) ? (float)(threadIdx.x + c) : __uint_as_float((threadIdx.x + c) & 1023u);
	__syncthreads();
	const long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS / 8; it++) {
#pragma unroll
		for (int r = 0; r < 8; r++)
#pragma unroll
			for (int c = 0; c < 8; c++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(a), "f"(b));
	}
	const long long t1 = clock64();
	float acc = 0.f;
	for (int c = 0; c < 8; c++) acc += f[c];
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void exact(uint32_t *out)
{
	// integer arithmetic through denormal floats: (x * 32 + base), (pa * 8 + negative bias)
	uint32_t x = threadIdx.x, base = 70000u + blockIdx.x * 4;
	uint32_t r = __float_as_uint(fmaf(__uint_as_float(x), 32.0f, __uint_as_float(base)));
	uint32_t bias = 0x80000000u | 12345u;      // -12345 as a negative denormal
	uint32_t t = __float_as_uint(fmaf(__uint_as_float(r), 8.0f, __uint_as_float(bias)));
	float zf = (float)(threadIdx.x & 7);
	uint32_t z4 = __float_as_uint(fmaf(zf, __uint_as_float(4u), __uint_as_float(r)));
	out[(blockIdx.x * blockDim.x + threadIdx.x) * 3 + 0] = r - (x * 32 + base);
	out[(blockIdx.x * blockDim.x + threadIdx.x) * 3 + 1] = t - (r * 8 - 12345u);
	out[(blockIdx.x * blockDim.x + threadIdx.x) * 3 + 2] = z4 - (r + 4 * (threadIdx.x & 7));
}
int main()
{
	float *out; long long *cyc; uint32_t *eo;
	cudaMalloc(&out, 4 * 148 * 1024); cudaMalloc(&cyc, 8 * 148); cudaMalloc(&eo, 4 * 3 * 256 * 64);
	for (int mode = 0; mode < 2; mode++) {
		float a = mode ? 1.0f : 1.0000001f, b = mode ? __builtin_bit_cast(float, 3u) : 1e-9f;
		bench<<<148, 1024>>>(out, cyc, a, b, mode); cudaDeviceSynchronize();
		bench<<<148, 1024>>>(out, cyc, a, b, mode); cudaDeviceSynchronize();
		long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
		double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
		printf("FFMA %s operands: %.3f warp-instr/clk/SMSP\n", mode ? "denormal" : "normal", (double)ITERS * 8 * 8 / avg);
	}
	exact<<<64, 256>>>(eo); cudaDeviceSynchronize();
	static uint32_t h[3 * 256 * 64]; cudaMemcpy(h, eo, sizeof(h), cudaMemcpyDeviceToHost);
	long bad = 0; for (size_t i = 0; i < sizeof(h) / 4; i++) bad += h[i] != 0;
	printf("denormal integer arithmetic mismatches: %ld (%s)\n", bad, cudaGetErrorString(cudaGetLastError()));
	return 0;
}
