// pipes.cu -- issue-rate microbenchmarks for the instruction mix of zq_sweep on sm_100a.
// Each test runs NCH independent dependency chains per thread inside an unrolled loop and
// reports warp-instructions per clock per SM sub-partition (SMSP), from clock64() inside
// the kernel (one CTA per SM, W warps per SMSP).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define NCH 8

typedef unsigned long long u64;

template <int T>
__device__ __forceinline__ void body(float (&f)[NCH], u64 (&d)[NCH], uint32_t (&u)[NCH], float k1, float k2, uint32_t m)
{
#pragma unroll
	for (int c = 0; c < NCH; c++) {
		if (T == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(k1), "f"(k2));
		if (T == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[c]) : "l"(d[(c + 1) % NCH]), "l"(d[(c + 2) % NCH]));
		if (T == 2) asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(k1), "f"(k2));
		if (T == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[c]) : "l"(d[(c + 1) % NCH]));
		if (T == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(m), "r"(u[(c + 1) % NCH]));
		if (T == 5) { uint32_t lo, hi; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(d[c]) : "r"(u[c]), "r"(m)); lo = (uint32_t)d[c]; hi = (uint32_t)(d[c] >> 32); u[c] = lo ^ hi; }
		if (T == 6) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(k1), "f"(k2)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(m), "r"(u[(c + 1) % NCH])); }
		if (T == 7) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[c]) : "l"(d[(c + 1) % NCH]), "l"(d[(c + 2) % NCH])); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(m), "r"(u[(c + 1) % NCH])); }
		if (T == 8) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(f[(c + 1) % NCH]), "f"(f[(c + 2) % NCH]));   // 3 distinct regs
		if (T == 9) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(m), "r"(u[(c + 1) % NCH]));
		if (T == 10) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[c]) : "f"(f[(c + 1) % NCH]));
		if (T == 11) { asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %2; selp.f32 %0, %1, %2, p;}" : "=f"(f[c]) : "f"(f[c]), "f"(f[(c + 1) % NCH])); }
		if (T == 12) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(m), "r"(u[(c + 1) % NCH]));
		if (T == 13) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[c]) : "l"(d[(c + 1) % NCH]), "l"(d[(c + 2) % NCH])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(k1), "f"(k2)); }
		if (T == 14) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(d[c]) : "r"(u[c]), "r"(m)); u[c] = (uint32_t)d[c] ^ (uint32_t)(d[c] >> 32); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[(c + 3) % NCH]) : "l"(d[(c + 1) % NCH]), "l"(d[(c + 2) % NCH])); }
		if (T == 15) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(m), "r"(u[(c + 1) % NCH]));
		if (T == 16) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(u[(c + 1) % NCH]), "r"(m));
		if (T == 17) { asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(u[c]) : "f"(f[c])); f[c] = __uint_as_float(u[c] | 0x3f000000u); }
		if (T == 18) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(u[(c + 1) % NCH]), "r"(m));
		if (T == 20) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(k1), "f"(k2)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(k1), "f"(k2)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(m), "r"(u[(c + 1) % NCH])); }
		if (T == 21) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[c]) : "l"(d[(c + 1) % NCH]), "l"(d[(c + 2) % NCH])); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[c]) : "r"(m), "r"(u[(c + 1) % NCH])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(k1), "f"(k2)); }
		if (T == 22) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[c]) : "r"(m), "r"(u[(c + 1) % NCH])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(k1), "f"(k2)); }
		if (T == 23) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(d[c]) : "r"(u[c]), "r"(m)); u[c] = (uint32_t)d[c]; u[(c+1)%NCH] ^= (uint32_t)(d[c] >> 32);  asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(k1), "f"(k2)); }
		if (T == 19) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[c]) : "r"(u[(c + 1) % NCH]));
	}
}

template <int T>
__global__ void __launch_bounds__(1024) bench(float *out, long long *cyc, float k1, float k2, uint32_t m)
{
	float f[NCH];
	u64 d[NCH];
	uint32_t u[NCH];
#pragma unroll
	for (int c = 0; c < NCH; c++) {
		f[c] = (float)(threadIdx.x + c) * 1e-3f;
		u[c] = threadIdx.x * 2654435761u + c;
		float2 t = make_float2(f[c], f[c] * 0.5f);
		d[c] = *reinterpret_cast<u64 *>(&t);
	}
	__syncthreads();
	const long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS / 8; it++) {
#pragma unroll
		for (int r = 0; r < 8; r++) body<T>(f, d, u, k1, k2, m);
	}
	const long long t1 = clock64();
	__syncthreads();
	float acc = 0.f;
#pragma unroll
	for (int c = 0; c < NCH; c++) acc += f[c] + (float)u[c] + (float)(uint32_t)d[c] + (float)(uint32_t)(d[c] >> 32);
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static const char *names[] = {"FFMA (imm/const operands)", "FFMA2", "FFMA.SAT", "FADD2", "LOP3", "IMAD.WIDE + LOP3(xor)", "FFMA + LOP3 pair", "FFMA2 + LOP3 pair",
                              "FFMA 3-reg", "IMAD.LO", "FMNMX", "FSETP+FSEL pair", "IMAD.HI", "FFMA2 + FFMA pair", "IMAD.WIDE + LOP3 + FFMA2", "HFMA2", "PRMT", "F2I+LOP", "SHF", "IADD", "2 FFMA + LOP3", "FFMA2 + LOP3 + FFMA", "IMAD.LO + FFMA", "IMAD.WIDE + LOP3 + FFMA"};
static const int per_iter[] = {1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 2, 1, 2, 3, 1, 2, 1, 1, 3, 3, 2, 3};

template <int T>
void run(int threads, float *out, long long *cyc, int sms)
{
	bench<T><<<sms, threads>>>(out, cyc, 1.0000001f, 1e-9f, 0x9E3779B9u);
	cudaDeviceSynchronize();
	bench<T><<<sms, threads>>>(out, cyc, 1.0000001f, 1e-9f, 0x9E3779B9u);
	cudaDeviceSynchronize();
	long long h[512];
	cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
	double avg = 0;
	for (int i = 0; i < sms; i++) avg += (double)h[i];
	avg /= sms;
	const double warps_per_smsp = threads / 32 / 4.0;
	const double instr = (double)ITERS * NCH * per_iter[T] * warps_per_smsp;
	printf("%-28s warps/SMSP=%4.1f  %.3f warp-instr/clk/SMSP\n", names[T], warps_per_smsp, instr / avg);
}

int main()
{
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	float *out;
	long long *cyc;
	cudaMalloc(&out, sizeof(float) * sms * 1024);
	cudaMalloc(&cyc, sizeof(long long) * 512);
	for (int threads : {512, 1024}) {
		run<0>(threads, out, cyc, sms); run<8>(threads, out, cyc, sms); run<1>(threads, out, cyc, sms); run<2>(threads, out, cyc, sms);
		run<3>(threads, out, cyc, sms); run<4>(threads, out, cyc, sms); run<5>(threads, out, cyc, sms); run<6>(threads, out, cyc, sms);
		run<7>(threads, out, cyc, sms); run<9>(threads, out, cyc, sms); run<12>(threads, out, cyc, sms); run<10>(threads, out, cyc, sms);
		run<11>(threads, out, cyc, sms); run<13>(threads, out, cyc, sms); run<14>(threads, out, cyc, sms); run<15>(threads, out, cyc, sms);
		run<16>(threads, out, cyc, sms); run<17>(threads, out, cyc, sms); run<18>(threads, out, cyc, sms); run<19>(threads, out, cyc, sms); run<20>(threads, out, cyc, sms); run<21>(threads, out, cyc, sms); run<22>(threads, out, cyc, sms); run<23>(threads, out, cyc, sms);
	}
	printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
