// sweep_common.cuh -- device helpers shared by the sweep kernels (zq_sweep.cu, tetra.cu):
// mbarrier / TMA bulk copy, cache-hinted global and shared-space accesses, the uniform draw,
// the FMA-pipe categorical search and denormal-float integer arithmetic.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace ig {

// --------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy (cp.async.bulk -> SASS UBLKCP), cache-hinted
// 128-bit global accesses, shared-space loads / reductions on 32-bit addresses.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "WAIT_LOOP:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	    "@p bra DONE;\n\t"
	    "bra WAIT_LOOP;\n\t"
	    "DONE:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 smem_addr(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
	             : "memory");
}
__device__ __forceinline__ int4 ldg_stream(const int4 *p)   // read-once data: bypass L1 allocation
{
	int4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ int4 ldg_rw(const int4 *p)       // data this kernel also writes: no .nc
{
	int4 r;
	asm volatile("ld.global.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ void stg_stream(int4 *p, const int4 &v)
{
	asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// the P chunk is read-only between the mbarrier wait and the end of the kernel: plain asm,
// so that ptxas may schedule these loads freely
__device__ __forceinline__ float4 lds_f4(uint32_t addr)
{
	float4 v;
	asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ float lds_f(uint32_t addr)
{
	float v;
	asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ void red_inc(uint32_t addr)
{
	asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ float lg2_fast(float x)           // MUFU.LG2; |abs err| <= 2^-22 on [0.5,2], 2 ulp elsewhere
{
	float r;
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

constexpr float BIG126 = 8.507059173023462e37f;          // 2^126
constexpr float U_SCALE = 8.507059173023462e37f;         // (f - 1 + 2^-24) * 2^126, f in [1,2)
constexpr float U_OFFS = -8.5070586659632355e37f;        // (-1 + 2^-24) * 2^126
constexpr float U_OFFS16 = -8.507059173023462e37f;       // 16-bit draw: the half step sits in the mantissa (bit 6), so the offset is -2^126
constexpr uint32_t U16_MANT = 0x007fff80u, U16_ONE = 0x3f800040u;   // RegConst of the 16-bit draw: mantissa bits 22..7, half step in bit 6

struct RegConst { uint32_t mant, one; };                 // 0x007fffff, 0x3f800000 held in registers

// uniform in (0,1) scaled by 2^126, from the low 23 bits of r: ((r & m) + 0.5) * 2^-23 * 2^126.
// One LOP3 builds the float 1.mantissa (the and-or needs its two constants in registers to
// stay ONE instruction), one FFMA rescales it.
__device__ __forceinline__ float uniform_big(uint32_t r, const RegConst &k)
{
	uint32_t b;
	asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(b) : "r"(r), "r"(k.mant), "r"(k.one));     // (r & mant) | one
	return fmaf(__uint_as_float(b), U_SCALE, U_OFFS);
}

// The bulk ancestry draw takes SIXTEEN random bits per allele copy, so that one Philox4x32 block serves four genotypes
// (eight copies): u = (r16 + 0.5) / 65536 scaled by 2^126.  Copy 0 of a genotype uses bits 22..7 of its word where they
// lie -- they are the mantissa -- and copy 1 the other sixteen, rotated into place by one PRMT (__byte_perm(r, 0, 0x1032)).
// k = {U16_MANT, U16_ONE}: (r & mant) | one = 1.r16|1000000b, i.e. 1 + u exactly.  A draw probability is therefore exact
// to 2^-16; where in a weight interval the 65536 grid points fall varies from locus to locus and individual to
// individual, so nothing is systematically favoured (chi-square tests against the exact conditional in tests/).
__device__ __forceinline__ float uniform_big16(uint32_t r, const RegConst &k)
{
	uint32_t b;
	asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(b) : "r"(r), "r"(k.mant), "r"(k.one));
	return fmaf(__uint_as_float(b), U_SCALE, U_OFFS16);
}

// index of the first cumulative weight that exceeds t, as a FLOAT in {0..KP-1}:
//     #{k < KP-1 : t > c_k} = sum_k sat((t - c_k) * 2^126)          (FFMA.SAT + FADD)
// t and c_k are fp32 values whose difference is 0 or at least one ulp(t) >= 2^-126 in
// magnitude (t >= 2^-24 * total, total >= P_FLOOR / K), so every term is exactly 0 or 1.
// Padded populations have q = 0, hence c_k = total > t: they never count.
template <int KP>
__device__ __forceinline__ float pick_category(const float (&c)[KP], float ub)
{
#ifdef IG_PICK_BSEARCH
	// The cumulative weights are non-decreasing, so the count #{k : t > c_k} is a three-level compare / select search for
	// KP = 8 -- identical decisions, fewer issue slots, but on the half-rate ALU pipe.  Defined by zq_sweep.cu only (measured
	// there: profiles/r2_zq_levers.md); the tetraploid passes and zq_snp.cu keep the saturating-FFMA form.
	if (KP == 8) {
		const float t = (ub * 1.1754943508222875e-38f) * c[KP - 1];          // 2^-126: a power of two, so t > c_k decides exactly as above
		const bool p2 = t > c[3];
		const float m1 = p2 ? c[5] : c[1];
		const bool p1 = t > m1;
		const float lo = p1 ? c[2] : c[0], hi = p1 ? c[6] : c[4];
		const float m0 = p2 ? hi : lo;
		const bool p0 = t > m0;
		float z = p2 ? 4.0f : 0.0f;
		z = p1 ? z + 2.0f : z;
		z = p0 ? z + 1.0f : z;
		return z;
	}
#endif
	const float tb = ub * c[KP - 1];                                      // t * 2^126, t = u * total
	float s[KP - 1];
#pragma unroll
	for (int k = 0; k < KP - 1; k++) s[k] = __saturatef(fmaf(c[k], -BIG126, tb));
	// balanced tree: shorter dependency chain than a running sum
#pragma unroll
	for (int w = 1; w < KP - 1; w <<= 1)
#pragma unroll
		for (int k = 0; k + w < KP - 1; k += 2 * w) s[k] += s[k + w];
	return s[0];
}

// Small non-negative integers (shared-memory addresses < 2^23, allele and population indices)
// are the bit patterns of denormal floats, and FFMA on denormals is exact integer arithmetic
// at full FMA-pipe rate (tools/ubench/denorm.cu).  All address arithmetic of the inner loop
// is phrased that way: it costs one FFMA where the integer form costs LEA / IADD3 on the
// half-rate ALU pipe, which is this kernel's second-busiest resource.
__device__ __forceinline__ float as_dn(uint32_t v) { return __uint_as_float(v); }
__device__ __forceinline__ float as_dn_signed(int v) { return __uint_as_float(v >= 0 ? (uint32_t)v : (0x80000000u | (uint32_t)(-v))); }

}  // namespace ig
