// zq_snp.cu -- the sweep pass for biallelic data (allelenum_max == 2: the SNP-scale configuration, BASELINE configs[3]).
//
// Same work as zq_sweep.cu -- update_ZQ's Z draw (mcmc.c:1133-1174), the tally of the next update_P (:810-845), the
// ancestry counts (:1176-1194) and the log_ld_indv pieces of update_G / cal_lkh (:1726-1773, :1053-1091, :1916-1942),
// in one pass that reads 2 B of genotype + 1 B of old z and writes 1 B of new z per allele copy -- on a different
// decomposition, chosen from the kernel-lab measurements of round 2 (profiles/r2_zq_levers.md):
//
//   * A WARP WALKS ONE INDIVIDUAL, lanes are loci.  The individual's genotypes inside a chunk of TLC loci are stored
//     sorted by class -- homozygotes, then heterozygotes, then missing -- so every row of 32 genotypes but the two at the
//     class boundaries is pure, and the class is a warp-uniform comparison with two counts:
//       - a homozygote's two copies share their P row and their cumulative weights: one row fetch and one prefix sum
//         instead of two (measured on the generic kernel with every genotype forced homozygous: 3.99 -> 3.38 ms);
//       - a heterozygote never has a generation-dependent term (genofreq, mcmc.c:1700): no old-Z piece, no second
//         likelihood under g' (forced heterozygous: 3.99 -> 3.45 ms).
//     The sort is a per-(individual, chunk) permutation of loci made once at load; the store entry carries the locus, so
//     Z lives in the same permuted order and the canonical view is rebuilt only for the state hooks.
//   * lanes on different loci hit different tally bins: the shared-memory histogram needs no replicas and its atomics no
//     conflict replays (105 M bank-conflict wavefronts per launch in the generic kernel's capture); with A = 2 the chunk
//     is 1024 loci (98 chunks at config 4 instead of 569), so the per-chunk partials shrink by the same factor.
//   * shared memory holds P twice, both landed by ONE TMA bulk copy from a chunk-ordered copy that p_dirichlet writes:
//     as float4 planes [k/4][x][l] for the two 128-bit row loads of a copy, and as words [k][x][l] -- locus fastest,
//     like the histogram [k][x][l] -- so that the f lookups and the tally RED of a warp fall into 32 different banks.
//     A store entry is (x0 * TLC + l) | (x1 * TLC + l) << 16: every shared address of a copy is that offset times a
//     constant plus a base (+ z times the k stride).
//
// Per-(individual, chunk) sums are reduced across the warp in a fixed butterfly order and written as the same partials
// zq_sweep.cu writes, so indiv_epilogue is shared.  Randomness: Philox4x32-7 block (chunk step, global individual, sweep,
// TAG_ZS | lane), 16 bits per copy -- a pure function of the chunk decomposition, hence shard-invariant.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "ig_internal.h"
#include "philox.cuh"
#include "sweep_common.cuh"

namespace ig {

#define LN2_D 0.69314718055994530942

template <int KP, int TLC>
struct SnpLayout {
	static constexpr uint32_t PLANE_BYTES = 2u * TLC * 16u;                 // one plane: float4 [2][TLC]
	static constexpr uint32_t PLANES_BYTES = (KP / 4) * PLANE_BYTES;
	static constexpr uint32_t KSTRIDE = 2u * TLC * 4u;                      // bytes between consecutive k in Pf / hist
	static constexpr uint32_t PF_OFF = PLANES_BYTES;
	static constexpr uint32_t PF_BYTES = KP * KSTRIDE;
	static constexpr uint32_t HIST_OFF = PF_OFF + PF_BYTES;
	static constexpr uint32_t HIST_DELTA = PF_BYTES;                        // &hist[k][x][l] - &Pf[k][x][l]
	static constexpr uint32_t CNT_OFF = HIST_OFF + PF_BYTES;
	static constexpr uint32_t TOTAL = CNT_OFF + KP * SNP_THREADS * 4u;
	static constexpr uint32_t PC_BLOCK_BYTES = PLANES_BYTES + PF_BYTES;     // one chunk of the global chunk-ordered P
};

struct SnpAcc {
	float mA, mB, mH, mD;      // log2 sums: homozygotes on the new Z under g and under g', heterozygotes (common to both), old-Z ratio
	int nsh;                   // same-z heterozygotes on the new Z
};
struct SnpThr {
	float plane0;              // shared address of plane 0, as a denormal (sweep_common.cuh)
	float pf;                  // shared address of Pf[0][0][0]
	float cnt_t;               // shared address of this thread's counter column
	float omh_g, h_g, omh_p, h_p;
};

template <uint32_t OFF>
__device__ __forceinline__ void red_inc_at(uint32_t addr)
{
	asm volatile("red.shared.add.u32 [%0+%1], 1;" ::"r"(addr), "n"(OFF) : "memory");
}

// One genotype of the lane.  e = (x0 TLC + l) | (x1 TLC + l) << 16, zw = old z0 | z1 << 8, r = 32 random bits.
// CLS 0: every lane of the row is a homozygote; 1: every lane a heterozygote; 2: anything (class from the entry).
template <int KP, int TLC, int CLS>
__device__ __forceinline__ uint32_t snp_genotype(uint32_t e, uint32_t zw, uint32_t r, const float (&q)[KP], const SnpThr &t, SnpAcc &acc,
                                                 const RegConst &kc)
{
	using LY = SnpLayout<KP, TLC>;
	const uint32_t o0 = __byte_perm(e, 0u, 0x4410);
	const float o0f = as_dn(o0);
	const uint32_t prow0 = __float_as_uint(fmaf(o0f, 16.0f, t.plane0));
	const float fb0f = fmaf(o0f, 4.0f, t.pf);
	float c0[KP], c1[KP];
	{
		float p0[KP];
#pragma unroll
		for (int v = 0; v < KP / 4; v++) {
			const float4 t0 = lds_f4(prow0 + v * LY::PLANE_BYTES);
			p0[4 * v] = t0.x; p0[4 * v + 1] = t0.y; p0[4 * v + 2] = t0.z; p0[4 * v + 3] = t0.w;
		}
		c0[0] = q[0] * p0[0];
#pragma unroll
		for (int k = 1; k < KP; k++) c0[k] = fmaf(q[k], p0[k], c0[k - 1]);
	}
	float fb1f;
	bool het;
	if (CLS == 0) {
		het = false;
		fb1f = fb0f;
#pragma unroll
		for (int k = 0; k < KP; k++) c1[k] = c0[k];
	} else {
		const uint32_t o1 = __byte_perm(e, 0u, 0x4432);
		const float o1f = as_dn(o1);
		const uint32_t prow1 = __float_as_uint(fmaf(o1f, 16.0f, t.plane0));
		fb1f = fmaf(o1f, 4.0f, t.pf);
		het = (CLS == 1) ? true : (o0 != o1);
		float p1[KP];
#pragma unroll
		for (int v = 0; v < KP / 4; v++) {
			const float4 t1 = lds_f4(prow1 + v * LY::PLANE_BYTES);
			p1[4 * v] = t1.x; p1[4 * v + 1] = t1.y; p1[4 * v + 2] = t1.z; p1[4 * v + 3] = t1.w;
		}
		c1[0] = q[0] * p1[0];
#pragma unroll
		for (int k = 1; k < KP; k++) c1[k] = fmaf(q[k], p1[k], c1[k - 1]);
	}
	// ---- old-Z piece of update_G's ratio (log_ld_indv, mcmc.c:1752-1759): same-z homozygotes only
	if (CLS != 1) {
		const uint32_t zo0 = __byte_perm(zw, 0u, 0x4440), zo1 = __byte_perm(zw, 0u, 0x4441);
		const float fo = lds_f(__float_as_uint(fmaf(as_dn(zo0), (float)LY::KSTRIDE, fb0f)));
		const float fe = (zo0 == zo1 && !het) ? fo : 1.0f;                // h + 1 * (1 - h) == 1 exactly
		acc.mD += lg2_fast(fmaf(fe, t.omh_p, t.h_p)) - lg2_fast(fmaf(fe, t.omh_g, t.h_g));
	}
	// ---- categorical draws (disc_unif, random.c:403-430), 16 random bits each
	const float zf0 = pick_category<KP>(c0, uniform_big16(r, kc));
	const float zf1 = pick_category<KP>(c1, uniform_big16(__byte_perm(r, 0u, 0x1032), kc));
	const uint32_t pa0 = __float_as_uint(fmaf(zf0, as_dn(LY::KSTRIDE), fb0f));          // &Pf[z0][x0][l]
	const uint32_t pa1 = __float_as_uint(fmaf(zf1, as_dn(LY::KSTRIDE), fb1f));
	// ---- n[l][a][k] for the next update_P (mcmc.c:815-845) and the ancestry counts (mcmc.c:1176-1194)
	red_inc_at<LY::HIST_DELTA>(pa0);
	red_inc_at<LY::HIST_DELTA>(pa1);
	red_inc(__float_as_uint(fmaf(zf0, as_dn(4u * SNP_THREADS), t.cnt_t)));
	red_inc(__float_as_uint(fmaf(zf1, as_dn(4u * SNP_THREADS), t.cnt_t)));
	// ---- new-Z likelihood pieces (cal_lkh and the accepted-G selection)
	const float f0 = lds_f(pa0), f1 = lds_f(pa1);
	const bool same_n = (zf0 == zf1);
	if (CLS == 1) {
		acc.mH += lg2_fast(f0 * f1);
		acc.nsh += same_n ? 1 : 0;
	} else if (CLS == 0) {
		acc.mA += lg2_fast(f0 * (same_n ? fmaf(f0, t.omh_g, t.h_g) : f1));
		acc.mB += lg2_fast(f0 * (same_n ? fmaf(f0, t.omh_p, t.h_p) : f1));
	} else {
		const bool sh_n = same_n && !het;
		acc.mA += lg2_fast(f0 * (sh_n ? fmaf(f0, t.omh_g, t.h_g) : f1));
		acc.mB += lg2_fast(f0 * (sh_n ? fmaf(f0, t.omh_p, t.h_p) : f1));
		acc.nsh += (same_n && het) ? 1 : 0;
	}
	return __float_as_uint(fmaf(zf1, as_dn(256u), zf0 * as_dn(1u)));                    // z0 | z1 << 8
}

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ uint2 ldg_rw_u2(const uint2 *p)
{
	uint2 r;
	asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
	return r;
}
__device__ __forceinline__ void stg_stream_u2(uint2 *p, uint32_t a, uint32_t b)
{
	asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}

// One step: 128 genotypes of one individual, 4 rows of 32 (this lane's genotypes are e.x .. e.w).  Rows of one class take
// the straight-line specialised bodies; the (at most two per individual and chunk) steps that hold a class boundary walk
// their rows one by one.
template <int KP, int TLC, int ROUNDS>
__device__ __forceinline__ void snp_step(int s, const uint4 &e, const uint2 &z, uint2 *zdst, int nhom, int nhh, uint32_t site, uint32_t ig_global,
                                         uint32_t iter, int lane, uint32_t key0, uint32_t key1, const float (&q)[KP], const SnpThr &t, SnpAcc &acc,
                                         const RegConst &kc, float &lgA, float &lgB, float &lgH, float &lgD)
{
	const int p0 = 128 * s;
	if (p0 >= nhh) return;                                            // only missing genotypes: their z stays (mcmc.c:1137)
	const u32x4 rnd = philox4x32<ROUNDS>(u32x4{site, ig_global, iter, TAG_ZS | (uint32_t)lane}, key0, key1);
	acc.mA = acc.mB = acc.mH = acc.mD = 0.0f;
	uint32_t zlo, zhi;
	if (p0 + 128 <= nhom) {
		const uint32_t a0 = snp_genotype<KP, TLC, 0>(e.x, z.x & 0xFFFFu, rnd.x, q, t, acc, kc);
		const uint32_t a1 = snp_genotype<KP, TLC, 0>(e.y, z.x >> 16, rnd.y, q, t, acc, kc);
		const uint32_t a2 = snp_genotype<KP, TLC, 0>(e.z, z.y & 0xFFFFu, rnd.z, q, t, acc, kc);
		const uint32_t a3 = snp_genotype<KP, TLC, 0>(e.w, z.y >> 16, rnd.w, q, t, acc, kc);
		zlo = a0 | (a1 << 16); zhi = a2 | (a3 << 16);
	} else if (p0 >= nhom && p0 + 128 <= nhh) {
		const uint32_t a0 = snp_genotype<KP, TLC, 1>(e.x, z.x & 0xFFFFu, rnd.x, q, t, acc, kc);
		const uint32_t a1 = snp_genotype<KP, TLC, 1>(e.y, z.x >> 16, rnd.y, q, t, acc, kc);
		const uint32_t a2 = snp_genotype<KP, TLC, 1>(e.z, z.y & 0xFFFFu, rnd.z, q, t, acc, kc);
		const uint32_t a3 = snp_genotype<KP, TLC, 1>(e.w, z.y >> 16, rnd.w, q, t, acc, kc);
		zlo = a0 | (a1 << 16); zhi = a2 | (a3 << 16);
	} else {
		zlo = z.x; zhi = z.y;
#pragma unroll 1
		for (int j = 0; j < 4; j++) {
			const uint32_t ej = j == 0 ? e.x : (j == 1 ? e.y : (j == 2 ? e.z : e.w));
			const uint32_t rj = j == 0 ? rnd.x : (j == 1 ? rnd.y : (j == 2 ? rnd.z : rnd.w));
			const uint32_t zw = j < 2 ? z.x : z.y;
			const uint32_t zo = (j & 1) ? (zw >> 16) : (zw & 0xFFFFu);
			const int r0 = p0 + 32 * j;
			uint32_t zn = zo;
			if (r0 + 32 <= nhom) zn = snp_genotype<KP, TLC, 0>(ej, zo, rj, q, t, acc, kc);
			else if (r0 >= nhom && r0 + 32 <= nhh) zn = snp_genotype<KP, TLC, 1>(ej, zo, rj, q, t, acc, kc);
			else if (r0 < nhh) { if ((int)ej >= 0) zn = snp_genotype<KP, TLC, 2>(ej, zo, rj, q, t, acc, kc); }
			const uint32_t sh = (j & 1) ? 16u : 0u, keep = (j & 1) ? 0x0000FFFFu : 0xFFFF0000u;
			if (j < 2) zlo = (zlo & keep) | (zn << sh);
			else zhi = (zhi & keep) | (zn << sh);
		}
	}
	stg_stream_u2(zdst, zlo, zhi);
	lgA += acc.mA; lgB += acc.mB; lgH += acc.mH; lgD += acc.mD;
}

// --------------------------------------------------------------------------------------
// zq_snp: grid (chunks, individual blocks), 512 threads = 16 warps, one CTA per SM.
//   global : Es u32 [chunk][Nloc][TLC]  class-sorted entries, 128-blocks stored [lane][4]: one 128-bit load per lane per step
//            Zs u16 [chunk][Nloc][TLC]  same order: one 64-bit load and one 64-bit store per lane per step
//            Hs u32 [chunk][Nloc]       homozygotes | heterozygotes << 16 of the list
//            Pc     [chunk][planes | words]  chunk-ordered P (written by p_dirichlet), one TMA bulk copy
// --------------------------------------------------------------------------------------
template <int KP, int TLC, int ROUNDS>
__global__ void __launch_bounds__(SNP_THREADS, 1) zq_snp_kernel(const SnpArgs a)
{
	using LY = SnpLayout<KP, TLC>;
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) unsigned long long bar;
	const Geometry &g = a.geo;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int chunk = blockIdx.x;
	const int l0 = chunk * TLC;

	if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
	__syncthreads();
	if (tid == 0) {
		mbar_expect_tx(&bar, LY::PC_BLOCK_BYTES);
		tma_bulk_g2s(smem, reinterpret_cast<const unsigned char *>(a.Pc) + (size_t)chunk * LY::PC_BLOCK_BYTES, LY::PC_BLOCK_BYTES, &bar);
	}
	{
		uint4 *hz = reinterpret_cast<uint4 *>(smem + LY::HIST_OFF);
		for (int j = tid; j < (int)((LY::PF_BYTES + KP * SNP_THREADS * 4u) / 16u); j += SNP_THREADS) hz[j] = make_uint4(0u, 0u, 0u, 0u);
	}
	__syncthreads();
	mbar_wait(&bar, 0);

	const uint32_t sbase = smem_addr(smem);
	const RegConst kc{a.k_mant, a.k_one};
	const uint32_t iter = a.iter_dev ? *a.iter_dev : a.iter;
	SnpThr t;
	t.plane0 = as_dn(sbase);
	t.pf = as_dn(sbase + LY::PF_OFF);
	t.cnt_t = as_dn(sbase + LY::CNT_OFF + (uint32_t)tid * 4u);
	int *cnt_col = reinterpret_cast<int *>(smem + LY::CNT_OFF) + tid;
	const int Nloc = g.Nloc;
	const int ib0 = blockIdx.y * g.subs_per_blk;
	const int ib1 = min(ib0 + g.subs_per_blk, Nloc);
	constexpr int NSTEP = TLC / 128;
	// The warp's individuals are il(k) = ib0 + wid + 16 k, each NSTEP steps of 128 genotypes (4 rows).  Two steps are in
	// flight ahead of the arithmetic, across individual boundaries (one CTA of 16 warps per SM: one step ahead leaves 12 KB
	// in flight per SM, half of what HBM latency x this kernel's bandwidth asks for): the loop body is a PAIR of steps, the
	// next pair is requested at its top and only moved into place at its bottom.  Pointers advance by constants.
	constexpr int WARPS = SNP_THREADS / 32;
	constexpr int ESTEP = 32;                                  // one step, in uint4 of entries = in uint2 of z
	constexpr int EIND = WARPS * TLC / 4;                      // this warp's next individual
	const int il_first = ib0 + wid;
	const int nk = il_first < ib1 ? (ib1 - il_first + WARPS - 1) / WARPS : 0;
	if (nk > 0) {
		const uint4 *ep = reinterpret_cast<const uint4 *>(a.Es + ((size_t)chunk * Nloc + il_first) * TLC) + lane;
		uint2 *zp = reinterpret_cast<uint2 *>(a.Zs + ((size_t)chunk * Nloc + il_first) * TLC) + lane;
		uint4 eA = ldg_stream_u4(ep), eB = ldg_stream_u4(ep + ESTEP);
		uint2 zA = ldg_rw_u2(zp), zB = ldg_rw_u2(zp + ESTEP);
		const float4 *qp = reinterpret_cast<const float4 *>(a.Qf + (size_t)il_first * KP);
		float4 qn0 = __ldg(qp), qn1 = (KP == 8) ? __ldg(qp + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
		int2 ggn = __ldg(a.gpair + il_first);
		uint32_t hdrn = __ldg(a.Hs + (size_t)chunk * Nloc + il_first);
		float q[KP];
		SnpAcc acc;
		for (int k = 0; k < nk; ++k) {
			const int il = il_first + k * WARPS;
			q[0] = qn0.x; q[1] = qn0.y; q[2] = qn0.z; q[3] = qn0.w;
			if (KP == 8) { q[KP - 4] = qn1.x; q[KP - 3] = qn1.y; q[KP - 2] = qn1.z; q[KP - 1] = qn1.w; }
			const int2 gg = ggn;
			// 1 - h(g) = 2^-(g-1); exact in fp32 down to 2^-126, 0 beyond
			t.omh_g = (gg.x <= 127) ? __int_as_float((128 - gg.x) << 23) : 0.0f;
			t.omh_p = (gg.y <= 127) ? __int_as_float((128 - gg.y) << 23) : 0.0f;
			t.h_g = 1.0f - t.omh_g;
			t.h_p = 1.0f - t.omh_p;
			const int nhom = (int)(hdrn & 0xFFFFu);                    // end of the homozygotes
			const int nhh = nhom + (int)(hdrn >> 16);                 // end of the heterozygotes; only missing genotypes behind it
			const bool more = k + 1 < nk;
			const uint32_t ig_global = (uint32_t)(g.i0 + il);
			acc.nsh = 0;
			float lgA = 0.0f, lgB = 0.0f, lgH = 0.0f, lgD = 0.0f;
#pragma unroll 1
			for (int sp = 0; sp < NSTEP / 2; ++sp) {
				const bool last = (sp == NSTEP / 2 - 1);
				const int adv = last ? EIND - (NSTEP - 2) * ESTEP : 2 * ESTEP;
				uint4 nA = eA, nB = eB;
				uint2 mA = zA, mB = zB;
				if (!last || more) {
					nA = ldg_stream_u4(ep + adv); nB = ldg_stream_u4(ep + adv + ESTEP);
					mA = ldg_rw_u2(zp + adv); mB = ldg_rw_u2(zp + adv + ESTEP);
				}
				if (last && more) {
					const float4 *qq = reinterpret_cast<const float4 *>(a.Qf + (size_t)(il + WARPS) * KP);
					qn0 = __ldg(qq);
					if (KP == 8) qn1 = __ldg(qq + 1);
					ggn = __ldg(a.gpair + il + WARPS);
					hdrn = __ldg(a.Hs + (size_t)chunk * Nloc + il + WARPS);
				}
				snp_step<KP, TLC, ROUNDS>(2 * sp, eA, zA, zp, nhom, nhh, (uint32_t)(chunk * NSTEP + 2 * sp), ig_global, iter, lane, a.key0, a.key1, q, t, acc, kc, lgA, lgB, lgH, lgD);
				snp_step<KP, TLC, ROUNDS>(2 * sp + 1, eB, zB, zp + ESTEP, nhom, nhh, (uint32_t)(chunk * NSTEP + 2 * sp + 1), ig_global, iter, lane, a.key0, a.key1, q, t, acc, kc, lgA, lgB, lgH, lgD);
				eA = nA; eB = nB; zA = mA; zB = mB;
				ep += adv; zp += adv;
			}
			// ---- this (chunk, individual) is complete: reduce over the lanes in a fixed order, lane 0 writes the partials
			uint32_t pk[KP / 2];
#pragma unroll
			for (int j = 0; j < KP / 2; j++) {
				const int ca = cnt_col[(2 * j) * SNP_THREADS], cb = cnt_col[(2 * j + 1) * SNP_THREADS];
				cnt_col[(2 * j) * SNP_THREADS] = 0;
				cnt_col[(2 * j + 1) * SNP_THREADS] = 0;
				pk[j] = __reduce_add_sync(0xffffffffu, (uint32_t)ca | ((uint32_t)cb << 16));     // <= 2 TLC per field
			}
			const int nsh = (int)__reduce_add_sync(0xffffffffu, (uint32_t)acc.nsh);
			const double dA = warp_sum_d((double)lgA + (double)lgH), dB = warp_sum_d((double)lgB + (double)lgH), dD = warp_sum_d((double)lgD);
			if (lane == 0) {
				uint32_t *pc = reinterpret_cast<uint32_t *>(a.pcnt + ((size_t)chunk * Nloc + il) * KP);
#pragma unroll
				for (int j = 0; j < KP / 2; j++) pc[j] = pk[j];
				double *pl = a.plog + (size_t)chunk * 3 * Nloc + il;
				pl[0] = dD * LN2_D;
				pl[(size_t)Nloc] = dA * LN2_D - (double)nsh * (double)(gg.x - 1) * LN2_D;
				pl[(size_t)2 * Nloc] = dB * LN2_D - (double)nsh * (double)(gg.y - 1) * LN2_D;
				a.pnsh[(size_t)chunk * Nloc + il] = (uint16_t)nsh;
			}
		}
	}
	__syncthreads();
	// ---- push this CTA's tally into global n[l][x][k] (RED); eight consecutive threads fill one 32-byte sector
	const int nl = min(TLC, g.Lpad - l0);
	int32_t *ng = a.n + (size_t)l0 * 2 * KP;
	const int *hist = reinterpret_cast<const int *>(smem + LY::HIST_OFF);
	for (int b = tid; b < nl * 2 * KP; b += SNP_THREADS) {
		const int k = b % KP, lx = b / KP, x = lx & 1, l = lx >> 1;
		const int v = hist[(k * 2 + x) * TLC + l];
		if (v) atomicAdd(ng + b, v);
	}
}

template <int KP, int TLC>
static cudaError_t launch_snp_t(const SnpArgs &a, int rounds, cudaStream_t s)
{
	using LY = SnpLayout<KP, TLC>;
	dim3 grid(a.geo.nchunks, a.geo.nblk), block(SNP_THREADS);
	cudaError_t e;
	if (rounds == 7) {
		e = cudaFuncSetAttribute(zq_snp_kernel<KP, TLC, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LY::TOTAL);
		if (e != cudaSuccess) return e;
		zq_snp_kernel<KP, TLC, 7><<<grid, block, LY::TOTAL, s>>>(a);
	} else {
		e = cudaFuncSetAttribute(zq_snp_kernel<KP, TLC, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LY::TOTAL);
		if (e != cudaSuccess) return e;
		zq_snp_kernel<KP, TLC, 10><<<grid, block, LY::TOTAL, s>>>(a);
	}
	return cudaGetLastError();
}

cudaError_t launch_zq_snp(const SnpArgs &a, int rounds, cudaStream_t s)
{
	const int key = a.geo.KP * 10000 + a.geo.TL;
	switch (key) {
	case 8 * 10000 + 1024: return launch_snp_t<8, 1024>(a, rounds, s);
	case 8 * 10000 + 256: return launch_snp_t<8, 256>(a, rounds, s);
#ifndef IG_FAST_BUILD
	case 4 * 10000 + 1024: return launch_snp_t<4, 1024>(a, rounds, s);
	case 4 * 10000 + 256: return launch_snp_t<4, 256>(a, rounds, s);
#endif
	default: return cudaErrorInvalidValue;
	}
}

// Decide whether this context takes the biallelic path and, if so, its decomposition.  The chunk length is a compile-time
// constant (the shared-memory strides are immediates): 1024 loci, or 256 for small data sets.
bool snp_eligible(const Geometry &g, int mode, int type_freq)
{
#ifdef IG_FAST_BUILD
	if (g.KP != 8) return false;
#endif
	// Opt-in (IG_SNP_PATH=1): at config 4 this kernel runs 4.14 ms per launch against 3.95 ms for zq_sweep with the same
	// 16-bit draw -- the class-pure rows cost 84 / 81 warp instructions per 32 genotypes as designed (generic: 116), but the
	// per-step, per-pair and per-individual work that the lanes-are-loci decomposition adds (Philox per lane and step,
	// cross-lane reductions per individual and chunk, rows at the class boundaries) takes 26 more per row
	// (profiles/r2_zq_snp_ncu.md).  Kept, tested and measured; not the default.
	if (!getenv("IG_SNP_PATH")) return false;
	return g.A == 2 && (g.KP == 8 || g.KP == 4) && g.fmode == 0 && type_freq == 1 && mode >= 1 && mode <= 3;
}

cudaError_t snp_configure(Geometry &g, int device)
{
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
	const int tlc = g.Lpad > 512 ? 1024 : 256;
	g.snp = 1;
	g.TL = tlc;
	g.nchunks = (g.Lpad + tlc - 1) / tlc;
	// individual blocks: enough CTAs for several full waves (one CTA per SM), the last wave as full as possible, and enough
	// individuals per CTA to amortise its prologue (128 KB of P, zeroed histogram) and the tally flush
	const double overhead = 24.0 * tlc;                       // prologue + flush, in genotype-equivalents (about 4 us at TLC = 1024)
	int best = 1;
	double best_eff = -1.0;
	const int max_blk = (g.Nloc + 15) / 16 < 64 ? (g.Nloc + 15) / 16 : 64;
	for (int nb = 1; nb <= (max_blk < 1 ? 1 : max_blk); nb++) {
		const int ipb = (g.Nloc + nb - 1) / nb;
		const int nblk = (g.Nloc + ipb - 1) / ipb;
		const long ctas = (long)g.nchunks * nblk;
		const long waves = (ctas + sms - 1) / sms;
		const double fill = (double)ctas / (double)(waves * sms);
		const double work = (double)ipb * tlc;
		// many waves also even out the partly filled last chunk and unequal lists
		const double eff = fill * work / (work + overhead) * (waves >= 6 ? 1.0 : 0.94 + 0.01 * waves);
		if (eff > best_eff) { best_eff = eff; best = nblk; }
	}
	if (const char *e = getenv("IG_SNP_NBLK")) { const int v = atoi(e); if (v >= 1) best = v < g.Nloc ? v : g.Nloc; }
	g.subs_per_blk = (g.Nloc + best - 1) / best;              // individuals per CTA
	g.nblk = (g.Nloc + g.subs_per_blk - 1) / g.subs_per_blk;
	g.R = 1;
	g.zq_smem = g.KP == 8 ? (tlc == 1024 ? SnpLayout<8, 1024>::TOTAL : SnpLayout<8, 256>::TOTAL)
	                      : (tlc == 1024 ? SnpLayout<4, 1024>::TOTAL : SnpLayout<4, 256>::TOTAL);
	return cudaSuccess;
}

size_t snp_pc_floats(const Geometry &g) { return (size_t)g.nchunks * 2 * (size_t)g.KP * 2 * g.TL; }

// --------------------------------------------------------------------------------------
// The class-sorted store.  One warp per (chunk, individual): two passes over the individual's TLC genotypes in the
// micro-tiled store Xt (monomorphic and padded loci are already marked missing there, mcmc.c:817): count, then place.
// Position p of the list is stored at (p & ~127) + 4 (p & 31) + ((p >> 5) & 3), i.e. lane-major inside a block of 128.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t snp_index(uint32_t p) { return (p & ~127u) + 4u * (p & 31u) + ((p >> 5) & 3u); }

__global__ void snp_tile_kernel(const int16_t *Xt, uint32_t *Es, uint32_t *Hs, Geometry g)
{
	const int tlc = g.TL;
	const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (w >= (long)g.nchunks * g.Nloc) return;
	const int chunk = (int)(w / g.Nloc), il = (int)(w % g.Nloc);
	const int *X32 = reinterpret_cast<const int *>(Xt);
	const int ns = tlc / 32;                          // loci of this lane: l = 32 s + lane, s < ns <= 32
	// Within a class the order is ROUND-ROBIN BY RESIDUE l mod 32, not locus order: lane r owns the loci of residue r, round
	// n of a class holds the n-th such locus of every lane that has one.  While all 32 lanes still have one, a row of the
	// sweep is exactly one round and its lanes sit on 32 different shared-memory banks (the 16-byte row slots of a quarter
	// warp on 8 different bank groups); only the thinning rounds at the end of a class can collide.  In locus order the
	// gaps left by the other classes made nearly every row collide (first capture: 44 % of the shared wavefronts).
	uint32_t mh = 0u, mt = 0u;                        // bit s: locus 32 s + lane is a homozygote / a heterozygote
	for (int s = 0; s < ns; s++) {
		const int l = chunk * tlc + 32 * s + lane;
		int v = -1;
		if (l < g.Lpad) v = X32[((size_t)(l / TILE) * g.Nloc + il) * TILE + (l % TILE)];
		if (v >= 0) { if ((v & 0xFFFF) == ((v >> 16) & 0xFFFF)) mh |= 1u << s; else mt |= 1u << s; }
	}
	const uint32_t mm = ~(mh | mt) & (ns == 32 ? 0xFFFFFFFFu : ((1u << ns) - 1u));
	const int ch = __popc(mh), ct = __popc(mt), cm = __popc(mm);
	const int nhom = (int)__reduce_add_sync(0xffffffffu, (uint32_t)ch), nhet = (int)__reduce_add_sync(0xffffffffu, (uint32_t)ct);
	const uint32_t lt = (1u << lane) - 1u;
	uint32_t *out = Es + ((size_t)chunk * g.Nloc + il) * tlc;
	int base = 0;
	for (int cls = 0; cls < 3; cls++) {
		const uint32_t mask = cls == 0 ? mh : (cls == 1 ? mt : mm);
		const int cnt = cls == 0 ? ch : (cls == 1 ? ct : cm);
		for (int r = 0; r < ns; r++) {
			const uint32_t b = __ballot_sync(0xffffffffu, r < cnt);
			if (b == 0u) break;
			if (r < cnt) {
				const int s = (int)__fns(mask, 0u, r + 1);          // the (r+1)-th locus of this class in this lane
				const int ll = 32 * s + lane;
				const int l = chunk * tlc + ll;
				uint32_t e;
				if (cls == 2) e = 0x80000000u | (uint32_t)ll;
				else {
					const int v = X32[((size_t)(l / TILE) * g.Nloc + il) * TILE + (l % TILE)];
					e = (uint32_t)((v & 0xFFFF) * tlc + ll) | ((uint32_t)(((v >> 16) & 0xFFFF) * tlc + ll) << 16);
				}
				out[snp_index((uint32_t)(base + __popc(b & lt)))] = e;
			}
			base += __popc(b);
		}
	}
	if (lane == 0) Hs[(size_t)chunk * g.Nloc + il] = (uint32_t)nhom | ((uint32_t)nhet << 16);
}
cudaError_t launch_snp_tile(const int16_t *Xt, uint32_t *Es, uint32_t *Hs, Geometry g, cudaStream_t s)
{
	const long threads = (long)g.nchunks * g.Nloc * 32;
	snp_tile_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(Xt, Es, Hs, g);
	return cudaGetLastError();
}

// Z between the sorted store and the micro-tiled canonical one (state hooks only)
__global__ void snp_z_convert_kernel(const uint32_t *Es, uint16_t *Zs, uint16_t *Zt16, Geometry g, int to_sorted)
{
	const int tlc = g.TL;
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t total = (size_t)g.nchunks * g.Nloc * tlc;
	if (t >= total) return;
	const int il = (int)((t / tlc) % g.Nloc);
	const int chunk = (int)(t / ((size_t)tlc * g.Nloc));
	const uint32_t e = Es[t];
	const int ll = (e >> 31) ? (int)(e & 0xFFFFu) : (int)((e & 0xFFFFu) & (uint32_t)(tlc - 1));
	const int l = chunk * tlc + ll;
	if (l >= g.Lpad) { if (to_sorted) Zs[t] = 0; return; }
	const size_t src = ((size_t)(l / TILE) * g.Nloc + il) * TILE + (l % TILE);
	if (to_sorted) Zs[t] = Zt16[src];
	else Zt16[src] = Zs[t];
}
cudaError_t launch_snp_z_convert(const uint32_t *Es, uint16_t *Zs, int8_t *Zt, Geometry g, int to_sorted, cudaStream_t s)
{
	const size_t total = (size_t)g.nchunks * g.Nloc * g.TL;
	snp_z_convert_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(Es, Zs, reinterpret_cast<uint16_t *>(Zt), g, to_sorted);
	return cudaGetLastError();
}

// chunk-ordered P from the canonical P[l][x][k] (state injection and the initial pass; p_dirichlet writes both itself)
__global__ void snp_pc_from_p_kernel(const float *P, float *Pc, Geometry g)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (size_t)g.Lpad * 2 * g.KP) return;
	const int k = (int)(t % g.KP), x = (int)((t / g.KP) & 1), l = (int)(t / (2 * g.KP));
	snp_pc_store(Pc, g.TL, g.KP, l, x, k, P[t]);
}
cudaError_t launch_snp_pc_from_p(const float *P, float *Pc, Geometry g, cudaStream_t s)
{
	const size_t total = (size_t)g.Lpad * 2 * g.KP;
	snp_pc_from_p_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(P, Pc, g);
	return cudaGetLastError();
}

}  // namespace ig
