// zq_sweep.cu -- the one big pass of the diploid sweep (sm_100a).
//
//   new Z | P, Q            update_ZQ, mcmc.c:1133-1174 (disc_unif, random.c:403-430)
//   n[l][a][k] += ...       the tally half of the NEXT update_P, mcmc.c:810-845
//   cnt[i][k]               ancestry counts for the Q draw, mcmc.c:1176-1194
//   log-likelihood pieces   log_ld_indv (mcmc.c:1726-1773) on the OLD Z for update_G's accept
//                           (mcmc.c:1062-1089) and on the NEW Z for cal_lkh (mcmc.c:1916-1942)
//
// update_G reads the OLD Z and update_ZQ never reads G, so both ride one pass over the
// genotype store: the pass reads x (int16) and old z (int8) and writes new z (int8) -- the
// 4 bytes per allele copy of SURVEY.md section 8d.
//
// The kernel is bound by instruction issue, not by HBM, so every choice
// below is about warp instructions per genotype.  Measured pipe rates on B200
// (tools/ubench/pipes.cu): FFMA/FADD/FMUL/FFMA.SAT issue at 1 per clock per sub-partition,
// ALU-pipe ops (LOP3, SHF, IADD3, LEA, ISETP, FSEL) and IMAD at 1 per 2 clocks, IMAD.WIDE and
// IMAD.HI at 1 per 4, F2I at 1 per 8+.  Hence:
//   * the categorical search is a sum of saturating FFMAs (FMA pipe), its result becomes an
//     integer by a multiplication with a denormal (2^-147 * z has the bit pattern 4z), not F2I;
//   * the log-likelihood pieces are sums of MUFU.LG2 (the otherwise idle XU pipe) instead of
//     mantissa/exponent product accumulators (IMAD.HI + LOP3 + FMUL per factor);
//   * the new z pair is packed by denormal FFMAs, shared-memory addresses come from one LEA /
//     IADD each, per-genotype missing-data branches exist only in micro-tiles that contain
//     a missing genotype somewhere in the warp.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
// the K <= 8 categorical search as three compare/select levels: measured at config 4 together with the 16-bit draw,
// 3.99 -> 3.95 ms per launch (alone, on the 23-bit draw, it lost 1.5 %: profiles/r2_zq_levers.md)
#define IG_PICK_BSEARCH 1
#include "ig_internal.h"
#include "philox.cuh"
#include "sweep_common.cuh"

namespace ig {

#define LN2_D 0.69314718055994530942

// per-thread running state of one (chunk, individual)
struct Acc {
	float lgA, lgB, lgD;       // chunk totals of log2: new Z under g, new Z under g', old Z (g' - g)
	float mA, mB, mD;          // the same within the current micro-tile (two-level fp32 summation)
	int nsh_new;               // same-z heterozygotes on the new Z (the 2^-(g-1) factor of genofreq, mcmc.c:1692-1699)
};

struct Thr {
	float hist_bias;           // shared address of this lane's tally replica column minus R * psm, as a signed denormal
	float cnt_t;               // shared address of this thread's ancestry-counter column, as a denormal
	float omh_g, h_g, omh_p, h_p;
	float ftabf;               // FM == 2: shared address of the per-population table {h, 1-h, h', 1-h'}[KP], as a denormal
	float dltabf;              // FM == 2: shared address of lg2(1-h'_k) - lg2(1-h_k), as a denormal
	float dcolf, ecolf;        // FM == 2: shared addresses of this thread's per-population accumulator columns
};

__device__ __forceinline__ void sm_fadd(uint32_t addr, float v)      // this thread's private column: plain read-modify-write
{
	float o;
	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o) : "r"(addr) : "memory");
	o += v;
	asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(o) : "memory");
}

// One genotype.  xw = packed allele pair (x0 | x1 << 16), zw = word holding the old pair
// (z0 | z1 << 8) in its half H, r0/r1 = 32 random bits per copy, rowbf = shared address of
// P[l][0][0] as a denormal.  Returns the new packed pair (z0 | z1 << 8).
//
// FM = 2 (mode 4, mcmc_POP_inbreedcoff): the homozygosity excess h is the inbreeding coefficient
// of the population a same-z genotype sits in (log_ld_F_pop, mcmc.c:1776; genofreq_inbreedcoff
// :1707 is genofreq with h := F), so h comes from a table indexed by z, and the old-Z / new-Z
// differences between the proposed and the current coefficients are kept per population
// (all K Metropolis steps of update_inbreedcoff_POP, mcmc.c:986, ride this one pass: a
// genotype's term depends on one F_k only).
template <int KP, bool TF0, int LR, int H, int FM>
__device__ __forceinline__ uint32_t genotype(int xw, uint32_t zw, uint32_t r0, uint32_t r1, float rowbf, const float (&q)[KP],
                                             const Thr &t, Acc &acc, const RegConst &kc)
{
	const uint32_t x0 = __byte_perm((uint32_t)xw, 0u, 0x4410), x1 = __byte_perm((uint32_t)xw, 0u, 0x4432);
	const float row0f = fmaf(as_dn(x0), (float)(KP * 4), rowbf), row1f = fmaf(as_dn(x1), (float)(KP * 4), rowbf);
	const uint32_t row0 = __float_as_uint(row0f), row1 = __float_as_uint(row1f);
#if defined(IG_ABL_HOM)
	const bool het = false;
#elif defined(IG_ABL_HET)
	const bool het = true;
#else
	const bool het = (x0 != x1);
#endif
	// ---- cumulative weights w_k = sum_{m<=k} Q_im P_m,l,x (mcmc.c:1141-1149)
	float p0[KP], p1[KP], c0[KP], c1[KP];
#ifdef IG_ABL_HOM
	// ABLATION (timing only, wrong chain): every genotype treated as a homozygote -- one row, one prefix
#pragma unroll
	for (int v = 0; v < KP / 4; v++) {
		const float4 t0 = lds_f4(row0 + 16 * v);
		p0[4 * v] = t0.x; p0[4 * v + 1] = t0.y; p0[4 * v + 2] = t0.z; p0[4 * v + 3] = t0.w;
	}
	c0[0] = q[0] * p0[0];
#pragma unroll
	for (int k = 1; k < KP; k++) c0[k] = fmaf(q[k], p0[k], c0[k - 1]);
#pragma unroll
	for (int k = 0; k < KP; k++) { c1[k] = c0[k]; p1[k] = p0[k]; }
#else
#pragma unroll
	for (int v = 0; v < KP / 4; v++) {
		const float4 t0 = lds_f4(row0 + 16 * v), t1 = lds_f4(row1 + 16 * v);
		p0[4 * v] = t0.x; p0[4 * v + 1] = t0.y; p0[4 * v + 2] = t0.z; p0[4 * v + 3] = t0.w;
		p1[4 * v] = t1.x; p1[4 * v + 1] = t1.y; p1[4 * v + 2] = t1.z; p1[4 * v + 3] = t1.w;
	}
	c0[0] = q[0] * p0[0];
	c1[0] = q[0] * p1[0];
#pragma unroll
	for (int k = 1; k < KP; k++) { c0[k] = fmaf(q[k], p0[k], c0[k - 1]); c1[k] = fmaf(q[k], p1[k], c1[k - 1]); }
#endif
	// ---- old-Z piece of update_G's ratio (log_ld_indv, mcmc.c:1752-1759): only same-z
	//      homozygotes depend on g.  (The 2^-(g-1) count of same-z heterozygotes on the old
	//      Z is the previous pass's nsh_new; indiv_epilogue carries it over.)
#if defined(IG_ABL_NOD) || defined(IG_ABL_HET)
	if (false) {
#else
	if (!TF0 && FM != 2) {
#endif
		const uint32_t zo0 = __byte_perm(zw, 0u, H ? 0x4442 : 0x4440), zo1 = __byte_perm(zw, 0u, H ? 0x4443 : 0x4441);
		const float fo = lds_f(__float_as_uint(fmaf(as_dn(zo0), 4.0f, row0f)));
		const float fe = (zo0 == zo1 && !het) ? fo : 1.0f;                // h + 1 * (1 - h) == 1 exactly
		acc.mD += lg2_fast(fmaf(fe, t.omh_p, t.h_p)) - lg2_fast(fmaf(fe, t.omh_g, t.h_g));
	}
	if (FM == 2) {
		const uint32_t zo0 = __byte_perm(zw, 0u, H ? 0x4442 : 0x4440), zo1 = __byte_perm(zw, 0u, H ? 0x4443 : 0x4441);
		if (zo0 == zo1) {
			const float fo = lds_f(__float_as_uint(fmaf(as_dn(zo0), 4.0f, row0f)));
			const float4 ft = lds_f4(__float_as_uint(fmaf(as_dn(zo0), 16.0f, t.ftabf)));
			const float dl = lds_f(__float_as_uint(fmaf(as_dn(zo0), 4.0f, t.dltabf)));
			const float term = het ? dl : lg2_fast(fmaf(fo, ft.w, ft.z)) - lg2_fast(fmaf(fo, ft.y, ft.x));
			sm_fadd(__float_as_uint(fmaf(as_dn(zo0), (float)(4 * ZQ_THREADS), t.dcolf)), term);
		}
	}
	// ---- categorical draws (disc_unif, random.c:403-430)
	const float zf0 = pick_category<KP>(c0, uniform_big16(r0, kc));
	const float zf1 = pick_category<KP>(c1, uniform_big16(r1, kc));
	const float pa0f = fmaf(zf0, as_dn(4u), row0f), pa1f = fmaf(zf1, as_dn(4u), row1f);           // &P[l][x][z]
	// ---- n[l][a][k] tally for the next update_P (mcmc.c:815-845) and the individual's
	//      ancestry counts (mcmc.c:1176-1194): shared-memory RED
#ifndef IG_ABL_NOTALLY
	red_inc(__float_as_uint(fmaf(pa0f, (float)(1 << LR), t.hist_bias)));
	red_inc(__float_as_uint(fmaf(pa1f, (float)(1 << LR), t.hist_bias)));
#endif
#ifndef IG_ABL_NOCNT
	red_inc(__float_as_uint(fmaf(zf0, as_dn(4u * ZQ_THREADS), t.cnt_t)));
	red_inc(__float_as_uint(fmaf(zf1, as_dn(4u * ZQ_THREADS), t.cnt_t)));
#endif
	// ---- new-Z likelihood pieces (cal_lkh and the accepted-G selection)
	float f0, f1;
	bool same_n;
	if (TF0) { f0 = c0[KP - 1]; f1 = c1[KP - 1]; same_n = true; }                                 // mcmc.c:1739-1749
#if defined(IG_ABL_NOLIK)
	else { f0 = f1 = 1.0f; same_n = (zf0 == zf1); }
#else
	else { f0 = lds_f(__float_as_uint(pa0f)); f1 = lds_f(__float_as_uint(pa1f)); same_n = (zf0 == zf1); }
#endif
	const bool sh_n = same_n && !het;
	if (FM == 2) {
		float fac = f1;
		if (same_n) {
			const float4 fn = lds_f4(__float_as_uint(fmaf(zf0, as_dn(16u), t.ftabf)));
			const float dl = lds_f(__float_as_uint(fmaf(zf0, as_dn(4u), t.dltabf)));
			const float cur = fmaf(f0, fn.y, fn.x);
			fac = het ? f1 * fn.y : cur;                                                      // 2 f0 f1 (1 - F) : f0 (F + f0 (1 - F))
			const float term = het ? dl : lg2_fast(fmaf(f0, fn.w, fn.z)) - lg2_fast(cur);
			sm_fadd(__float_as_uint(fmaf(zf0, as_dn(4u * ZQ_THREADS), t.ecolf)), term);
		}
		acc.mA += lg2_fast(f0 * fac);
	} else {
#if defined(IG_ABL_NOLIK)
		acc.nsh_new += (same_n && het) ? 1 : 0;
#elif defined(IG_ABL_HET)
		acc.mA += lg2_fast(f0 * f1);
		acc.nsh_new += same_n ? 1 : 0;
#else
		acc.mA += lg2_fast(f0 * (sh_n ? fmaf(f0, t.omh_g, t.h_g) : f1));
		acc.mB += lg2_fast(f0 * (sh_n ? fmaf(f0, t.omh_p, t.h_p) : f1));
		acc.nsh_new += (same_n && het) ? 1 : 0;
#endif
	}
	return __float_as_uint(fmaf(zf1, as_dn(256u), zf0 * as_dn(1u)));                              // z0 | z1 << 8
}

// One micro-tile of 8 loci.  CHECK = false: no thread of the warp holds a missing genotype
// here, so the per-genotype sign test and its divergence bookkeeping are compiled out.
template <int KP, int ROUNDS, bool TF0, int LR, bool CHECK, int FM>
__device__ __forceinline__ void micro_tile(const int (&xw)[8], const uint32_t (&zwo)[4], uint32_t (&zwn)[4], float rowb0f, float rowstridef,
                                           uint32_t mt_global, uint32_t ig_global, uint32_t iter, const ZQArgs &a, const float (&q)[KP], const Thr &t,
                                           Acc &acc, const RegConst &kc)
{
	// 16 random bits per allele copy: one Philox block serves FOUR genotypes.  Copy 0 of a genotype takes bits 22..7 of
	// its word as they lie (the mantissa of uniform_big), copy 1 the other sixteen, rotated into the same place by one PRMT.
#pragma unroll
	for (int ph = 0; ph < 2; ++ph) {
		const u32x4 rnd = philox4x32<ROUNDS>(u32x4{mt_global, ig_global, iter, TAG_Z16 | (uint32_t)ph}, a.key0, a.key1);
		const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
		for (int qq = 0; qq < 2; ++qq) {
			const int pr = 2 * ph + qq;
			uint32_t pair[2];
#pragma unroll
			for (int h2 = 0; h2 < 2; ++h2) {
				const int j = 2 * pr + h2;
				const uint32_t ra = rr[2 * qq + h2], rb = __byte_perm(ra, 0u, 0x1032);
				const float rowbf = fmaf((float)j, rowstridef, rowb0f);
				if (CHECK) {
					pair[h2] = h2 ? (zwo[pr] >> 16) : (zwo[pr] & 0xFFFFu);
					if (xw[j] >= 0) {
						if (h2) pair[h2] = genotype<KP, TF0, LR, 1, FM>(xw[j], zwo[pr], ra, rb, rowbf, q, t, acc, kc);
						else pair[h2] = genotype<KP, TF0, LR, 0, FM>(xw[j], zwo[pr], ra, rb, rowbf, q, t, acc, kc);
					}
				} else {
					if (h2) pair[h2] = genotype<KP, TF0, LR, 1, FM>(xw[j], zwo[pr], ra, rb, rowbf, q, t, acc, kc);
					else pair[h2] = genotype<KP, TF0, LR, 0, FM>(xw[j], zwo[pr], ra, rb, rowbf, q, t, acc, kc);
				}
			}
			zwn[pr] = __byte_perm(pair[0], pair[1], 0x5410);
		}
	}
}

// --------------------------------------------------------------------------------------
// zq_sweep: grid (locus chunks, individual blocks), 256 threads, one thread = one
// individual, marching over the chunk's micro-tiles of 8 loci.
//   global  : X  int16 [LT][Nloc][8][2]  two 128-bit loads per thread per micro-tile
//             Z  int8  [LT][Nloc][8][2]  one 128-bit load + one 128-bit store
//   shared  : P chunk [TL][A][KP] fp32, landed by ONE TMA bulk copy on an mbarrier;
//             n chunk [TL][A][KP][R] int32 histogram, R lane-replicas to thin out conflicts,
//             reduced and pushed to global n with RED at the end of the CTA;
//             per-thread ancestry counters [KP][256] int32 (column tid: conflict-free RED)
//   output  : per (chunk, individual) partials: K counts (u16) + 3 log-likelihood pieces
// --------------------------------------------------------------------------------------
template <int KP, int ROUNDS, bool TF0, int LR, int FM>
__global__ void __launch_bounds__(ZQ_THREADS, ZQ_MIN_CTAS) zq_sweep_kernel(const ZQArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ __align__(8) unsigned long long bar;
	constexpr int R = 1 << LR;
	const Geometry &g = a.geo;
	const int tid = threadIdx.x;
	const int chunk = blockIdx.x;
	const int l0 = chunk * g.TL;
	const int nl = min(g.TL, g.Lpad - l0);
	const int nmt = nl / TILE;
	const int rowsz = g.A * KP;                     // floats per locus
	float *Psm = reinterpret_cast<float *>(smem_raw);
	int *hist = reinterpret_cast<int *>(Psm + (size_t)g.TL * rowsz);
	int *cntsm = hist + (size_t)g.TL * rowsz * R;                   // [KP][ZQ_THREADS]
	float *dsm = reinterpret_cast<float *>(cntsm + KP * ZQ_THREADS); // FM == 2: [KP][ZQ_THREADS] old-Z differences per population
	float *esm = dsm + KP * ZQ_THREADS;                             //          [KP][ZQ_THREADS] new-Z differences
	float *ftab = esm + KP * ZQ_THREADS;                            //          [KP][4] + [KP]
	const int nbins = nl * rowsz;

	if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
	__syncthreads();
	if (tid == 0) {
		mbar_expect_tx(&bar, (uint32_t)nbins * 4u);
		tma_bulk_g2s(Psm, a.P + (size_t)l0 * rowsz, (uint32_t)nbins * 4u, &bar);
	}
	for (int j = tid; j < nbins * R; j += ZQ_THREADS) hist[j] = 0;
#pragma unroll
	for (int k = 0; k < KP; k++) cntsm[k * ZQ_THREADS + tid] = 0;
	if (FM == 2) {
#pragma unroll
		for (int k = 0; k < KP; k++) { dsm[k * ZQ_THREADS + tid] = 0.0f; esm[k * ZQ_THREADS + tid] = 0.0f; }
		if (tid < KP * 5) ftab[tid] = a.ftab[tid];
	}
	__syncthreads();
	mbar_wait(&bar, 0);

	const int Nloc = g.Nloc;
	const int mt0 = l0 / TILE;
	const int sub0 = blockIdx.y * g.subs_per_blk;
	const int nsub_total = (Nloc + ZQ_THREADS - 1) / ZQ_THREADS;
	const int sub1 = min(sub0 + g.subs_per_blk, nsub_total);
	const uint32_t psm = smem_addr(Psm);
	const float rowstridef = as_dn((uint32_t)rowsz * 4u);           // bytes per locus in the P chunk
	const float tilestridef = as_dn((uint32_t)rowsz * 4u * TILE);
	const RegConst kc{a.k_mant, a.k_one};
	const uint32_t iter = a.iter_dev ? *a.iter_dev : a.iter;
	Thr t;
	t.hist_bias = as_dn_signed((int)smem_addr(hist) + (tid & (R - 1)) * 4 - (int)(psm << LR));
	t.cnt_t = as_dn(smem_addr(cntsm) + (uint32_t)tid * 4u);
	t.ftabf = t.dltabf = t.dcolf = t.ecolf = 0.0f;
	if (FM == 2) {
		t.ftabf = as_dn(smem_addr(ftab));
		t.dltabf = as_dn(smem_addr(ftab + KP * 4));
		t.dcolf = as_dn(smem_addr(dsm) + (uint32_t)tid * 4u);
		t.ecolf = as_dn(smem_addr(esm) + (uint32_t)tid * 4u);
	}

	for (int sub = sub0; sub < sub1; ++sub) {
		// Warps vote inside the loop (__any_sync), so a warp stays together: a warp with no
		// individual at all leaves as a whole, and in the last, partly filled warp the surplus
		// lanes shadow individual Nloc-1 with every genotype marked missing and nothing stored.
		if (sub * ZQ_THREADS + (tid & ~31) >= Nloc) continue;
		const bool live = sub * ZQ_THREADS + tid < Nloc;
		const int il = live ? sub * ZQ_THREADS + tid : Nloc - 1;
		{
			float q[KP];
			{
				const float4 *qp = reinterpret_cast<const float4 *>(a.Qf + (size_t)il * KP);
#pragma unroll
				for (int v = 0; v < KP / 4; v++) {
					float4 w = __ldg(qp + v);
					q[4 * v] = w.x; q[4 * v + 1] = w.y; q[4 * v + 2] = w.z; q[4 * v + 3] = w.w;
				}
			}
			const int2 gg = __ldg(a.gpair + il);
			// 1 - h(g) = 2^-(g-1); exact in fp32 down to 2^-126, 0 beyond (g can start huge in mode 3)
			t.omh_g = (gg.x <= 127) ? __int_as_float((128 - gg.x) << 23) : 0.0f;
			t.omh_p = (gg.y <= 127) ? __int_as_float((128 - gg.y) << 23) : 0.0f;
			// mode 5 (mcmc_INDV_inbreedcoff): h is the individual's inbreeding coefficient, current and
			// proposed (genofreq_inbreedcoff, mcmc.c:1707, is genofreq with h := F); hpair holds 1 - F
			if (FM == 0 && a.hpair) { const float2 hp = __ldg(a.hpair + il); t.omh_g = hp.x; t.omh_p = hp.y; }
			t.h_g = 1.0f - t.omh_g;                                         // omh + (1 - omh) == 1 exactly in fp32
			t.h_p = 1.0f - t.omh_p;
			Acc acc;
			acc.lgA = acc.lgB = acc.lgD = 0.0f;
			acc.nsh_new = 0;

			const int4 *xp = reinterpret_cast<const int4 *>(a.Xt) + ((size_t)mt0 * Nloc + il) * 2;
			int4 *zp = reinterpret_cast<int4 *>(a.Zt) + ((size_t)mt0 * Nloc + il);
			const size_t xstride = (size_t)Nloc * 2, zstride = (size_t)Nloc;
			const uint32_t ig_global = (uint32_t)(g.i0 + il);
			float rowb0f = as_dn(psm);

			// Software prefetch, one micro-tile ahead: without it a fifth of all warp time was
			// spent waiting on these loads at the first use of x (ncu r1c: long_scoreboard on one
			// instruction), because four warps per scheduler do not cover ~1 us of HBM latency.
			int4 xa = ldg_stream(xp), xb = ldg_stream(xp + 1);
			int4 zz = ldg_rw(zp);
			for (int mt = 0; mt < nmt; ++mt, rowb0f += tilestridef) {
				const int4 cxa = xa, cxb = xb, czz = zz;
				if (mt + 1 < nmt) {
					xa = ldg_stream(xp + (size_t)(mt + 1) * xstride);
					xb = ldg_stream(xp + (size_t)(mt + 1) * xstride + 1);
					zz = ldg_rw(zp + (size_t)(mt + 1) * zstride);
				}
				const int dead = live ? 0 : -1;
				const int xw[8] = {cxa.x | dead, cxa.y | dead, cxa.z | dead, cxa.w | dead, cxb.x | dead, cxb.y | dead, cxb.z | dead, cxb.w | dead};
				const uint32_t zwo[4] = {(uint32_t)czz.x, (uint32_t)czz.y, (uint32_t)czz.z, (uint32_t)czz.w};
				uint32_t zwn[4];
				acc.mA = acc.mB = acc.mD = 0.0f;
				// sign of the AND of all eight words: set iff every genotype is usable
				const int any_missing = (xw[0] | xw[1] | xw[2] | xw[3] | xw[4] | xw[5] | xw[6] | xw[7]) < 0;
				if (__any_sync(0xffffffffu, any_missing))
					micro_tile<KP, ROUNDS, TF0, LR, true, FM>(xw, zwo, zwn, rowb0f, rowstridef, (uint32_t)(mt0 + mt), ig_global, iter, a, q, t, acc, kc);
				else
					micro_tile<KP, ROUNDS, TF0, LR, false, FM>(xw, zwo, zwn, rowb0f, rowstridef, (uint32_t)(mt0 + mt), ig_global, iter, a, q, t, acc, kc);
				if (live) stg_stream(zp + (size_t)mt * zstride, make_int4((int)zwn[0], (int)zwn[1], (int)zwn[2], (int)zwn[3]));
				acc.lgA += acc.mA; acc.lgB += acc.mB; acc.lgD += acc.mD;
			}
			// ---- partials of this (chunk, individual)
			if (live) {
				uint32_t *pc = reinterpret_cast<uint32_t *>(a.pcnt + ((size_t)chunk * Nloc + il) * KP);
				int *cnt_col = cntsm + tid;
#pragma unroll
				for (int j = 0; j < KP / 2; j++) {
					const int ca = cnt_col[(2 * j) * ZQ_THREADS], cb = cnt_col[(2 * j + 1) * ZQ_THREADS];
					cnt_col[(2 * j) * ZQ_THREADS] = 0;
					cnt_col[(2 * j + 1) * ZQ_THREADS] = 0;
					pc[j] = (uint32_t)ca | ((uint32_t)cb << 16);
				}
				double *pl = a.plog + (size_t)chunk * 3 * Nloc + il;
				const double la = (double)acc.lgA * LN2_D, lb = (double)acc.lgB * LN2_D;
				if (FM == 2) {
					pl[0] = 0.0; pl[(size_t)Nloc] = la; pl[(size_t)2 * Nloc] = la;
					float *pd = a.pfk + ((size_t)chunk * Nloc + il) * 2 * KP;
#pragma unroll
					for (int k = 0; k < KP; k++) {
						pd[k] = dsm[k * ZQ_THREADS + tid]; pd[KP + k] = esm[k * ZQ_THREADS + tid];
						dsm[k * ZQ_THREADS + tid] = 0.0f; esm[k * ZQ_THREADS + tid] = 0.0f;
					}
				} else if (a.hpair) {
					// same-z heterozygotes carry (1 - F) (genofreq_inbreedcoff, mcmc.c:1719); 0 * log 0 is kept at 0
					const double wg = acc.nsh_new ? (double)acc.nsh_new * log((double)t.omh_g) : 0.0;
					const double wp = acc.nsh_new ? (double)acc.nsh_new * log((double)t.omh_p) : 0.0;
					pl[0] = (double)acc.lgD * LN2_D; pl[(size_t)Nloc] = la + wg; pl[(size_t)2 * Nloc] = lb + wp;
				} else {
					// TF0: the likelihood does not depend on Z, so the old-Z ratio is the new-Z one
					pl[0] = TF0 ? (lb - la) - (double)acc.nsh_new * (double)(gg.y - gg.x) * LN2_D : (double)acc.lgD * LN2_D;
					pl[(size_t)Nloc] = la - (double)acc.nsh_new * (double)(gg.x - 1) * LN2_D;
					pl[(size_t)2 * Nloc] = lb - (double)acc.nsh_new * (double)(gg.y - 1) * LN2_D;
				}
				a.pnsh[(size_t)chunk * Nloc + il] = (uint16_t)acc.nsh_new;
			}
		}
	}
	__syncthreads();
	// ---- reduce the replicas and push this CTA's tally into global n (RED, no return value)
	int32_t *ng = a.n + (size_t)l0 * rowsz;
	for (int b = tid; b < nbins; b += ZQ_THREADS) {
		int s = 0;
#pragma unroll
		for (int r = 0; r < R; r++) s += hist[b * R + r];
		if (s) atomicAdd(ng + b, s);
	}
}

template <int KP, int LR>
static cudaError_t launch_zq_kp(const ZQArgs &a, int rounds, cudaStream_t s)
{
	dim3 grid(a.geo.nchunks, a.geo.nblk), block(ZQ_THREADS);
	const size_t sm = a.geo.zq_smem;
#define IG_LAUNCH(RND, TF, FM)                                                                               \
	do {                                                                                                 \
		cudaError_t e = cudaFuncSetAttribute(zq_sweep_kernel<KP, RND, TF, LR, FM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
		if (e != cudaSuccess) return e;                                                                  \
		zq_sweep_kernel<KP, RND, TF, LR, FM><<<grid, block, sm, s>>>(a);                                  \
	} while (0)
	if (a.fmode == 2) { if (rounds == 7) IG_LAUNCH(7, false, 2); else IG_LAUNCH(10, false, 2); }
	else if (a.type_freq == 0) { if (rounds == 7) IG_LAUNCH(7, true, 0); else IG_LAUNCH(10, true, 0); }
	else { if (rounds == 7) IG_LAUNCH(7, false, 0); else IG_LAUNCH(10, false, 0); }
#undef IG_LAUNCH
	return cudaGetLastError();
}

cudaError_t launch_zq_sweep(const ZQArgs &a, int rounds, cudaStream_t s)
{
	const int key = a.geo.KP * 16 + a.geo.R;
	switch (key) {
#ifndef IG_FAST_BUILD
	case 4 * 16 + 8: return launch_zq_kp<4, 3>(a, rounds, s);
	case 4 * 16 + 1: return launch_zq_kp<4, 0>(a, rounds, s);
	case 8 * 16 + 1: return launch_zq_kp<8, 0>(a, rounds, s);
	case 16 * 16 + 8: return launch_zq_kp<16, 3>(a, rounds, s);
	case 16 * 16 + 1: return launch_zq_kp<16, 0>(a, rounds, s);
#endif
	case 8 * 16 + 8: return launch_zq_kp<8, 3>(a, rounds, s);
	default: return cudaErrorInvalidValue;
	}
}

// Choose the decomposition: chunks of TL loci x blocks of individuals.  The tally of a chunk
// is complete inside one CTA when nblk == 1 (no contention on global n); per-individual
// pieces are always combined by indiv_epilogue in chunk order (deterministic).
cudaError_t zq_configure(Geometry &g, int device)
{
	int sms = 148, smem_optin = 227 * 1024;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
	cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
	const int target_ctas = 2 * ZQ_MIN_CTAS * sms;         // ZQ_MIN_CTAS resident CTAs per SM, two waves
	// per-thread counter columns; mode 4 adds two float columns per population and the F table
	const size_t cnt_bytes = (size_t)g.KP * ZQ_THREADS * sizeof(int) * (g.fmode == 2 ? 3 : 1) + (g.fmode == 2 ? (size_t)g.KP * 5 * 4 + 16 : 0);
	const size_t budget = (size_t)min(smem_optin, 227 * 1024) / ZQ_MIN_CTAS - 2048 - cnt_bytes;
	const size_t per_locus = (size_t)g.A * g.KP * 4;
	const int nsub_total = (g.Nloc + ZQ_THREADS - 1) / ZQ_THREADS;
	// 8 lane-replicas of the tally histogram, or none when allelenum_max is so large that
	// eight do not fit (many alleles spread the lanes over many bins anyway)
	int R = 8;
	if (per_locus * (1 + R) * TILE > budget) R = 1;
	if (per_locus * (1 + R) * TILE > (size_t)smem_optin - 2048 - cnt_bytes) return cudaErrorInvalidConfiguration;
	int tl_max = (int)(budget / (per_locus * (1 + R)));
	if (tl_max < TILE) tl_max = (int)(((size_t)smem_optin - 2048 - cnt_bytes) / (per_locus * (1 + R)));
	tl_max = (tl_max / TILE) * TILE;
	if (tl_max < TILE) return cudaErrorInvalidConfiguration;
	if (tl_max > 1024) tl_max = 1024;
	int tl = ((g.Lpad + target_ctas - 1) / target_ctas + TILE - 1) / TILE * TILE;
	if (tl < TILE) tl = TILE;
	if (tl > tl_max) tl = tl_max;
	g.TL = tl;
	g.nchunks = (g.Lpad + tl - 1) / tl;
	// about 15 waves of CTAs: measured at config 4 (569 chunks, 40 passes of 256 individuals), 2 / 4 / 8 / 10 / 20
	// individual blocks give 4.42 / 4.38 / 4.26 / 4.30 / 4.31 ms per launch -- a coarse grid loses its last wave
	const int want_ctas = 15 * ZQ_MIN_CTAS * sms;
	int nblk = min(nsub_total, (want_ctas + g.nchunks - 1) / g.nchunks);
	// a CTA's prologue (P chunk, zeroed histogram) wants at least two passes behind it once the chunks alone fill the
	// GPU: a shard of 1250 individuals (5 passes) runs 0.591 ms per launch as 2 blocks, 0.603 ms as 5
	if (g.nchunks >= ZQ_MIN_CTAS * sms && nsub_total >= 2) nblk = min(nblk, nsub_total / 2);
	if (nblk < 1) nblk = 1;
	if (const char *e = getenv("IG_ZQ_NBLK")) nblk = max(1, min(nsub_total, atoi(e)));       // tuning hook
	g.subs_per_blk = (nsub_total + nblk - 1) / nblk;
	g.nblk = (nsub_total + g.subs_per_blk - 1) / g.subs_per_blk;
	g.R = R;
	g.zq_smem = (size_t)tl * per_locus * (1 + R) + cnt_bytes;
	return cudaSuccess;
}

}  // namespace ig
