// philox.cuh -- counter-based Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11), host + device.
//
// Replaces the reference's global-state Wichmann-Hill generator (random.c:10-47).  Every
// random number in a chain is a pure function of
//     key     = (seed, chain)
//     counter = (site index 0, site index 1, sweep iteration, purpose tag | draw number)
// so the stream does not depend on how individuals are sharded over GPUs, on grid shape or
// on launch order.  The site indices are (locus micro-tile, global individual) for the Z
// draw, (locus, population) for the P draw, (individual, 0) for Q / G, and so on.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define IG_HD __host__ __device__ __forceinline__
#else
#define IG_HD inline
#endif

namespace ig {

struct u32x4 { uint32_t x, y, z, w; };

constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

IG_HD void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
{
#if defined(__CUDA_ARCH__)
	uint64_t p;
	asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));     // one IMAD.WIDE.U32
	lo = (uint32_t)p;
	hi = (uint32_t)(p >> 32);
#else
	uint64_t p = (uint64_t)a * b;
	lo = (uint32_t)p;
	hi = (uint32_t)(p >> 32);
#endif
}

template <int ROUNDS = 10>
IG_HD u32x4 philox4x32(u32x4 c, uint32_t k0, uint32_t k1)
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int r = 0; r < ROUNDS; r++) {
		uint32_t h0, l0, h1, l1;
		mulhilo(PHILOX_M0, c.x, h0, l0);
		mulhilo(PHILOX_M1, c.z, h1, l1);
		c = u32x4{h1 ^ c.y ^ k0, l1, h0 ^ c.w ^ k1, l0};
		k0 += PHILOX_W0;
		k1 += PHILOX_W1;
	}
	return c;
}

// purpose tags (upper bits of the 4th counter word; the low 16 bits count draws at a site)
enum : uint32_t {
	TAG_Z = 1u << 16,       // per-allele-copy ancestry draw      (update_ZQ, mcmc.c:1139-1153)
	TAG_ZINIT = 2u << 16,   // uniform initial assignment          (init_flag = 1)
	TAG_P = 3u << 16,       // allele-frequency Dirichlet          (update_P, mcmc.c:846-857)
	TAG_Q = 4u << 16,       // admixture Dirichlet                 (mcmc.c:1196-1198)
	TAG_GPROP = 5u << 16,   // generation proposal                 (update_G, mcmc.c:1074)
	TAG_GACC = 6u << 16,    // generation accept                   (mcmc.c:1086)
	TAG_SPOP = 7u << 16,    // population selfing-rate MH          (update_S_POP, mcmc.c:913)
	TAG_SIND = 8u << 16,    // individual selfing-rate MH          (update_S_IND, mcmc.c:864)
	TAG_ALPHA = 9u << 16,   // alpha MH                            (update_alpha, mcmc.c:1244)
	TAG_INIT = 10u << 16,   // chain initialisation                (mcmc.c:196-205,326-331,479)
	TAG_DP = 11u << 16,     // host-side Dirichlet-process step    (DPMM.c:124-199)
	TAG_GENO = 12u << 16,   // tetraploid dosage resolution        (poly_geno.c:520)
	TAG_TETRA = 13u << 16,  // tetraploid selfing-rate MH          (poly_geno.c:584)
	TAG_ZS = 15u << 16,     // ancestry draw of the biallelic path (zq_snp.cu): block (chunk step, individual), draw number = lane
	TAG_Z16 = 14u << 16,    // ancestry draw, 16 random bits per allele copy (one block per four genotypes)
};

// 24-bit uniform in [0,1) as float, 53-bit uniform in (0,1) as double
IG_HD float u01f(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
IG_HD double u01d(uint32_t hi, uint32_t lo)
{
	uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;            // 53 bits
	return ((double)v + 0.5) * (1.0 / 9007199254740992.0);     // never 0, never 1
}

// A small sequential stream on top of the counter-based generator, for the low-volume
// floating-point draws (gamma / normal / uniform): site (a, b, iter, tag) and a running
// draw counter in the low 16 bits of the 4th word.  64 K draws per site per sweep is far
// beyond what any rejection loop here uses.
struct Stream {
	uint32_t a, b, it, tag, k0, k1, n;
	u32x4 buf;
	int have;
	IG_HD Stream(uint32_t a_, uint32_t b_, uint32_t it_, uint32_t tag_, uint32_t k0_, uint32_t k1_)
	    : a(a_), b(b_), it(it_), tag(tag_), k0(k0_), k1(k1_), n(0), buf{0u, 0u, 0u, 0u}, have(0) {}
	IG_HD double uniform()
	{
		if (!have) { buf = philox4x32<10>(u32x4{a, b, it, tag | (n & 0xFFFFu)}, k0, k1); n++; have = 2; }
		double u = (have == 2) ? u01d(buf.x, buf.y) : u01d(buf.z, buf.w);
		have--;
		return u;
	}
};

}  // namespace ig
