// ig_internal.h -- context layout and kernel launch interfaces shared by ig_api.cu and the
// kernel translation units.  Not part of the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/instruct_b200.h"

namespace ig {

constexpr int TILE = 8;          // loci per micro-tile (one 128-bit Z vector, two 128-bit X vectors)
constexpr int ZQ_THREADS = 256;  // individuals per CTA pass
constexpr int ZQ_MIN_CTAS = 2;   // resident CTAs per SM the sweep kernel is compiled and sized for
constexpr int SNP_THREADS = 512; // zq_snp.cu: 16 warps, one CTA per SM
constexpr int MAX_K = 16;
constexpr int SC_MAX_CTAS = 592;   // cooperative scalar-update kernels: at most 4 CTAs per SM
constexpr float P_FLOOR = 1e-18f;   // keeps f0*f1 a normal fp32 number in the product accumulators

// device-resident scalars of one chain
struct DevScalars {
	double alpha;
	double totallkh;
	double sumlogq;          // sum_i sum_k log q_ik  (sufficient statistic of update_alpha)
	double cur_prop_ll;      // proposal() at the current S, kept between update_S_POP steps
	double qcol[MAX_K];      // column sums of Q (check_empty_cluster)
	int32_t alpha_accepts;
	int32_t s_accepts;
	int32_t flags;
	uint32_t iter;           // the NEXT sweep's counter (RNG key); read by the kernels of a graph-replayed sweep, advanced by post_sweep
};

// running moments (store_chn, mcmc.c:1320) on the device
struct Moments {
	double *tot;       // [2] totallkh, totallkh2
	double *indvlkh;   // [N]
	double *qq, *qq2;  // [N][K]
	double *self, *self2;
	double *gen, *gen2;
	double *freq, *freq2;   // [K][L][A] (print_freq)
	double *convg;     // [ckrep]
	double *convg_S;   // [ckrep][K] the population rates (selfing rates / inbreeding coefficients) of the same retained sweeps, or null
};

struct Geometry {
	int N, Nloc, i0;         // global individuals, local shard size and first global index
	int L, Lpad, LT;         // loci, padded to TILE, micro-tiles
	int K, KP, A;            // populations, padded populations, allelenum_max
	int REC;                 // doubles per individual record: Q[K], lkh, slq, G (mode 5: F); mode 4 appends D[K], E[K]
	int TL;                  // loci per chunk (multiple of TILE)
	int nchunks;             // ceil(Lpad / TL)
	int nblk;                // individual blocks (grid.y)
	int subs_per_blk;        // 256-individual passes per CTA
	int R;                   // smem histogram replicas
	int snp;                 // 1: the biallelic path (zq_snp.cu): TL = compile-time chunk length, subs_per_blk = individuals per CTA
	int fmode;               // 0: selfing generations (modes 1-3), 1: inbreeding coefficient per individual (mode 5), 2: per population (mode 4)
	size_t zq_smem;          // dynamic shared memory bytes of zq_sweep
	int tab_stage;           // ploid 4, PASS B: the chunk's genotype-frequency tables and the code -> index bytes are staged in shared memory
	int c2i_bytes;           // ploid 4: size of the code -> index byte array (all catalogues)
};

struct ZQArgs {
	const int16_t *Xt;       // [LT][Nloc][TILE][2]
	int8_t *Zt;              // [LT][Nloc][TILE][2]
	const float *P;          // [Lpad][A][KP]
	int32_t *n;              // [Lpad][A][KP]
	const float *Qf;         // [Nloc][KP]
 	const int2 *gpair;       // [Nloc] (g, g')
	uint16_t *pcnt;          // [nchunks][Nloc][KP]
	double *plog;            // [nchunks][3][Nloc]  old-Z ratio piece, new-Z likelihood under g and under g'
	uint16_t *pnsh;          // [nchunks][Nloc]     same-z heterozygotes on the new Z
	Geometry geo;
	uint32_t iter;
	uint32_t key0, key1;
	const uint32_t *iter_dev;   // non-null: read the sweep counter from device memory (CUDA-graph replay)
	const float2 *hpair;        // mode 5: [Nloc] (1 - F, 1 - F') in place of the generation pair
	const float *ftab;          // mode 4: {F, 1-F, F', 1-F'}[KP] then lg2(1-F') - lg2(1-F) [KP]
	float *pfk;                 // mode 4: [nchunks][Nloc][2][KP] per-population old-Z / new-Z differences (log2 units)
	int fmode;
	int type_freq;
	uint32_t k_mant, k_one;  // 0x007fffff, 0x3f800000 kept in registers on purpose (see uniform_big)
};

// ---- launchers (ig_kernels.cu) -------------------------------------------------------
cudaError_t launch_zq_sweep(const ZQArgs &a, int rounds, cudaStream_t s);
cudaError_t zq_configure(Geometry &g, int device);

// ---- the biallelic path (zq_snp.cu) ---------------------------------------------------
struct SnpArgs {
	const uint32_t *Es;      // [nchunks][Nloc][TL] class-sorted store entries (x0 TL + l) | (x1 TL + l) << 16, missing: 0x80000000 | l
	uint16_t *Zs;            // [nchunks][Nloc][TL] z0 | z1 << 8 in the same order
	const uint32_t *Hs;      // [nchunks][Nloc]     homozygotes | heterozygotes << 16
	const float *Pc;         // [nchunks][4 KP TL]  chunk-ordered P: float4 planes [k/4][x][l], then words [k][x][l]
	int32_t *n;              // [Lpad][2][KP]
	const float *Qf;
	const int2 *gpair;
	uint16_t *pcnt; double *plog; uint16_t *pnsh;     // the partials of ZQArgs
	Geometry geo;
	uint32_t iter, key0, key1;
	const uint32_t *iter_dev;
	uint32_t k_mant, k_one;
};
cudaError_t launch_zq_snp(const SnpArgs &a, int rounds, cudaStream_t s);
bool snp_eligible(const Geometry &g, int mode, int type_freq);
cudaError_t snp_configure(Geometry &g, int device);
size_t snp_pc_floats(const Geometry &g);
cudaError_t launch_snp_tile(const int16_t *Xt, uint32_t *Es, uint32_t *Hs, Geometry g, cudaStream_t s);
cudaError_t launch_snp_z_convert(const uint32_t *Es, uint16_t *Zs, int8_t *Zt, Geometry g, int to_sorted, cudaStream_t s);
cudaError_t launch_snp_pc_from_p(const float *P, float *Pc, Geometry g, cudaStream_t s);
// one element of the chunk-ordered P
__host__ __device__ inline void snp_pc_store(float *Pc, int tlc, int KP, int l, int x, int k, float v)
{
	float *b = Pc + (size_t)(l / tlc) * 4 * KP * tlc;
	const int ll = l % tlc;
	b[((size_t)(k / 4) * 2 * tlc + (size_t)x * tlc + ll) * 4 + (k % 4)] = v;
	b[(size_t)2 * KP * tlc + ((size_t)k * 2 + x) * tlc + ll] = v;
}

struct EpiArgs {
	const uint16_t *pcnt; const double *plog; const uint16_t *pnsh;
	const int32_t *nhet;     // [Nloc] usable heterozygous genotypes (data only; computed at load)
	int32_t *nsh;            // [Nloc] same-z heterozygotes on the current Z: read as the old-Z count, rewritten with the new one
	double *ind;             // [Npad][REC]
	float *Qf;               // [Nloc][KP]
	int32_t *cnt;            // [Nloc][K]
	double *llparts;         // [Nloc][4]
	const int2 *gpair;
	const DevScalars *sc;
	Geometry geo;
	uint32_t iter, key0, key1;
	int init;                // 1: initial assignment pass (no G accept, no likelihood)
	int type_freq;
	const uint32_t *iter_dev;
	int fmode;               // Geometry.fmode
	const double *S;         // mode 5: current F [N]
	const double *fprop;     // mode 5: proposed F [N]
	const float *pfk;        // mode 4: the sweep kernel's per-population partials
};
cudaError_t launch_epilogue(const EpiArgs &a, cudaStream_t s);
cudaError_t launch_fk_epilogue(const EpiArgs &a, cudaStream_t s);   // mode 4: per-population differences into the records

struct PArgs {
	int32_t *n; float *P; double *P64; const int32_t *allelenum;
	Geometry geo; uint32_t iter, key0, key1;
	const uint32_t *iter_dev;
	int mono_ok;             // 1: a locus with one allele gets P = 1 (update_P_auto has no allelenum > 1 guard, poly_geno.c:425)
	int sub;                 // 1: the second subgenome of the allotetraploid model (its own random stream)
	float *Pc;               // biallelic path: the chunk-ordered copy of P the sweep kernel stages (null otherwise)
	int tlc;
	int l0, nl;              // nl > 0: draw loci [l0, l0 + nl) only (a sharded chain draws 1/W of the loci per rank and all-gathers P)
};
cudaError_t launch_p_dirichlet(const PArgs &a, cudaStream_t s);

struct PreArgs {
	double *ind; double *S; const int32_t *state_in; int32_t *state_out; int32_t *gprop; int2 *gpair; DevScalars *sc;
	double *gpart;           // [2][SC_MAX_CTAS][20] partials of the grid-wide sums
	Geometry geo; uint32_t iter, key0, key1; int mode, prior_flag, back_refl;
	const uint32_t *iter_dev;
	double *fprop;           // modes 4/5: proposed inbreeding coefficients [K] / [N]
	float2 *hpair;           // mode 5: [Nloc] (1 - F, 1 - F')
	float *ftab;             // mode 4: table for the sweep kernel (ZQArgs.ftab)
	int grid;                // cooperative grid size of this context (scalar_grid), used when N > 1024
};
cudaError_t launch_pre_sweep(const PreArgs &a, cudaStream_t s);
int scalar_grid(int which /*0 pre, 1 post*/, int n_items, int device);

struct PostArgs {
	double *ind; DevScalars *sc; double *gpart; Geometry geo; uint32_t iter, key0, key1;
	const uint32_t *iter_dev;
	int mode, back_refl;
	double *S; const double *fprop; int32_t *state; const int32_t *state_prop;   // modes 4/5
	int grid;                // cooperative grid size of this context
};
cudaError_t launch_post_sweep(const PostArgs &a, cudaStream_t s);
// the same in two halves around an int64 all-reduce (sharded chains with local records): acc[K + 4]
cudaError_t launch_post_final(const PostArgs &a, const unsigned long long *acc, cudaStream_t s);

// update_S_POP of a sharded chain on the local records: proposal() summed for all 2^K subsets of replaced rates
// (acc[2^K][2]: fixed-point sum, -inf / NaN counts), and, after the all-reduce, the K decisions + the local G proposals
struct TreeArgs {
	const double *ind; const double *S; const int32_t *state; Geometry geo; uint32_t iter, key0, key1; int back_refl;
	unsigned long long *acc;
	double *S_out; int32_t *state_out; int32_t *gprop; int2 *gpair; DevScalars *sc;      // decide only
};

// All-reduce of the few KB of int64 sums over NVLink peer memory, fused with the post-sweep tail (ig_kernels.cu
// peer_allreduce_kernel).  Every rank owns one buffer that all ranks of the chain have mapped (CUDA IPC):
//   words [2 parities][W ranks][PX_WORDS] : rank r's contribution to the all-reduce of that parity, written BY rank r
//   words [2 parities][W ranks]           : rank r's sequence number, written by rank r after its data
constexpr int PX_WORDS = 32 + 2 * 1024;
//   words [2][W]                          : sequence numbers of the tally / P exchange below ("my tally is final", "my block of P is in your buffer")
static inline size_t px_buffer_words(int W) { return (size_t)2 * W * PX_WORDS + (size_t)4 * W; }
__host__ __device__ static inline size_t px_pflag_word(int W, int which) { return (size_t)2 * W * PX_WORDS + (size_t)(2 + which) * W; }

// The tally -> P exchange of a sharded sweep over peer memory (ig_kernels.cu p_peer_*): the arena every rank exports also
// holds its tally n and both P buffers.  Rank r owns loci block r: it PULLS that block of every rank's n over NVLink and adds
// (the reduce-scatter), draws the block's Dirichlets, and PUSHES the result into every rank's next-P buffer (the all-gather)
// -- one kernel between two single-CTA flag kernels.
struct PeerPArgs {
	unsigned long long *const *peers;    // device array [W] of arena bases
	int W, me;
	unsigned long long seq;
	size_t n_off, p_off;                 // byte offsets of n and of the P buffer being written, inside every arena
};
cudaError_t launch_p_peer_signal(const PeerPArgs &x, int which, cudaStream_t s);      // publish seq in flag array `which`, wait for all ranks'
cudaError_t launch_p_peer_draw(const PArgs &a, const PeerPArgs &x, cudaStream_t s);   // a.l0 / a.nl: this rank's block
struct PeerArgs {
	unsigned long long *const *peers;    // device array [W]: every rank's buffer as mapped here (own buffer at [me])
	int W, me;
	unsigned long long seq;              // 1, 2, 3, ... : the same on every rank for the same all-reduce
	int nwords;
	unsigned long long *acc;             // in: local sums (left zeroed)
	unsigned long long *out;             // global sums
	int do_final;                        // run the post-sweep tail (post_finish) on the summed totals at out[0 .. K + 4)
};
cudaError_t launch_peer_allreduce(const PeerArgs &x, const PostArgs &a, cudaStream_t s);
cudaError_t launch_local_sums(const TreeArgs *tree, const PostArgs *post, unsigned long long *post_acc, cudaStream_t s);   // either may be null
cudaError_t launch_spop_decide(const TreeArgs &a, cudaStream_t s);

struct MomArgs {
	const double *ind; const double *S; const DevScalars *sc; const float *P; Moments m;
	Geometry geo; int ns; long step; int convg_slot; int print_freq;
	const double *P64;       // print_freq: the allele frequencies in double, [K][L][A], as p_dirichlet drew them (the fp32 sweep copy is floored and rounded)
	int i_lo, n_ind;         // individuals whose moments this launch advances: all N, or the shard's own when the records are local
};
cudaError_t launch_moments(const MomArgs &a, cudaStream_t s);
cudaError_t launch_moments_reset(const MomArgs &a, cudaStream_t s);

// layout transforms and stand-alone evaluators (tests, state injection)
cudaError_t launch_tile_x(const int16_t *x_canon, int16_t *Xt, const int32_t *allelenum, Geometry g, cudaStream_t s);
cudaError_t launch_untile_x(const int16_t *Xt, int16_t *x_canon, Geometry g, cudaStream_t s);
cudaError_t launch_tile_z(const int8_t *z_canon, int8_t *Zt, Geometry g, cudaStream_t s);
cudaError_t launch_untile_z(const int8_t *Zt, int8_t *z_canon, Geometry g, cudaStream_t s);
cudaError_t launch_het_counts(const int16_t *Xt, const int8_t *Zt, int32_t *nhet, int32_t *nsh, Geometry g, cudaStream_t s);
cudaError_t launch_tally(const int16_t *Xt, const int8_t *Zt, int32_t *n, Geometry g, cudaStream_t s);
cudaError_t launch_loglik(const int16_t *Xt, const int8_t *Zt, const float *P, const float *Qf, const int32_t *gen,
                          double *out, Geometry g, int type_freq, cudaStream_t s);
cudaError_t launch_fill_f32(float *p, float v, size_t n, cudaStream_t s);
cudaError_t launch_fill_q_uniform(float *Qf, Geometry g, cudaStream_t s);
cudaError_t launch_init_chain(double *ind, double *S, int32_t *state, DevScalars *sc, const float *initd_dev,
                              int32_t *gprop, int2 *gpair, Geometry g, int mode, int prior_flag, int back_refl,
                              uint32_t key0, uint32_t key1, cudaStream_t s);
cudaError_t launch_proposal_ll(const double *ind, const double *S, double *out, Geometry g, cudaStream_t s);
cudaError_t launch_qf_from_ind(const double *ind, float *Qf, Geometry g, cudaStream_t s);
cudaError_t launch_pack_g(const double *ind, uint8_t *g8, Geometry g, cudaStream_t s);     // G of every individual as one byte (DP host step)

}  // namespace ig
