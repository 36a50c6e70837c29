// tetra.cu -- the autotetraploid sweep (`-p 4 -ap 1`) on sm_100a.
//
// Reference order (mcmc_POP_tetra_selfing, poly_geno.c:97-112):
//     update_P_auto -> calc_exfreq_auto -> update_S_POP -> update_ZQ -> update_geno -> cal_lkd
// where update_S_POP alone makes 2K+1 full passes over the data (cal_lkd_props - cal_lkd per
// population, poly_geno.c:625) and cal_lkd a further one.  Here a sweep is TWO passes:
//
//   p_dirichlet     P | n                              update_P_auto draw half      :425-434
//   tetra_propose   S'_k = reflect(S_k + U(-.05,.05))  update_S_POP proposals       :604-612
//   tetra_tables    log HWE table R, then the selfing tables of S and of S'          calc_exfreq_auto :1515,
//                   for every (locus, population)                                    auto_genfreq :1803
//   tetra_zs        PASS A over (geno, old z): the S statistics D_k = sum over genotypes whose four
//                   copies all sit in k of (table_S'[g] - table_S[g])  -- everything else cancels in
//                   cal_lkd_props(k) - cal_lkd(), and the K steps do not interact, so one pass
//                   serves all K Metropolis steps;  then the new z | P, Q and the counts  :750-836
//   tetra_accept    the K accept decisions, S and the tables of the accepted proposals   :628-634
//   tetra_q         Q_i ~ Dirichlet(cnt_i + alpha)                                       :831-833
//   tetra_geno      PASS B over (x, new z): dosage resolution | z, Q, tables (update_geno :520,
//                   choose_two_auto :854, choose_tri_auto :907), the likelihood of the result
//                   (cal_lkd :715, calc_genofq :1235) and the tally n[l][a][k] of the next update_P_auto
//   tetra_lkh       indvlkh, totallkh, column sums of Q
//
// update_geno needs the Q drawn from the finished Z of the same sweep, which is why there are two
// passes and not one.  Bytes per allele copy: pass A 3 (geno, old z, new z), pass B 4 (int16 x,
// z, new geno) against the 6 of SURVEY.md section 8d.
//
// Multi-GPU: chains over GPUs, or the individuals of one chain sharded (int32 all-reduce of n, all-gather of
// the per-individual S statistics and records; fixed-order reductions => bit-identical to one GPU).
// Restrictions of this version: -e 1 only
// (with -e 0 the reference's own tables are log(0), tests/test_tetra_oracle_vs_reference.py),
// allelenum_max <= 10, alpha is never updated (the reference's tetraploid driver never calls
// update_alpha).
#include <stdlib.h>
#include "ig_ctx.h"
#include "philox.cuh"
#include "samplers.cuh"
#include "sweep_common.cuh"

namespace ig {

constexpr int TT = 4;                 // loci per micro-tile: one 128-bit z / geno vector, two 128-bit x vectors
constexpr int TETRA_MAX_A = 10;       // catalogue of 715 genotypes (autotetraploid) / 3025 (allotetraploid): 16-bit indices
constexpr int TETRA_THREADS = 256;
#define LN2_D 0.69314718055994530942

// genotype catalogue of one allele count n (auto_geno_num / auto_geno_list, poly_geno.c:1698-1800):
// codes are the four alleles read as a base-n number, grouped by dosage class
struct TetraCat {
	int n;
	int cls[5];      // iiii, iiij, iijj, iijk, ijkl
	int total;
	int code_off;    // into codes[]
	int c2i_off;     // into c2i[] (n^4 entries: code -> index, 0xFFFF = not a catalogue genotype)
};

struct TetraState {
	int ncat = 0, Gmax = 0, Lq = 0, LTq = 0;
	TetraCat cat_h[TETRA_MAX_A + 1];
	int16_t *Xq = nullptr;            // [LTq][Nloc][TT][4] distinct alleles ascending, -1 padding
	int8_t *Zq = nullptr, *Gq = nullptr;   // [LTq][Nloc][TT][4]; geno = -1 on missing genotypes
	TetraCat *cats = nullptr;
	int32_t *codes = nullptr;
	uint16_t *c2i = nullptr;
	int32_t *loc_cat = nullptr;       // [Lq] catalogue of each locus (-1 beyond L)
	float *exf = nullptr, *tabC = nullptr, *tabP = nullptr;    // [Lq][K][Gmax] natural logs
	double *Sprop = nullptr, *dstat = nullptr;                  // [K]
	int32_t *accepted = nullptr;      // [K]
	unsigned long long *dfix = nullptr;   // [MAX_K] S statistics cal_lkd_props(k) - cal_lkd() of the current sweep, 2^-24 fixed point
	float *lpart = nullptr;           // [nchunks][2][Nloc] likelihood partials: natural-log part, log2 part
	bool timing = false;              // a profiled sweep is between its PASS A start and PASS B end events
	// allotetraploid (-ap 0): copies 0,1 and copies 2,3 are two subgenomes with their own allele frequencies
	bool allo = false;
	float *P2 = nullptr;              // UPMCMC.freq2, [Lpad][A][KP]
	int32_t *n2 = nullptr;            // tally of the second subgenome, same indexing
};

// --------------------------------------------------------------------------------------
// catalogue (host)
// --------------------------------------------------------------------------------------
static void build_catalogue(int n, std::vector<int> &code, int cls[5])
{
	cls[0] = n; cls[1] = n * (n - 1); cls[2] = n * (n - 1) / 2;
	cls[3] = n * (n - 1) * (n - 2) / 2; cls[4] = n * (n - 1) * (n - 2) * (n - 3) / 24;
	const int n2 = n * n, n3 = n2 * n;
	for (int j = 0; j < n; j++) code.push_back(j * (n3 + n2 + n + 1));
	for (int j = 0; j < n - 1; j++) for (int k = j + 1; k < n; k++) { code.push_back(j * (n3 + n2 + n) + k); code.push_back(k * (n3 + n2 + n) + j); }
	for (int j = 0; j < n - 1; j++) for (int k = j + 1; k < n; k++) code.push_back(j * (n3 + n2) + k * (n + 1));
	for (int j = 0; j < n - 2; j++) for (int k = j + 1; k < n - 1; k++) for (int a = k + 1; a < n; a++) {
		code.push_back(j * (n3 + n2) + k * n + a);
		code.push_back(k * (n3 + n2) + j * n + a);
		code.push_back(a * (n3 + n2) + j * n + k);
	}
	for (int j = 0; j < n - 3; j++) for (int k = j + 1; k < n - 2; k++) for (int a = k + 1; a < n - 1; a++) for (int b = a + 1; b < n; b++)
		code.push_back(j * n3 + k * n2 + a * n + b);
}

// allo_geno_num / allo_geno_list, poly_geno.c:2031-2120: iikk | iikl (k<l) | ijkk (i<j) | ijkl (i<j, k<l)
static void build_catalogue_allo(int n, std::vector<int> &code, int cls[5])
{
	cls[0] = n * n; cls[1] = n * (n - 1) / 2 * n; cls[2] = n * (n - 1) / 2 * n; cls[3] = n * (n - 1) * n * (n - 1) / 4; cls[4] = 0;
	const int n2 = n * n, n3 = n2 * n;
	for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) code.push_back(j * n2 * (n + 1) + k * (n + 1));
	for (int j = 0; j < n; j++) for (int k = 0; k < n - 1; k++) for (int a = k + 1; a < n; a++) code.push_back(j * n2 * (n + 1) + n * k + a);
	for (int j = 0; j < n - 1; j++) for (int k = j + 1; k < n; k++) for (int a = 0; a < n; a++) code.push_back((j * n + k) * n2 + a * (n + 1));
	for (int j = 0; j < n - 1; j++) for (int k = j + 1; k < n; k++) for (int a = 0; a < n - 1; a++) for (int b = a + 1; b < n; b++)
		code.push_back(j * n3 + k * n2 + a * n + b);
}

// --------------------------------------------------------------------------------------
// layout transforms: canonical [L][Nloc][4] <-> tiled [LTq][Nloc][TT][4]
// --------------------------------------------------------------------------------------
template <typename T>
__global__ void tile4_kernel(const T *canon, T *tiled, int L, int Nloc, int LTq, T fill)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (size_t)LTq * Nloc * TT) return;
	const int j = (int)(t % TT), i = (int)((t / TT) % Nloc), mt = (int)(t / ((size_t)TT * Nloc));
	const int l = mt * TT + j;
	for (int c = 0; c < 4; c++) tiled[t * 4 + c] = (l < L) ? canon[((size_t)l * Nloc + i) * 4 + c] : fill;
}
template <typename T>
__global__ void untile4_kernel(const T *tiled, T *canon, int L, int Nloc)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (size_t)L * Nloc) return;
	const int i = (int)(t % Nloc), l = (int)(t / Nloc);
	const size_t src = (((size_t)(l / TT) * Nloc + i) * TT + (l % TT)) * 4;
	for (int c = 0; c < 4; c++) canon[t * 4 + c] = tiled[src + c];
}
static inline unsigned nb(size_t n, int b) { return (unsigned)((n + b - 1) / b); }

// --------------------------------------------------------------------------------------
// tables: one thread per (locus, population)
// --------------------------------------------------------------------------------------
struct TabArgs {
	const float *P; const int32_t *allelenum; const int32_t *loc_cat; const TetraCat *cats; const int32_t *codes; const uint16_t *c2i;
	const double *S, *Sprop; float *exf, *tabC, *tabP;
	int L, K, KP, A, Gmax, do_cur, do_prop;
	const float *P2; int allo;
};

// gaussj (poly_geno.c:2384) for the 3x3 float system of the triallelic class: Gauss-Jordan with
// full pivoting, same pivot choice and elimination order
__device__ void solve3(float a[3][3], float b[3])
{
	int piv[3] = {0, 0, 0}, row = 0, col = 0;
	for (int i = 0; i < 3; i++) {
		float big = 0.0f;
		for (int j = 0; j < 3; j++)
			if (piv[j] != 1)
				for (int k = 0; k < 3; k++)
					if (piv[k] == 0 && fabsf(a[j][k]) >= big) { big = fabsf(a[j][k]); row = j; col = k; }
		++piv[col];
		if (row != col) {
			float t;
			for (int k = 0; k < 3; k++) { t = a[row][k]; a[row][k] = a[col][k]; a[col][k] = t; }
			t = b[row]; b[row] = b[col]; b[col] = t;
		}
		const float pivinv = (float)(1.0 / (double)a[col][col]);
		a[col][col] = 1.0f;
		for (int k = 0; k < 3; k++) a[col][k] *= pivinv;
		b[col] *= pivinv;
		for (int j = 0; j < 3; j++)
			if (j != col) {
				const float dum = a[j][col];
				a[j][col] = 0.0f;
				for (int k = 0; k < 3; k++) a[j][k] -= a[col][k] * dum;
				b[j] -= b[col] * dum;
			}
	}
}
__device__ __forceinline__ bool in3(int v, const int *d, int len) { for (int i = 0; i < len; i++) if (d[i] == v) return true; return false; }
// calc_val / calc_val2 (poly_geno.c:2307,2333): the ijkl code holding the given alleles
__device__ int quad_with(const int *d, int v, int n)
{
	const int n2 = n * n, n3 = n2 * n;
	if (v < d[0]) return v * n3 + d[0] * n2 + d[1] * n + d[2];
	if (v > d[0] && v < d[1]) return d[0] * n3 + v * n2 + d[1] * n + d[2];
	if (v > d[1] && v < d[2]) return d[0] * n3 + d[1] * n2 + v * n + d[2];
	if (v > d[2]) return d[0] * n3 + d[1] * n2 + d[2] * n + v;
	return 0;
}
__device__ int quad_with2(const int *d, int v1, int v2, int n)
{
	const int n2 = n * n, n3 = n2 * n;
	if (v2 < d[1]) return v1 * n3 + v2 * n2 + d[1] * n + d[0];
	if (v2 > d[1] && v2 < d[0] && v1 < d[1]) return v1 * n3 + d[1] * n2 + v2 * n + d[0];
	if (v1 > d[1] && v2 < d[0]) return d[1] * n3 + v1 * n2 + v2 * n + d[0];
	if (v1 > d[1] && v1 < d[0] && v2 > d[0]) return d[1] * n3 + v1 * n2 + d[0] * n + v2;
	if (v1 > d[0]) return d[1] * n3 + d[0] * n2 + v1 * n + v2;
	if (v1 < d[1] && v2 > d[0]) return v1 * n3 + d[1] * n2 + n * d[0] + v2;
	return 0;
}

// auto_genfreq (poly_geno.c:1803-2028): log genotype frequencies under partial selfing, solved
// class by class from the most to the least heterozygous, (I - sA) P = (1 - s) R.  Same float /
// double mix and operation order as the reference; its re-use of a stale index in the
// monoallelic class (:1984-1989) is reproduced as written.
__device__ void genfreq_locus(float self, const TetraCat &c, const int32_t *code, const uint16_t *c2i, const float *R, float *P)
{
	const int n = c.n, n2 = n * n;
	int hi = c.total, num = 0, d[3];
	float temp;
#define IDX(cd) ((int)c2i[(cd)])
	if (n >= 4)
		for (int i = hi - c.cls[4]; i < hi; i++) P[i] = (float)(log((double)(1 - self)) + (double)R[i] - log((double)(1 - self / 6)));
	hi -= c.cls[4];
	if (n >= 3)
		for (int i = 0; i < c.cls[3] / 3; i++) {
			const int base = hi - c.cls[3] + i * 3;
			float A[3][3], v[3];
			num = code[base];
			for (int j = 2; j >= 0; j--) { d[j] = num % n; num /= n; }
			temp = 0;
			if (n >= 4)
				for (int q = 0; q < n; q++)
					if (!in3(q, d, 3)) { num = IDX(quad_with(d, q, n)); temp = (float)((double)temp + exp((double)P[num])); }
			for (int j = 0; j < 3; j++) {
				for (int q = 0; q < 3; q++) A[j][q] = (j == q) ? (float)(1 - (double)self * 10.0 / 36.0) : (float)(-(double)self / 9.0);
				v[j] = (float)((double)self / 18.0 * (double)temp + (1.0 - (double)self) * exp((double)R[base + j]));
			}
			temp = v[0];
			for (int j = 0; j < 3; j++) v[j] /= temp;
			solve3(A, v);
			for (int j = 0; j < 3; j++) P[base + j] = (float)(log((double)v[j]) + log((double)temp));
		}
	hi -= c.cls[3];
	for (int i = hi - c.cls[2]; i < hi; i++) {                                  // iijj
		num = code[i];
		d[0] = num % n; num /= n2; d[1] = num % n;
		temp = 0;
		if (n >= 3)
			for (int j = 0; j < n; j++) {
				if (in3(j, d, 2)) continue;
				if (d[0] < j) num = IDX(d[1] * n2 * (n + 1) + d[0] * n + j);
				else if (d[0] > j) num = IDX(d[1] * n2 * (n + 1) + j * n + d[0]);
				temp = (float)((double)temp + exp((double)P[num]) / 9.0 * (double)self);
				if (d[1] < j) num = IDX(d[0] * n2 * (n + 1) + d[1] * n + j);
				else if (d[1] > j) num = IDX(d[0] * n2 * (n + 1) + j * n + d[1]);
				temp = (float)((double)temp + exp((double)P[num]) / 9.0 * (double)self);
				num = IDX(j * n2 * (n + 1) + d[1] * n + d[0]);
				temp = (float)((double)temp + exp((double)P[num]) / 36.0 * (double)self);
				if (n >= 4)
					for (int q = j + 1; q < n; q++)
						if (!in3(q, d, 2)) { num = IDX(quad_with2(d, j, q, n)); temp = (float)((double)temp + exp((double)P[num]) / 36.0 * (double)self); }
			}
		P[i] = (float)(log((double)(1 - self) * exp((double)R[i]) + (double)temp) - log(1 - (double)self / 2.0));
	}
	hi -= c.cls[2];
	for (int i = hi - c.cls[1]; i < hi; i++) {                                  // iiij
		num = code[i];
		d[0] = num % n; num /= n; d[1] = num % n;
		if (d[0] < d[1]) num = IDX((d[0] * n2 + d[1]) * (n + 1));
		else if (d[0] > d[1]) num = IDX((d[1] * n2 + d[0]) * (n + 1));
		temp = (float)(8.0 / 36.0 * exp((double)P[num]) * (double)self);
		if (n >= 3)
			for (int j = 0; j < n; j++) {
				if (in3(j, d, 2)) continue;
				if (d[0] < j) num = IDX(d[1] * n2 * (n + 1) + d[0] * n + j);
				else if (d[0] > j) num = IDX(d[1] * n2 * (n + 1) + j * n + d[0]);
				temp = (float)((double)temp + exp((double)P[num]) / 9.0 * (double)self);
			}
		P[i] = (float)(log((double)(1 - self) * exp((double)R[i]) + (double)temp) - log(1 - (double)self / 2.0));
	}
	hi -= c.cls[1];
	for (int i = hi - c.cls[0]; i < hi; i++) {                                  // iiii
		num = code[i];
		d[0] = num % n;
		temp = 0;
		for (int j = 0; j < n; j++) {
			if (j == d[0]) continue;
			num = IDX(d[0] * n * (n2 + n + 1) + j);
			temp = (float)((double)temp + exp((double)P[num]) / 4.0 * (double)self);
			if (d[0] < j) num = IDX(d[0] * n2 * (n + 1) + j * (n + 1));       // as written: otherwise the iiij index is re-used
			temp = (float)((double)temp + exp((double)P[num]) / 36.0 * (double)self);
			if (n >= 3)
				for (int q = j + 1; q < n; q++)
					if (q != d[0]) { num = IDX(d[0] * n2 * (n + 1) + j * n + q); temp = (float)((double)temp + exp((double)P[num]) / 36.0 * (double)self); }
		}
		P[i] = (float)(log((double)(1 - self) * exp((double)R[i]) + (double)temp) - log((double)(1 - self)));
	}
#undef IDX
}

// allo_genfreq, poly_geno.c:2122-2305: (I - sA) P = (1 - s) R solved class by class, from the doubly
// heterozygous genotypes down; operation order and float / double mix of the reference
__device__ void genfreq_locus_allo(float self, const TetraCat &c, const int32_t *code, const uint16_t *c2i, const float *R, float *P)
{
	const int n = c.n, n2 = n * n, n3 = n2 * n;
#define IDX(cd) ((int)c2i[(cd)])
	int tmp = c.total, num, d[3];
	float temp;
	for (int i = tmp - c.cls[3]; i < tmp; i++)                            // ijkl
		P[i] = (float)(log((double)(1 - self)) + (double)R[i] - log((double)(1 - self / 4)));
	tmp -= c.cls[3];
	for (int i = tmp - c.cls[2]; i < tmp; i++) {                          // ijkk
		num = code[i];
		d[0] = num % n; num /= n2; d[1] = num % n; num /= n; d[2] = num % n;
		temp = 0;
		for (int j = 0; j < n; j++)
			if (j != d[0]) {
				num = (d[0] < j) ? IDX(d[2] * n3 + d[1] * n2 + d[0] * n + j) : IDX(d[2] * n3 + d[1] * n2 + j * n + d[0]);
				temp = (float)((double)temp + exp((double)P[num]) * (double)self / 8.0);
			}
		P[i] = (float)(log((double)(1 - self) * exp((double)R[i]) + (double)temp) - log(1 - (double)self / 2.0));
	}
	tmp -= c.cls[2];
	for (int i = tmp - c.cls[1]; i < tmp; i++) {                          // iikl
		num = code[i];
		d[0] = num % n; num /= n; d[1] = num % n; num /= n; d[2] = num % n;
		temp = 0;
		for (int j = 0; j < n; j++)
			if (j != d[2]) {
				num = (d[2] < j) ? IDX(d[2] * n3 + j * n2 + d[1] * n + d[0]) : IDX(j * n3 + d[2] * n2 + d[1] * n + d[0]);
				temp = (float)((double)temp + exp((double)P[num]) * (double)self / 8.0);
			}
		P[i] = (float)(log((double)(1 - self) * exp((double)R[i]) + (double)temp) - log(1 - (double)self / 2.0));
	}
	tmp -= c.cls[1];
	for (int i = tmp - c.cls[0]; i < tmp; i++) {                          // iikk
		num = code[i];
		d[0] = num % n; num /= n2; d[1] = num % n;
		temp = 0;
		for (int j = 0; j < n; j++)
			if (j != d[0]) {
				num = (d[0] < j) ? IDX(d[1] * n2 * (n + 1) + d[0] * n + j) : IDX(d[1] * n2 * (n + 1) + j * n + d[0]);
				temp = (float)((double)temp + exp((double)P[num]) * (double)self / 4.0);
			}
		for (int j = 0; j < n; j++)
			if (j != d[1]) {
				num = (d[1] < j) ? IDX(d[1] * n3 + j * n2 + d[0] * (n + 1)) : IDX(j * n3 + d[1] * n2 + d[0] * (n + 1));
				temp = (float)((double)temp + exp((double)P[num]) * (double)self / 4.0);
			}
		for (int j = 0; j < n; j++)
			for (int kk = 0; kk < n; kk++)
				if (j != d[1] && kk != d[0]) {
					const int a0 = d[1] < j ? d[1] : j, a1 = d[1] < j ? j : d[1], b0 = d[0] < kk ? d[0] : kk, b1 = d[0] < kk ? kk : d[0];
					num = IDX(a0 * n3 + a1 * n2 + b0 * n + b1);
					temp = (float)((double)temp + exp((double)P[num]) * (double)self / 16.0);
				}
		P[i] = (float)(log((double)(1 - self) * exp((double)R[i]) + (double)temp) - log((double)(1 - self)));
	}
#undef IDX
}

__global__ void tetra_tables_kernel(const TabArgs a)
{
	// when both tables are wanted (every sweep), two threads share a (locus, population): one solves for the current
	// rate, the other for the proposed one; both write the same exfreq values first (identical stores)
	const int split = (a.do_cur && a.do_prop) ? 2 : 1;
	const int t2 = blockIdx.x * blockDim.x + threadIdx.x;
	const int t = t2 / split;
	if (t >= a.L * a.K) return;
	const bool second = (split == 2) && (t2 & 1);
	const bool want_cur = (split == 2) ? !second : (a.do_cur != 0), want_prop = (split == 2) ? second : (a.do_prop != 0);
	const int l = t / a.K, k = t % a.K;
	const TetraCat c = a.cats[a.loc_cat[l]];
	const int32_t *code = a.codes + c.code_off;
	const uint16_t *c2i = a.c2i + c.c2i_off;
	const int n = c.n;
	float *R = a.exf + (size_t)t * a.Gmax;
	double lf[TETRA_MAX_A];
	for (int al = 0; al < n; al++) lf[al] = log((double)a.P[((size_t)l * a.A + al) * a.KP + k]);
	if (a.allo) {
		// calc_exfreq_allo, poly_geno.c:1592-1671: subgenome 1 (copies 0,1) with freq, subgenome 2 with freq2
		double lf2[TETRA_MAX_A];
		for (int al = 0; al < n; al++) lf2[al] = log((double)a.P2[((size_t)l * a.A + al) * a.KP + k]);
		int lo = 0, d[4], w;
		for (int g = lo; g < lo + c.cls[0]; g++) { w = code[g]; d[0] = w % n; w /= (n * n); d[1] = w % n; R[g] = (float)((lf[d[1]] + lf2[d[0]]) * 2); }
		lo += c.cls[0];
		for (int g = lo; g < lo + c.cls[1]; g++) {
			w = code[g];
			for (int q = 0; q < 3; q++) { d[q] = w % n; w /= n; }
			R[g] = (float)(log(2.0) + lf[d[2]] * 2 + lf2[d[0]] + lf2[d[1]]);
		}
		lo += c.cls[1];
		for (int g = lo; g < lo + c.cls[2]; g++) {
			w = code[g]; w /= n;
			for (int q = 0; q < 3; q++) { d[q] = w % n; w /= n; }
			R[g] = (float)(log(2.0) + lf2[d[0]] * 2 + lf[d[2]] + lf[d[1]]);
		}
		lo += c.cls[2];
		for (int g = lo; g < lo + c.cls[3]; g++) {
			w = code[g];
			for (int q = 0; q < 4; q++) { d[q] = w % n; w /= n; }
			float r = (float)log(4.0);
			for (int q = 0; q < 2; q++) r += (float)lf2[d[q]];
			for (int q = 2; q < 4; q++) r += (float)lf[d[q]];
			R[g] = r;
		}
		if (split == 2) genfreq_locus_allo((float)(second ? a.Sprop[k] : a.S[k]), c, code, c2i, R, (second ? a.tabP : a.tabC) + (size_t)t * a.Gmax);
		else {
			if (want_cur) genfreq_locus_allo((float)a.S[k], c, code, c2i, R, a.tabC + (size_t)t * a.Gmax);
			if (want_prop) genfreq_locus_allo((float)a.Sprop[k], c, code, c2i, R, a.tabP + (size_t)t * a.Gmax);
		}
		return;
	}
	// calc_exfreq_auto, poly_geno.c:1515-1577
	int lo = 0, d[4], w;
	for (int g = lo; g < lo + c.cls[0]; g++) R[g] = (float)lf[code[g] % n] * 4.0f;
	lo += c.cls[0];
	for (int g = lo; g < lo + c.cls[1]; g++) { w = code[g]; d[0] = w % n; w /= n; d[1] = w % n; R[g] = (float)(log(4.0) + lf[d[1]] * (double)3.0f + lf[d[0]]); }
	lo += c.cls[1];
	for (int g = lo; g < lo + c.cls[2]; g++) { w = code[g]; d[0] = w % n; w /= (n * n); d[1] = w % n; R[g] = (float)(log(6.0) + (lf[d[1]] + lf[d[0]]) * 2); }
	lo += c.cls[2];
	for (int g = lo; g < lo + c.cls[3]; g++) {
		w = code[g];
		for (int q = 0; q < 3; q++) { d[q] = w % n; w /= n; }
		R[g] = (float)(log(12.0) + lf[d[2]] * 2 + lf[d[0]] + lf[d[1]]);
	}
	lo += c.cls[3];
	for (int g = lo; g < lo + c.cls[4]; g++) {
		w = code[g];
		for (int q = 0; q < 4; q++) { d[q] = w % n; w /= n; }
		float r = (float)log(24.0);
		for (int q = 0; q < 4; q++) r += (float)lf[d[q]];
		R[g] = r;
	}
	// calc_self_genofreq, poly_geno.c:1219: the rate arrives as double and is passed on as float
	if (split == 2) genfreq_locus((float)(second ? a.Sprop[k] : a.S[k]), c, code, c2i, R, (second ? a.tabP : a.tabC) + (size_t)t * a.Gmax);
	else {
		if (want_cur) genfreq_locus((float)a.S[k], c, code, c2i, R, a.tabC + (size_t)t * a.Gmax);
		if (want_prop) genfreq_locus((float)a.Sprop[k], c, code, c2i, R, a.tabP + (size_t)t * a.Gmax);
	}
}

// --------------------------------------------------------------------------------------
// proposals and accepts of update_S_POP (poly_geno.c:604-634), -e 1
// --------------------------------------------------------------------------------------
__global__ void tetra_propose_kernel(const double *S, double *Sprop, int K, uint32_t iter, uint32_t key0, uint32_t key1)
{
	const int k = threadIdx.x;
	if (k >= K) return;
	Stream st((uint32_t)k, 0u, iter, TAG_TETRA, key0, key1);
	double p = st.uniform() * 2 * 0.05 - 0.05;
	p += S[k];
	if (p <= 0.0) p = 0.0 - p;
	else if (p >= 1.0) p = 1.0 - (p - 1.0);
	Sprop[k] = p;
}

constexpr int RED1 = 1024;
__device__ double block_sum1(double v, double *sh)
{
	const int tid = threadIdx.x;
	__syncthreads();
	sh[tid] = v;
	__syncthreads();
	for (int s = RED1 / 2; s > 0; s >>= 1) { if (tid < s) sh[tid] += sh[tid + s]; __syncthreads(); }
	const double r = sh[0];
	__syncthreads();
	return r;
}

// D_k = cal_lkd_props(k) - cal_lkd() from the fixed-point totals of PASS A (all shards, all-reduced as
// integers: identical on every rank and for every shard count), then the K accept decisions
__global__ void tetra_accept_kernel(const unsigned long long *dfix, int K, double *S, const double *Sprop,
                                    double *dstat, int32_t *accepted, DevScalars *sc, uint32_t iter, uint32_t key0, uint32_t key1, int decide)
{
	const int k = threadIdx.x;
	if (k >= K) return;
	const double D = (double)(long long)dfix[k] * (1.0 / 16777216.0);
	dstat[k] = D;
	if (decide) {
		Stream st((uint32_t)k, 0u, iter, TAG_TETRA, key0, key1);
		(void)st.uniform();                                       // the proposal's draw
		const double u = st.uniform();
		// ran1() < exp(MIN2(0, mhratio)) (poly_geno.c:628)
		const bool acc = (u < exp(fmin(0.0, D)));
		accepted[k] = acc ? 1 : 0;
		if (acc) { S[k] = Sprop[k]; atomicAdd(&sc->s_accepts, 1); }
	}
}
// move_genofreq, poly_geno.c:738-748
__global__ void tetra_select_kernel(float *tabC, const float *tabP, const int32_t *accepted, int L, int K, int Gmax)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (size_t)L * K * Gmax) return;
	const int k = (int)((t / Gmax) % K);
	if (accepted[k]) tabC[t] = tabP[t];
}

// --------------------------------------------------------------------------------------
// PASS A: S statistics on the old z, then the new z and the ancestry counts
// --------------------------------------------------------------------------------------
struct ZsArgs {
	int8_t *Zq; const int8_t *Gq; const float *P; const float *Qf;
	const float *tabC, *tabP; const int32_t *loc_cat; const TetraCat *cats; const uint16_t *c2i;
	uint16_t *pcnt; unsigned long long *dfix;   // dfix [K]: S statistics, 2^-24 fixed point, summed over the launch
	Geometry geo; int Gmax; int init;
	uint32_t iter, key0, key1, k_mant, k_one;
};

template <int KP, int ROUNDS>
__global__ void __launch_bounds__(TETRA_THREADS, 3) tetra_zs_kernel(const ZsArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ __align__(8) unsigned long long bar;
	__shared__ unsigned long long dacc[MAX_K];              // this CTA's S statistics, fixed point
	const Geometry &g = a.geo;
	const int tid = threadIdx.x, chunk = blockIdx.x;
	const int l0 = chunk * g.TL, nl = min(g.TL, g.Lpad - l0), nmt = nl / TT, rowsz = g.A * KP;
	float *Psm = reinterpret_cast<float *>(smem_raw);
	int *cntsm = reinterpret_cast<int *>(Psm + (size_t)g.TL * rowsz);           // [KP][threads]
	float *dsm = reinterpret_cast<float *>(cntsm + KP * TETRA_THREADS);         // [KP][threads]
	int2 *locsm = reinterpret_cast<int2 *>(dsm + KP * TETRA_THREADS);           // [TL] (n, c2i offset)
	const int nbins = nl * rowsz;
	if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
	__syncthreads();
	if (tid == 0) { mbar_expect_tx(&bar, (uint32_t)nbins * 4u); tma_bulk_g2s(Psm, a.P + (size_t)l0 * rowsz, (uint32_t)nbins * 4u, &bar); }
	for (int k = 0; k < KP; k++) { cntsm[k * TETRA_THREADS + tid] = 0; dsm[k * TETRA_THREADS + tid] = 0.0f; }
	if (tid < MAX_K) dacc[tid] = 0ull;
	for (int j = tid; j < nl; j += TETRA_THREADS) {
		const int ci = (l0 + j < g.L) ? a.loc_cat[l0 + j] : -1;
		locsm[j] = ci >= 0 ? make_int2(a.cats[ci].n, a.cats[ci].c2i_off) : make_int2(1, 0);
	}
	__syncthreads();
	mbar_wait(&bar, 0);

	const int Nloc = g.Nloc, mt0 = l0 / TT;
	const int nsub_total = (Nloc + TETRA_THREADS - 1) / TETRA_THREADS;
	const int sub0 = blockIdx.y * g.subs_per_blk, sub1 = min(sub0 + g.subs_per_blk, nsub_total);
	const uint32_t psm = smem_addr(Psm);
	const RegConst kc{a.k_mant, a.k_one};
	const RegConst kc16{U16_MANT, U16_ONE};
	const float cnt_t = as_dn(smem_addr(cntsm) + (uint32_t)tid * 4u);

	for (int sub = sub0; sub < sub1; ++sub) {
		const int il = sub * TETRA_THREADS + tid;
		if (il < Nloc) {
		float q[KP];
#pragma unroll
		for (int v = 0; v < KP / 4; v++) {
			const float4 w = __ldg(reinterpret_cast<const float4 *>(a.Qf + (size_t)il * KP) + v);
			q[4 * v] = w.x; q[4 * v + 1] = w.y; q[4 * v + 2] = w.z; q[4 * v + 3] = w.w;
		}
		const uint32_t ig_global = (uint32_t)(g.i0 + il);
		int4 *zp = reinterpret_cast<int4 *>(a.Zq) + ((size_t)mt0 * Nloc + il);
		const int4 *gp = reinterpret_cast<const int4 *>(a.Gq) + ((size_t)mt0 * Nloc + il);
		// software prefetch, one micro-tile ahead (as in zq_sweep): four warps per scheduler do not cover HBM latency
		int4 gv_n = ldg_stream(gp), zv_n = ldg_rw(zp);
		for (int mt = 0; mt < nmt; ++mt) {
			const int4 gv = gv_n, zv = zv_n;
			if (mt + 1 < nmt) { gv_n = ldg_stream(gp + (size_t)(mt + 1) * Nloc); zv_n = ldg_rw(zp + (size_t)(mt + 1) * Nloc); }
			const uint32_t gw[4] = {(uint32_t)gv.x, (uint32_t)gv.y, (uint32_t)gv.z, (uint32_t)gv.w};
			const uint32_t zo[4] = {(uint32_t)zv.x, (uint32_t)zv.y, (uint32_t)zv.z, (uint32_t)zv.w};
			uint32_t zn[4];
			// 16 random bits per allele copy (sweep_common.cuh): two Philox blocks serve the micro-tile's 16 copies
			const u32x4 rndA = philox4x32<ROUNDS>(u32x4{(uint32_t)(mt0 + mt), ig_global, a.iter, TAG_Z16 | 0u}, a.key0, a.key1);
			const u32x4 rndB = philox4x32<ROUNDS>(u32x4{(uint32_t)(mt0 + mt), ig_global, a.iter, TAG_Z16 | 1u}, a.key0, a.key1);
			const uint32_t rw[8] = {rndA.x, rndA.y, rndA.z, rndA.w, rndB.x, rndB.y, rndB.z, rndB.w};
#pragma unroll
			for (int j = 0; j < TT; ++j) {
				zn[j] = zo[j];
				if ((gw[j] & 0x80u) != 0u) continue;                      // missing genotype: geno = -1
				const int lj = mt * TT + j;
				const uint32_t ga[4] = {gw[j] & 0xFFu, (gw[j] >> 8) & 0xFFu, (gw[j] >> 16) & 0xFFu, gw[j] >> 24};
				// ---- S statistics on the OLD z (cal_lkd_props - cal_lkd, poly_geno.c:625): only genotypes
				//      whose four copies sit in one population see that population's table
				if (!a.init) {
					const uint32_t z0 = zo[j] & 0xFFu;
					if (zo[j] == z0 * 0x01010101u) {
						const int2 li = locsm[lj];
						const int code = ((ga[0] * li.x + ga[1]) * li.x + ga[2]) * li.x + ga[3];
						const int gid = a.c2i[li.y + code];
						const size_t o = ((size_t)(l0 + lj) * g.K + z0) * a.Gmax + gid;
						dsm[z0 * TETRA_THREADS + tid] += __ldg(a.tabP + o) - __ldg(a.tabC + o);
					}
				}
				// ---- new z: one categorical draw per copy, weights Q_ik P_k,l,geno_c (poly_geno.c:766-779)
				const uint32_t rr[4] = {rw[2 * j], __byte_perm(rw[2 * j], 0u, 0x1032), rw[2 * j + 1], __byte_perm(rw[2 * j + 1], 0u, 0x1032)};
				const float rowbf = as_dn(psm + (uint32_t)(lj * rowsz) * 4u);
				uint32_t packed = 0;
#pragma unroll
				for (int cidx = 0; cidx < 4; ++cidx) {
					const uint32_t row = __float_as_uint(fmaf(as_dn(ga[cidx]), (float)(KP * 4), rowbf));
					float p[KP], cw[KP];
#pragma unroll
					for (int v = 0; v < KP / 4; v++) {
						const float4 w = lds_f4(row + 16 * v);
						p[4 * v] = w.x; p[4 * v + 1] = w.y; p[4 * v + 2] = w.z; p[4 * v + 3] = w.w;
					}
					cw[0] = q[0] * p[0];
#pragma unroll
					for (int k = 1; k < KP; k++) cw[k] = fmaf(q[k], p[k], cw[k - 1]);
					const float zf = pick_category<KP>(cw, uniform_big16(rr[cidx], kc16));
					red_inc(__float_as_uint(fmaf(zf, as_dn(4u * TETRA_THREADS), cnt_t)));
					packed |= __float_as_uint(zf * as_dn(1u)) << (8 * cidx);
				}
				zn[j] = packed;
			}
			stg_stream(zp + (size_t)mt * Nloc, make_int4((int)zn[0], (int)zn[1], (int)zn[2], (int)zn[3]));
		}
		// ---- partials of this (chunk, individual)
		uint32_t *pc = reinterpret_cast<uint32_t *>(a.pcnt + ((size_t)chunk * Nloc + il) * KP);
#pragma unroll
		for (int j = 0; j < KP / 2; j++) {
			const int ca = cntsm[(2 * j) * TETRA_THREADS + tid], cb = cntsm[(2 * j + 1) * TETRA_THREADS + tid];
			cntsm[(2 * j) * TETRA_THREADS + tid] = 0;
			cntsm[(2 * j + 1) * TETRA_THREADS + tid] = 0;
			pc[j] = (uint32_t)ca | ((uint32_t)cb << 16);
		}
		}
		// ---- S statistics of this (chunk, 256 individuals): each thread's fp32 sum over the chunk's loci becomes a
		//      2^-24 fixed-point integer, and integers add exactly in any order -- so the K totals are identical for
		//      every grid shape and every shard count without keeping per-individual partials
		if (!a.init)
			for (int k = 0; k < g.K; k++) {
				long long fx = __float2ll_rn(dsm[k * TETRA_THREADS + tid] * 16777216.0f);
				dsm[k * TETRA_THREADS + tid] = 0.0f;
#pragma unroll
				for (int off = 16; off > 0; off >>= 1) fx += __shfl_down_sync(0xFFFFFFFFu, fx, off);
				if ((tid & 31) == 0 && fx != 0) atomicAdd(&dacc[k], (unsigned long long)fx);
			}
	}
	__syncthreads();
	if (!a.init && tid < g.K && dacc[tid] != 0ull) atomicAdd(a.dfix + tid, dacc[tid]);
}

// chunk-group width of the per-individual reductions below (tetra_q, tetra_indv_lkh)
constexpr int CG = 8;

// --------------------------------------------------------------------------------------
// Q_i ~ Dirichlet(cnt_i + alpha), poly_geno.c:812-833
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * CG) tetra_q_kernel(const uint16_t *pcnt, double *ind, float *Qf, int32_t *cnt_out, const DevScalars *sc, Geometry g,
                                                          uint32_t iter, uint32_t key0, uint32_t key1)
{
	__shared__ int shc[CG][MAX_K][32];
	__shared__ double gq[MAX_K][32];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int il = blockIdx.x * 32 + lane;
	const bool live = il < g.Nloc;
	int cnt[MAX_K];
#pragma unroll
	for (int k = 0; k < MAX_K; k++) cnt[k] = 0;
	if (live)
		for (int c = w; c < g.nchunks; c += CG) {
			const uint16_t *pc = pcnt + ((size_t)c * g.Nloc + il) * g.KP;
#pragma unroll
			for (int k = 0; k < MAX_K; k++)
				if (k < g.K) cnt[k] += pc[k];
		}
#pragma unroll
	for (int k = 0; k < MAX_K; k++) shc[w][k][lane] = cnt[k];
	__syncthreads();
	// warp k draws population k's gamma for the CTA's 32 individuals (as in indiv_epilogue_kernel), warp 0 normalises
	const int ig_global = g.i0 + il;
	if (live)
		for (int k = w; k < g.K; k += CG) {
			int ck = 0;
			for (int ww = 0; ww < CG; ww++) ck += shc[ww][k][lane];
			Stream sq((uint32_t)ig_global, (uint32_t)k, iter, TAG_Q, key0, key1);
			gq[k][lane] = draw_gamma(sq, (double)ck + sc->alpha);
		}
	__syncthreads();
	if (w != 0 || !live) return;
#pragma unroll
	for (int k = 0; k < MAX_K; k++) {
		int t = 0;
		for (int ww = 0; ww < CG; ww++) t += shc[ww][k][lane];
		cnt[k] = t;
	}
	double *rec = ind + (size_t)ig_global * g.REC;
	double qv[MAX_K], sum = 0.0;
#pragma unroll
	for (int k = 0; k < MAX_K; k++)
		if (k < g.K) { qv[k] = gq[k][lane]; sum += qv[k]; }
	double slq = 0.0;
#pragma unroll
	for (int k = 0; k < MAX_K; k++)
		if (k < g.K) {
			const double qk = qv[k] / sum;
			rec[k] = qk;
			slq += log(qk);
			Qf[(size_t)il * g.KP + k] = (float)qk;
			cnt_out[(size_t)il * g.K + k] = cnt[k];
		}
	for (int k = g.K; k < g.KP; k++) Qf[(size_t)il * g.KP + k] = 0.0f;
	rec[g.K + 1] = slq;
	rec[g.K + 2] = 0.0;
}

// --------------------------------------------------------------------------------------
// PASS B: dosage resolution | new z, Q, tables; likelihood; tally for the next update_P_auto
// --------------------------------------------------------------------------------------
struct GenoArgs {
	const int16_t *Xq; const int8_t *Zq; int8_t *Gq; const float *P; const float *Qf; const float *tab;
	const int32_t *loc_cat; const TetraCat *cats; const uint16_t *c2i;
	int32_t *n; float *lpart;
	Geometry geo; int Gmax; int init;       // init: uniform resolution (initial_geno), no likelihood, no tally
	uint32_t iter, key0, key1;
};

__device__ __forceinline__ float ex2_fast(float x)            // MUFU.EX2
{
	float r;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

// The reference evaluates the resolution weights and the likelihood in double (libm log / exp); here
// they are fp32 with MUFU.LG2 / MUFU.EX2: the weights carry ~1e-6 relative error (the z draw's fp32
// weights carry as much), the likelihood sums stay inside the 1e-6 gate (tests/test_gpu_tetra.py).
//
// The kernel is bound by instruction issue and a third of its instructions were integer address
// arithmetic (ncu source page, profiles/r1_tetra_geno_v4_source.txt), so all shared-memory accesses use
// explicit 32-bit shared addresses that advance by additions, the four (allele, population) bin
// indices of a genotype come from ONE multiply on the packed bytes (geno * KP + z), a resolution
// is a PRMT selector over the observed alleles and its catalogue code one dp4a.
constexpr int TETRA_R = 4;            // lane replicas of the tally histogram (tetra_configure)
__device__ __forceinline__ int2 lds_i2(uint32_t addr)
{
	int2 v;
	asm("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
	uint32_t v;
	asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
// Catalogue code of a genotype (its four alleles, one per byte, as a base-n number).  n <= 6: one dp4a with the weights
// (n^3, n^2, n, 1), which fit a byte each.  Beyond that (npack's top byte is 2 instead of 1): two dp4a with the weights
// (n, 1) on the upper and the lower pair and one multiply-add, hi * n^2 + lo.  A locus is the same for the whole warp, so
// the branch is uniform.
__device__ __forceinline__ uint32_t tcode(uint32_t gpk, uint32_t npack)
{
	if ((npack >> 24) == 1u) return __dp4a(gpk, npack, 0u);
	const uint32_t n = npack & 0xFFu, w = n | (1u << 8);
	return __dp4a(gpk, w, 0u) * (n * n) + __dp4a(gpk, w << 16, 0u);
}
__host__ __device__ __forceinline__ int tetra_npack(int n)
{
	return n <= 6 ? ((n * n * n) | ((n * n) << 8) | (n << 16) | (1 << 24)) : (n | (2 << 24));
}
// table entry of the genotype with catalogue code `code`: staged copy (shared addresses) or global
template <bool STAGE>
__device__ __forceinline__ float tab_at(uint32_t tab_sa, const float *tab_g, uint32_t c2i_sa, const uint16_t *c2i_g, uint32_t code)
{
	if (STAGE) return lds_f(tab_sa + lds_u16(c2i_sa + 2u * code) * 4u);
	return __ldg(tab_g + c2i_g[code]);
}

template <int KP, int ROUNDS, bool STAGE>
__global__ void __launch_bounds__(TETRA_THREADS, 3) tetra_geno_kernel(const GenoArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ __align__(8) unsigned long long bar;
	const Geometry &g = a.geo;
	constexpr int R = TETRA_R;
	const int tid = threadIdx.x, chunk = blockIdx.x;
	const int l0 = chunk * g.TL, nl = min(g.TL, g.Lpad - l0), nmt = nl / TT, rowsz = g.A * KP;
	float *Psm = reinterpret_cast<float *>(smem_raw);
	int *hist = reinterpret_cast<int *>(Psm + (size_t)g.TL * rowsz);            // [TL][A][KP][R]
	int2 *locsm = reinterpret_cast<int2 *>(hist + (size_t)g.TL * rowsz * R);    // [TL] ((n^3, n^2, n, 1) bytes, c2i offset)
	// the chunk's slice of the genotype-frequency tables [TL][K][Gmax] (one more bulk copy) and the code -> index
	// bytes: the same-population look-ups become shared-memory loads instead of dependent L1 / L2 round trips
	float *tabsm = reinterpret_cast<float *>(locsm + g.TL);
	uint8_t *c2ism = reinterpret_cast<uint8_t *>(tabsm + (size_t)g.TL * g.K * a.Gmax);
	const int nbins = nl * rowsz;
	const uint32_t tab_bytes = STAGE ? (uint32_t)nl * (uint32_t)(g.K * a.Gmax) * 4u : 0u;
	if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
	__syncthreads();
	if (tid == 0) {
		mbar_expect_tx(&bar, (uint32_t)nbins * 4u + tab_bytes);
		tma_bulk_g2s(Psm, a.P + (size_t)l0 * rowsz, (uint32_t)nbins * 4u, &bar);
		if (STAGE) tma_bulk_g2s(tabsm, a.tab + (size_t)l0 * g.K * a.Gmax, tab_bytes, &bar);
	}
	if (STAGE) for (int j = tid; j < g.c2i_bytes; j += TETRA_THREADS) c2ism[j] = reinterpret_cast<const uint8_t *>(a.c2i)[j];
	for (int j = tid; j < nbins * R; j += TETRA_THREADS) hist[j] = 0;
	for (int j = tid; j < nl; j += TETRA_THREADS) {
		const int ci = (l0 + j < g.L) ? a.loc_cat[l0 + j] : -1;
		const int n = ci >= 0 ? a.cats[ci].n : 1;                                // n <= 6: n^3 fits a byte
		locsm[j] = make_int2(tetra_npack(n), ci >= 0 ? a.cats[ci].c2i_off : 0);
	}
	__syncthreads();
	mbar_wait(&bar, 0);

	const int Nloc = g.Nloc, mt0 = l0 / TT;
	const int nsub_total = (Nloc + TETRA_THREADS - 1) / TETRA_THREADS;
	const int sub0 = blockIdx.y * g.subs_per_blk, sub1 = min(sub0 + g.subs_per_blk, nsub_total);
	const uint32_t psm_sa = smem_addr(Psm), loc_sa = smem_addr(locsm), tab_sa0 = smem_addr(tabsm), c2i_sa0 = smem_addr(c2ism);
	const uint32_t hist_sa = smem_addr(hist) + (uint32_t)(tid & (R - 1)) * 4u;
	const uint32_t row4 = (uint32_t)rowsz * 4u;                 // bytes of one locus of P; one locus of the histogram is R of them
	const uint32_t G4 = (uint32_t)a.Gmax * 4u, KG4 = (uint32_t)g.K * G4;
	const float LOG2E = 1.4426950408889634f;
	const float LG2_6 = 2.584962500721156f;

	for (int sub = sub0; sub < sub1; ++sub) {
		const int il = sub * TETRA_THREADS + tid;
		if (il >= Nloc) continue;
		float q[KP];
#pragma unroll
		for (int v = 0; v < KP / 4; v++) {
			const float4 w = __ldg(reinterpret_cast<const float4 *>(a.Qf + (size_t)il * KP) + v);
			q[4 * v] = w.x; q[4 * v + 1] = w.y; q[4 * v + 2] = w.z; q[4 * v + 3] = w.w;
		}
		const uint32_t ig_global = (uint32_t)(g.i0 + il);
		const int4 *xp = reinterpret_cast<const int4 *>(a.Xq) + ((size_t)mt0 * Nloc + il) * 2;
		const int4 *zp = reinterpret_cast<const int4 *>(a.Zq) + ((size_t)mt0 * Nloc + il);
		int4 *gp = reinterpret_cast<int4 *>(a.Gq) + ((size_t)mt0 * Nloc + il);
		float ll_nat = 0.0f, ll_lg2 = 0.0f;      // natural-log part (tables, multiplicities) and log2 part (allele frequencies)
		int4 xa_n = ldg_stream(xp), xb_n = ldg_stream(xp + 1), zv_n = ldg_stream(zp);      // prefetch one micro-tile ahead
		uint32_t p_sa = psm_sa, l_sa = loc_sa, t_sa = tab_sa0;                // running shared addresses of the micro-tile's first locus
		const float *t_g = a.tab + (size_t)l0 * g.K * a.Gmax;
		for (int mt = 0; mt < nmt; ++mt) {
			const int4 xa = xa_n, xb = xb_n, zv = zv_n;
			if (mt + 1 < nmt) {
				xa_n = ldg_stream(xp + (size_t)(mt + 1) * Nloc * 2); xb_n = ldg_stream(xp + (size_t)(mt + 1) * Nloc * 2 + 1);
				zv_n = ldg_stream(zp + (size_t)(mt + 1) * Nloc);
			}
			// one Philox block per (micro-tile, individual): word j resolves the dosage of locus j
			const u32x4 rnd4 = philox4x32<ROUNDS>(u32x4{(uint32_t)(mt0 + mt), ig_global, a.iter, TAG_GENO}, a.key0, a.key1);
			const uint32_t rj[4] = {rnd4.x, rnd4.y, rnd4.z, rnd4.w};
			const uint32_t xw[8] = {(uint32_t)xa.x, (uint32_t)xa.y, (uint32_t)xa.z, (uint32_t)xa.w, (uint32_t)xb.x, (uint32_t)xb.y, (uint32_t)xb.z, (uint32_t)xb.w};
			const uint32_t zw[4] = {(uint32_t)zv.x, (uint32_t)zv.y, (uint32_t)zv.z, (uint32_t)zv.w};
			uint32_t gn[4];
			float m_nat = 0.0f, m_lg2 = 0.0f;
#pragma unroll
			for (int j = 0; j < TT; ++j) {
				gn[j] = 0xFFFFFFFFu;
				// distinct alleles ascending in int16, -1 padding; all -1 = missing (data_interface.c:636-650)
				if (xw[2 * j] & 0x8000u) continue;
				const int nd = 1 + ((xw[2 * j] >> 31) ^ 1) + (((xw[2 * j + 1] >> 15) & 1) ^ 1) + ((xw[2 * j + 1] >> 31) ^ 1);
				const uint32_t apack = __byte_perm(xw[2 * j], xw[2 * j + 1], 0x6420);   // the observed alleles, one per byte
				const int2 li = lds_i2(l_sa + 8u * j);
				const uint32_t npack = (uint32_t)li.x;                            // dp4a(genotype, npack) = catalogue code
				const uint32_t pj_sa = p_sa + (uint32_t)j * row4;                 // &P[l][0][0]
				const uint32_t z0 = zw[j] & 0xFFu;
				const bool same = (zw[j] == z0 * 0x01010101u);
				const uint32_t tj_sa = t_sa + (uint32_t)j * KG4 + z0 * G4;        // population z0's table of this locus
				const float *tj_g = t_g + ((size_t)j * g.K + z0) * a.Gmax;
				const uint32_t cj_sa = c2i_sa0 + 2u * (uint32_t)li.y;
				const uint16_t *cj_g = a.c2i + li.y;
				uint32_t gpk;
				float lmul;                                                        // heterozygote multiplicities log 4, 6, 12, 24 (poly_geno.c:1262-1268)
				if (nd == 1) { gpk = (apack & 0xFFu) * 0x01010101u; lmul = 0.0f; }
				else if (nd == 4) { gpk = apack; lmul = 3.1780538303479458f; }
				else {
					// ---- three dosage resolutions (choose_two_auto :854, choose_tri_auto :907), weights in log2; as PRMT
					//      selectors over the observed alleles (two_allele_auto :2440 / tri_allele_auto :2509):
					//      nd 2: a0a0a0a1 | a1a1a1a0 | a0a0a1a1     nd 3: a0a0a1a2 | a1a1a0a2 | a2a2a0a1
					const uint32_t s0 = (nd == 2) ? 0x1000u : 0x2100u, s1 = (nd == 2) ? 0x0111u : 0x2011u, s2 = (nd == 2) ? 0x1100u : 0x1022u;
					float w0, w1, w2;
					if (a.init) { w0 = w1 = w2 = 0.0f; }                                // choose_unif, poly_geno.c:842
					else if (same) {
						w0 = tab_at<STAGE>(tj_sa, tj_g, cj_sa, cj_g, tcode(__byte_perm(apack, 0u, s0), npack)) * LOG2E;
						w1 = tab_at<STAGE>(tj_sa, tj_g, cj_sa, cj_g, tcode(__byte_perm(apack, 0u, s1), npack)) * LOG2E;
						w2 = tab_at<STAGE>(tj_sa, tj_g, cj_sa, cj_g, tcode(__byte_perm(apack, 0u, s2), npack)) * LOG2E;
					} else {
						const uint32_t r0 = pj_sa + (apack & 0xFFu) * (KP * 4u), r1 = pj_sa + ((apack >> 8) & 0xFFu) * (KP * 4u);
						const uint32_t r2 = (nd == 3) ? pj_sa + ((apack >> 16) & 0xFFu) * (KP * 4u) : r0;
						float f0 = 0.0f, f1 = 0.0f, f2 = 0.0f;
#pragma unroll
						for (int v = 0; v < KP / 4; v++) {
							const float4 t0 = lds_f4(r0 + 16 * v), t1 = lds_f4(r1 + 16 * v), t2 = lds_f4(r2 + 16 * v);
							f0 = fmaf(q[4 * v], t0.x, f0); f0 = fmaf(q[4 * v + 1], t0.y, f0); f0 = fmaf(q[4 * v + 2], t0.z, f0); f0 = fmaf(q[4 * v + 3], t0.w, f0);
							f1 = fmaf(q[4 * v], t1.x, f1); f1 = fmaf(q[4 * v + 1], t1.y, f1); f1 = fmaf(q[4 * v + 2], t1.z, f1); f1 = fmaf(q[4 * v + 3], t1.w, f1);
							f2 = fmaf(q[4 * v], t2.x, f2); f2 = fmaf(q[4 * v + 1], t2.y, f2); f2 = fmaf(q[4 * v + 2], t2.z, f2); f2 = fmaf(q[4 * v + 3], t2.w, f2);
						}
						const float l0f = lg2_fast(f0), l1f = lg2_fast(f1), l2f = lg2_fast(f2);
						if (nd == 2) { w0 = 2.0f + 3.0f * l0f + l1f; w1 = 2.0f + 3.0f * l1f + l0f; w2 = LG2_6 + 2.0f * l0f + 2.0f * l1f; }
						else { w0 = 2.0f * l0f + l1f + l2f; w1 = 2.0f * l1f + l0f + l2f; w2 = 2.0f * l2f + l1f + l0f; }
					}
					const float e1 = ex2_fast(w1 - w0), e2 = ex2_fast(w2 - w0);          // e0 = 1
					const float c1w = 1.0f + e1, c2w = c1w + e2;
					const float u = u01f(rj[j]) * c2w;
					const uint32_t sel = (u < 1.0f) ? s0 : (u < c1w ? s1 : s2);
					gpk = __byte_perm(apack, 0u, sel);
					lmul = (nd == 3) ? 2.4849066497880004f : ((u < c1w) ? 1.3862943611198906f : 1.791759469228055f);
				}
				gn[j] = gpk;
				if (a.init) continue;
				// the four (allele, population) bins of the genotype, one per byte: geno * KP + z (< 256, no carries)
				const uint32_t ipk = gpk * (uint32_t)KP + zw[j];
				const uint32_t i0 = ipk & 0xFFu, i1 = (ipk >> 8) & 0xFFu, i2 = (ipk >> 16) & 0xFFu, i3 = ipk >> 24;
				// ---- likelihood of the result (calc_genofq, poly_geno.c:1235-1286)
				if (same) m_nat += tab_at<STAGE>(tj_sa, tj_g, cj_sa, cj_g, tcode(gpk, npack));
				else {
					m_nat += lmul;
					m_lg2 += lg2_fast(lds_f(pj_sa + i0 * 4u) * lds_f(pj_sa + i1 * 4u)) + lg2_fast(lds_f(pj_sa + i2 * 4u) * lds_f(pj_sa + i3 * 4u));
				}
				// ---- tally of the next update_P_auto over the latent genotype (poly_geno.c:403-424): shared-space RED
				const uint32_t hj_sa = hist_sa + (pj_sa - psm_sa) * (uint32_t)R;
				red_inc(hj_sa + i0 * (4u * R));
				red_inc(hj_sa + i1 * (4u * R));
				red_inc(hj_sa + i2 * (4u * R));
				red_inc(hj_sa + i3 * (4u * R));
			}
			*(gp + (size_t)mt * Nloc) = make_int4((int)gn[0], (int)gn[1], (int)gn[2], (int)gn[3]);
			ll_nat += m_nat; ll_lg2 += m_lg2;
			p_sa += TT * row4; l_sa += TT * 8u; t_sa += TT * KG4; t_g += (size_t)TT * g.K * a.Gmax;
		}
		if (!a.init) {
			a.lpart[((size_t)chunk * 2) * Nloc + il] = ll_nat;
			a.lpart[((size_t)chunk * 2 + 1) * Nloc + il] = ll_lg2;
		}
	}
	__syncthreads();
	if (a.init) return;
	int32_t *ng = a.n + (size_t)l0 * rowsz;
	for (int b = tid; b < nbins; b += TETRA_THREADS) {
		int s = 0;
		for (int r = 0; r < R; r++) s += hist[b * R + r];
		if (s) atomicAdd(ng + b, s);
	}
}

// --------------------------------------------------------------------------------------
// PASS B, allotetraploid (-ap 0): update_geno :524-548 with choose_two_allo :962 (7 resolutions),
// choose_tri_allo :1043 (12), choose_tetra_allo :1144 (6); calc_genofq :1267-1283; tally of
// update_P_allo :441-489 (copies 0,1 -> n, copies 2,3 -> n2).  Same thread mapping and staging as
// the autotetraploid kernel.  A resolution is a PRMT selector: nibble c = index into the observed
// allele set of copy c, so that byte_perm(observed alleles, selector) IS the resolved genotype, and
// dp4a(genotype, (n^3, n^2, n, 1)) its catalogue code.  The three lists are compile-time constants
// and each observed-allele count runs its own fully unrolled path: no table in memory, no
// dynamically indexed array.
// --------------------------------------------------------------------------------------
#define ASEL(a, b, c, d) ((uint32_t)((a) | ((b) << 4) | ((c) << 8) | ((d) << 12)))
template <int ND> struct AlloRes { static constexpr int N = (ND == 2) ? 7 : (ND == 3 ? 12 : 6); };
template <int ND>
__host__ __device__ constexpr uint32_t allo_sel(int r)
{
	// two_allele_allo, poly_geno.c:2465 | tri_allele_allo :2533 | tetra_allele_allo :2602
	constexpr uint32_t S2[7] = {ASEL(0,0,0,1), ASEL(0,1,0,0), ASEL(0,0,1,1), ASEL(1,1,0,0), ASEL(0,1,1,1), ASEL(1,1,0,1), ASEL(0,1,0,1)};
	constexpr uint32_t S3[12] = {ASEL(0,0,1,2), ASEL(1,2,0,0), ASEL(1,1,0,2), ASEL(0,2,1,1), ASEL(2,2,0,1), ASEL(0,1,2,2),
	                             ASEL(0,1,1,2), ASEL(1,2,0,1), ASEL(1,2,0,2), ASEL(0,2,1,2), ASEL(0,2,0,1), ASEL(0,1,0,2)};
	constexpr uint32_t S4[6] = {ASEL(0,1,2,3), ASEL(2,3,0,1), ASEL(0,2,1,3), ASEL(1,3,0,2), ASEL(0,3,1,2), ASEL(1,2,0,3)};
	return ND == 2 ? S2[r < 7 ? r : 0] : (ND == 3 ? S3[r < 12 ? r : 0] : S4[r < 6 ? r : 0]);
}
#undef ASEL

// One genotype with ND observed alleles (bytes of apack, ascending): draws the resolution, returns the
// resolved genotype g0 | g1 << 8 | g2 << 16 | g3 << 24.
template <int ND, int KP, bool STAGE>
__device__ __forceinline__ uint32_t allo_resolve(uint32_t apack, uint32_t npack, int init, bool same, uint32_t tab_sa, const float *tab_g,
                                                 uint32_t c2i_sa, const uint16_t *c2i_g, uint32_t p1_sa, uint32_t p2_sa, const float (&q)[KP], float u01)
{
	using RS = AlloRes<ND>;
	constexpr int NR = RS::N;
	const float LOG2E = 1.4426950408889634f;
	float w[NR];
	if (init) {                                                        // choose_unif, poly_geno.c:842
#pragma unroll
		for (int r = 0; r < NR; r++) w[r] = 0.0f;
	} else if (same) {                                                 // population z's table at the resolution's genotype
#pragma unroll
		for (int r = 0; r < NR; r++)
			w[r] = tab_at<STAGE>(tab_sa, tab_g, c2i_sa, c2i_g, tcode(__byte_perm(apack, 0u, allo_sel<ND>(r)), npack)) * LOG2E;
	} else {                                                           // admixture-averaged frequencies of the two subgenomes
		float lf[ND], lf2[ND];
#pragma unroll
		for (int t = 0; t < ND; t++) {
			const uint32_t off = ((apack >> (8 * t)) & 0xFFu) * (KP * 4u);
			float f = 0.0f, f2 = 0.0f;
#pragma unroll
			for (int v = 0; v < KP / 4; v++) {
				const float4 t1 = lds_f4(p1_sa + off + 16 * v), t2 = lds_f4(p2_sa + off + 16 * v);
				f = fmaf(q[4 * v], t1.x, f); f = fmaf(q[4 * v + 1], t1.y, f); f = fmaf(q[4 * v + 2], t1.z, f); f = fmaf(q[4 * v + 3], t1.w, f);
				f2 = fmaf(q[4 * v], t2.x, f2); f2 = fmaf(q[4 * v + 1], t2.y, f2); f2 = fmaf(q[4 * v + 2], t2.z, f2); f2 = fmaf(q[4 * v + 3], t2.w, f2);
			}
			lf[t] = lg2_fast(f); lf2[t] = lg2_fast(f2);
		}
#pragma unroll
		for (int r = 0; r < NR; r++) {
			const int i0 = allo_sel<ND>(r) & 3, i1 = (allo_sel<ND>(r) >> 4) & 3, i2 = (allo_sel<ND>(r) >> 8) & 3, i3 = (allo_sel<ND>(r) >> 12) & 3;
			// the reference adds log 2 only where BOTH pairs are heterozygous, and never with four alleles
			const bool both_het = (ND != 4) && (i0 != i1) && (i2 != i3);
			w[r] = lf[i0] + lf[i1] + lf2[i2] + lf2[i3] + (both_het ? 1.0f : 0.0f);
		}
	}
	float cum[NR];
	cum[0] = 1.0f;
#pragma unroll
	for (int r = 1; r < NR; r++) cum[r] = cum[r - 1] + ex2_fast(w[r] - w[0]);
	const float u = u01 * cum[NR - 1];
	uint32_t sel = allo_sel<ND>(0);
#pragma unroll
	for (int r = 1; r < NR; r++) sel = (u >= cum[r - 1]) ? allo_sel<ND>(r) : sel;
	return __byte_perm(apack, 0u, sel);
}

struct GenoAlloArgs {
	const int16_t *Xq; const int8_t *Zq; int8_t *Gq; const float *P, *P2; const float *Qf; const float *tab;
	const int32_t *loc_cat; const TetraCat *cats; const uint16_t *c2i;
	int32_t *n, *n2; float *lpart;
	Geometry geo; int Gmax; int init;
	uint32_t iter, key0, key1;
};

template <int KP, int ROUNDS, bool STAGE>
__global__ void __launch_bounds__(TETRA_THREADS, 2) tetra_geno_allo_kernel(const GenoAlloArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ __align__(8) unsigned long long bar;
	const Geometry &g = a.geo;
	constexpr int R = TETRA_R;
	const int tid = threadIdx.x, chunk = blockIdx.x;
	const int l0 = chunk * g.TL, nl = min(g.TL, g.Lpad - l0), nmt = nl / TT, rowsz = g.A * KP;
	float *Psm = reinterpret_cast<float *>(smem_raw);
	float *P2sm = Psm + (size_t)g.TL * rowsz;
	int *hist = reinterpret_cast<int *>(P2sm + (size_t)g.TL * rowsz);          // [2][TL][A][KP][R]
	int2 *locsm = reinterpret_cast<int2 *>(hist + (size_t)2 * g.TL * rowsz * R); // [TL] ((n^3, n^2, n, 1) bytes, c2i offset)
	float *tabsm = reinterpret_cast<float *>(locsm + g.TL);                       // staged tables, as in tetra_geno_kernel
	uint8_t *c2ism = reinterpret_cast<uint8_t *>(tabsm + (size_t)g.TL * g.K * a.Gmax);
	const int nbins = nl * rowsz;
	const uint32_t tab_bytes = STAGE ? (uint32_t)nl * (uint32_t)(g.K * a.Gmax) * 4u : 0u;
	if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
	__syncthreads();
	if (tid == 0) {
		mbar_expect_tx(&bar, (uint32_t)nbins * 8u + tab_bytes);
		tma_bulk_g2s(Psm, a.P + (size_t)l0 * rowsz, (uint32_t)nbins * 4u, &bar);
		tma_bulk_g2s(P2sm, a.P2 + (size_t)l0 * rowsz, (uint32_t)nbins * 4u, &bar);
		if (STAGE) tma_bulk_g2s(tabsm, a.tab + (size_t)l0 * g.K * a.Gmax, tab_bytes, &bar);
	}
	if (STAGE) for (int j = tid; j < g.c2i_bytes; j += TETRA_THREADS) c2ism[j] = reinterpret_cast<const uint8_t *>(a.c2i)[j];
	for (int j = tid; j < 2 * g.TL * rowsz * R; j += TETRA_THREADS) hist[j] = 0;
	for (int j = tid; j < nl; j += TETRA_THREADS) {
		const int ci = (l0 + j < g.L) ? a.loc_cat[l0 + j] : -1;
		const int n = ci >= 0 ? a.cats[ci].n : 1;                              // tetra_npack: one dp4a up to n = 6, two beyond
		locsm[j] = make_int2(tetra_npack(n), ci >= 0 ? a.cats[ci].c2i_off : 0);
	}
	__syncthreads();
	mbar_wait(&bar, 0);

	const int Nloc = g.Nloc, mt0 = l0 / TT;
	const int nsub_total = (Nloc + TETRA_THREADS - 1) / TETRA_THREADS;
	const int sub0 = blockIdx.y * g.subs_per_blk, sub1 = min(sub0 + g.subs_per_blk, nsub_total);
	const uint32_t psm_sa = smem_addr(Psm), loc_sa = smem_addr(locsm), tab_sa0 = smem_addr(tabsm), c2i_sa0 = smem_addr(c2ism);
	const uint32_t p2off = (uint32_t)(g.TL * rowsz) * 4u;             // P2 chunk sits right behind the P chunk
	const uint32_t hist1_sa = smem_addr(hist) + (uint32_t)(tid & (R - 1)) * 4u;
	const uint32_t hist2off = (uint32_t)(g.TL * rowsz) * (4u * R);
	const uint32_t row4 = (uint32_t)rowsz * 4u;
	const uint32_t G4 = (uint32_t)a.Gmax * 4u, KG4 = (uint32_t)g.K * G4;

	for (int sub = sub0; sub < sub1; ++sub) {
		const int il = sub * TETRA_THREADS + tid;
		if (il >= Nloc) continue;
		float q[KP];
#pragma unroll
		for (int v = 0; v < KP / 4; v++) {
			const float4 w = __ldg(reinterpret_cast<const float4 *>(a.Qf + (size_t)il * KP) + v);
			q[4 * v] = w.x; q[4 * v + 1] = w.y; q[4 * v + 2] = w.z; q[4 * v + 3] = w.w;
		}
		const uint32_t ig_global = (uint32_t)(g.i0 + il);
		const int4 *xp = reinterpret_cast<const int4 *>(a.Xq) + ((size_t)mt0 * Nloc + il) * 2;
		const int4 *zp = reinterpret_cast<const int4 *>(a.Zq) + ((size_t)mt0 * Nloc + il);
		int4 *gp = reinterpret_cast<int4 *>(a.Gq) + ((size_t)mt0 * Nloc + il);
		float ll_nat = 0.0f, ll_lg2 = 0.0f;
		int4 xa_n = ldg_stream(xp), xb_n = ldg_stream(xp + 1), zv_n = ldg_stream(zp);      // prefetch one micro-tile ahead
		uint32_t p_sa = psm_sa, l_sa = loc_sa, t_sa = tab_sa0;                // running shared addresses of the micro-tile's first locus
		const float *t_g = a.tab + (size_t)l0 * g.K * a.Gmax;
		for (int mt = 0; mt < nmt; ++mt) {
			const int4 xa = xa_n, xb = xb_n, zv = zv_n;
			if (mt + 1 < nmt) {
				xa_n = ldg_stream(xp + (size_t)(mt + 1) * Nloc * 2); xb_n = ldg_stream(xp + (size_t)(mt + 1) * Nloc * 2 + 1);
				zv_n = ldg_stream(zp + (size_t)(mt + 1) * Nloc);
			}
			const u32x4 rnd4 = philox4x32<ROUNDS>(u32x4{(uint32_t)(mt0 + mt), ig_global, a.iter, TAG_GENO}, a.key0, a.key1);
			const uint32_t rj[4] = {rnd4.x, rnd4.y, rnd4.z, rnd4.w};
			const uint32_t xw[8] = {(uint32_t)xa.x, (uint32_t)xa.y, (uint32_t)xa.z, (uint32_t)xa.w, (uint32_t)xb.x, (uint32_t)xb.y, (uint32_t)xb.z, (uint32_t)xb.w};
			const uint32_t zw[4] = {(uint32_t)zv.x, (uint32_t)zv.y, (uint32_t)zv.z, (uint32_t)zv.w};
			uint32_t gn[4];
			float m_nat = 0.0f, m_lg2 = 0.0f;
#pragma unroll
			for (int j = 0; j < TT; ++j) {
				gn[j] = 0xFFFFFFFFu;
				// distinct alleles ascending in int16, -1 padding; all -1 = missing (data_interface.c:636-650)
				if (xw[2 * j] & 0x8000u) continue;
				const int nd = 1 + ((xw[2 * j] >> 31) ^ 1) + (((xw[2 * j + 1] >> 15) & 1) ^ 1) + ((xw[2 * j + 1] >> 31) ^ 1);
				const uint32_t apack = __byte_perm(xw[2 * j], xw[2 * j + 1], 0x6420);     // low bytes of the four int16
				const int2 li = lds_i2(l_sa + 8u * j);
				const uint32_t npack = (uint32_t)li.x;
				const uint32_t p1_sa = p_sa + (uint32_t)j * row4, p2_sa = p1_sa + p2off;    // &P[l][0][0], &P2[l][0][0]
				const uint32_t z0 = zw[j] & 0xFFu;
				const bool same = (zw[j] == z0 * 0x01010101u);
				const uint32_t tj_sa = t_sa + (uint32_t)j * KG4 + z0 * G4;
				const float *tj_g = t_g + ((size_t)j * g.K + z0) * a.Gmax;
				const uint32_t cj_sa = c2i_sa0 + 2u * (uint32_t)li.y;
				const uint16_t *cj_g = a.c2i + li.y;
				const float u01 = u01f(rj[j]);
				uint32_t gpk;
				if (nd == 1) gpk = (apack & 0xFFu) * 0x01010101u;
				else if (nd == 2) gpk = allo_resolve<2, KP, STAGE>(apack, npack, a.init, same, tj_sa, tj_g, cj_sa, cj_g, p1_sa, p2_sa, q, u01);
				else if (nd == 3) gpk = allo_resolve<3, KP, STAGE>(apack, npack, a.init, same, tj_sa, tj_g, cj_sa, cj_g, p1_sa, p2_sa, q, u01);
				else gpk = allo_resolve<4, KP, STAGE>(apack, npack, a.init, same, tj_sa, tj_g, cj_sa, cj_g, p1_sa, p2_sa, q, u01);
				gn[j] = gpk;
				if (a.init) continue;
				// the four (allele, population) bins, one per byte: geno * KP + z (< 256, no carries)
				const uint32_t ipk = gpk * (uint32_t)KP + zw[j];
				const uint32_t i0 = ipk & 0xFFu, i1 = (ipk >> 8) & 0xFFu, i2 = (ipk >> 16) & 0xFFu, i3 = ipk >> 24;
				// ---- likelihood of the result (calc_genofq, poly_geno.c:1235-1286, allotetraploid branch)
				if (same) m_nat += tab_at<STAGE>(tj_sa, tj_g, cj_sa, cj_g, tcode(gpk, npack));
				else {
					// classes 1,2: log 2; class 3: log 4 -- one log 2 per heterozygous pair
					const uint32_t hx = gpk ^ (gpk >> 8);                             // byte 0: g0 ^ g1, byte 2: g2 ^ g3
					m_nat += (float)(((hx & 0xFFu) ? 1 : 0) + ((hx & 0xFF0000u) ? 1 : 0)) * 0.6931471805599453f;
					m_lg2 += lg2_fast(lds_f(p1_sa + i0 * 4u) * lds_f(p1_sa + i1 * 4u)) + lg2_fast(lds_f(p2_sa + i2 * 4u) * lds_f(p2_sa + i3 * 4u));
				}
				const uint32_t h1_sa = hist1_sa + (p1_sa - psm_sa) * (uint32_t)R, h2_sa = h1_sa + hist2off;
				red_inc(h1_sa + i0 * (4u * R));
				red_inc(h1_sa + i1 * (4u * R));
				red_inc(h2_sa + i2 * (4u * R));
				red_inc(h2_sa + i3 * (4u * R));
			}
			*(gp + (size_t)mt * Nloc) = make_int4((int)gn[0], (int)gn[1], (int)gn[2], (int)gn[3]);
			ll_nat += m_nat; ll_lg2 += m_lg2;
			p_sa += TT * row4; l_sa += TT * 8u; t_sa += TT * KG4; t_g += (size_t)TT * g.K * a.Gmax;
		}
		if (!a.init) {
			a.lpart[((size_t)chunk * 2) * Nloc + il] = ll_nat;
			a.lpart[((size_t)chunk * 2 + 1) * Nloc + il] = ll_lg2;
		}
	}
	__syncthreads();
	if (a.init) return;
	for (int h2 = 0; h2 < 2; h2++) {
		int32_t *ng = (h2 ? a.n2 : a.n) + (size_t)l0 * rowsz;
		const int *hh = hist + (size_t)h2 * g.TL * rowsz * R;
		for (int b = tid; b < nbins; b += TETRA_THREADS) {
			int sacc = 0;
			for (int r = 0; r < R; r++) sacc += hh[b * R + r];
			if (sacc) atomicAdd(ng + b, sacc);
		}
	}
}

// indvlkh (cal_lkd :715): chunk partials in chunk order, one thread per individual
__global__ void __launch_bounds__(32 * CG) tetra_indv_lkh_kernel(const float *lpart, double *ind, Geometry g)
{
	__shared__ double sh[CG][32];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int il = blockIdx.x * 32 + lane;
	const bool live = il < g.Nloc;
	double s = 0.0;
	if (live)
		for (int c = w; c < g.nchunks; c += CG)
			s += (double)lpart[((size_t)c * 2) * g.Nloc + il] + (double)lpart[((size_t)c * 2 + 1) * g.Nloc + il] * LN2_D;
	sh[w][lane] = s;
	__syncthreads();
	if (w != 0 || !live) return;
	double t = 0.0;
	for (int ww = 0; ww < CG; ww++) t += sh[ww][lane];
	ind[(size_t)(g.i0 + il) * g.REC + g.K] = t;
}
// totallkh and the column sums of Q (check_empty_cluster mcmc.c:1944): one CTA, fixed tree
__global__ void __launch_bounds__(RED1) tetra_lkh_kernel(const double *ind, DevScalars *sc, Geometry g)
{
	__shared__ double sh[RED1];
	double tot = 0.0, qc[MAX_K];
	for (int k = 0; k < MAX_K; k++) qc[k] = 0.0;
	for (int i = threadIdx.x; i < g.N; i += RED1) {
		const double *rec = ind + (size_t)i * g.REC;
		tot += rec[g.K];
		for (int k = 0; k < g.K; k++) qc[k] += rec[k];
	}
	const double T = block_sum1(tot, sh);
	if (threadIdx.x == 0) sc->totallkh = T;
	for (int k = 0; k < g.K; k++) {
		const double v = block_sum1(qc[k], sh);
		if (threadIdx.x == 0) sc->qcol[k] = v;
	}
}

// stand-alone tally over (z, geno), for state injection and the start of a chain
// One CTA per micro-tile of TT loci: a shared histogram [TT][A][KP] with one replica per lane
// (the threads of a warp share their loci, so un-replicated atomics would all collide), flushed
// to n with one atomicAdd per non-empty bin.
constexpr int TALLY_REP = 16;
__global__ void __launch_bounds__(256) tetra_tally_kernel(const int8_t *Zq, const int8_t *Gq, int32_t *n, int32_t *n2, Geometry g, int LTq)
{
	extern __shared__ int th[];                                    // [TT * A * KP][TALLY_REP] (twice for the allotetraploid model)
	const int mt = blockIdx.x, tid = threadIdx.x;
	const int nb = TT * g.A * g.KP;
	const int nsub = n2 ? 2 : 1;
	for (int b = tid; b < nsub * nb * TALLY_REP; b += blockDim.x) th[b] = 0;
	__syncthreads();
	const int rep = tid & (TALLY_REP - 1);
	const int4 *zp = reinterpret_cast<const int4 *>(Zq) + (size_t)mt * g.Nloc;
	const int4 *gp = reinterpret_cast<const int4 *>(Gq) + (size_t)mt * g.Nloc;
	for (int il = tid; il < g.Nloc; il += blockDim.x) {
		const int4 zv = zp[il], gv = gp[il];
		const uint32_t zw[4] = {(uint32_t)zv.x, (uint32_t)zv.y, (uint32_t)zv.z, (uint32_t)zv.w};
		const uint32_t gw[4] = {(uint32_t)gv.x, (uint32_t)gv.y, (uint32_t)gv.z, (uint32_t)gv.w};
#pragma unroll
		for (int j = 0; j < TT; j++) {
			if ((int8_t)(gw[j] & 0xFF) < 0) continue;              // missing genotype
#pragma unroll
			for (int cpy = 0; cpy < 4; cpy++) {
				const int a = (gw[j] >> (8 * cpy)) & 0xFF, z = (zw[j] >> (8 * cpy)) & 0xFF;
				atomicAdd(&th[(((n2 && cpy >= 2) ? nb : 0) + (j * g.A + a) * g.KP + z) * TALLY_REP + rep], 1);   // update_P_allo: copies 2,3 apart
			}
		}
	}
	__syncthreads();
	for (int b = tid; b < nb; b += blockDim.x) {
		int sacc = 0;
		for (int r = 0; r < TALLY_REP; r++) sacc += th[b * TALLY_REP + r];
		if (sacc) atomicAdd(&n[(size_t)mt * TT * g.A * g.KP + b], sacc);
	}
	if (n2)
		for (int b = tid; b < nb; b += blockDim.x) {
			int sacc = 0;
			for (int r = 0; r < TALLY_REP; r++) sacc += th[(nb + b) * TALLY_REP + r];
			if (sacc) atomicAdd(&n2[(size_t)mt * TT * g.A * g.KP + b], sacc);
		}
}
static cudaError_t launch_tetra_tally(const int8_t *Zq, const int8_t *Gq, int32_t *n, int32_t *n2, const Geometry &g, int LTq, cudaStream_t s)
{
	const size_t sm = (size_t)TT * g.A * g.KP * TALLY_REP * sizeof(int) * (n2 ? 2 : 1);
	cudaError_t e = cudaFuncSetAttribute(tetra_tally_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
	if (e != cudaSuccess) return e;
	tetra_tally_kernel<<<LTq, 256, sm, s>>>(Zq, Gq, n, n2, g, LTq);
	return cudaGetLastError();
}

// geno = -1 wherever the genotype is missing (the sweep kernels read the mask from there)
__global__ void tetra_mask_geno_kernel(const int16_t *Xq, int8_t *Gq, size_t ngeno)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= ngeno) return;
	if (Xq[t * 4] < 0) for (int c = 0; c < 4; c++) Gq[t * 4 + c] = -1;
}

__global__ void tetra_init_scalars_kernel(DevScalars *sc, double *S, const float *initd, int K, uint32_t key0, uint32_t key1)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	Stream st(0u, 1u, 0u, TAG_INIT, key0, key1);
	sc->alpha = st.uniform() * 10;                       // initial_chn, poly_geno.c:386
	sc->totallkh = 0.0; sc->sumlogq = 0.0; sc->cur_prop_ll = 0.0;
	sc->alpha_accepts = 0; sc->s_accepts = 0; sc->flags = 0;
	for (int k = 0; k < K; k++) S[k] = (double)initd[k];
}

// --------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------
static cudaError_t tetra_configure(Geometry &g, int device, bool allo, int Gmax, int c2i_bytes)
{
	int sms = 148, smem_optin = 227 * 1024;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
	cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
	const int target = 4 * sms;
	const size_t per_locus = (size_t)g.A * g.KP * 4;
	const int R = TETRA_R;
	const size_t fixedA = (size_t)2 * g.KP * TETRA_THREADS * 4 + 2048;
	const size_t budget = (size_t)smem_optin / 2 - 1024;
	int tl = (int)((budget - fixedA) / (per_locus * (1 + R) * (allo ? 2 : 1) + 8));     // allotetraploid: two P chunks, two histograms
	tl = tl / TT * TT;
	if (tl < TT) return cudaErrorInvalidConfiguration;
	if (tl > 256) tl = 256;
	int want = ((g.Lpad + target - 1) / target + TT - 1) / TT * TT;
	if (want < TT) want = TT;
	if (want < tl) tl = want;
	// PASS B stages its chunk of the tables [TL][K][Gmax] and the code -> index bytes in shared memory when that
	// leaves the chunk at (nearly) the length the grid wants; it runs 3 resident CTAs per SM (allotetraploid: 2)
	g.tab_stage = 0;
	g.c2i_bytes = c2i_bytes;
	{
		const size_t per_b = per_locus * (1 + R) * (allo ? 2 : 1) + 8 + (size_t)g.K * Gmax * 4;
		const size_t bud_b = (size_t)smem_optin / (allo ? 2 : 3) - 1024 - (size_t)((c2i_bytes + 15) / 16 * 16);
		int tl_s = (int)(bud_b / per_b) / TT * TT;
		if (tl_s >= TT && 4 * tl_s >= 3 * tl) { g.tab_stage = 1; if (tl_s < tl) tl = tl_s; }
	}
	g.TL = tl;
	g.nchunks = (g.Lpad + tl - 1) / tl;
	const int nsub_total = (g.Nloc + TETRA_THREADS - 1) / TETRA_THREADS;
	// about 15 waves of CTAs (3 resident per SM, allotetraploid 2): at config 5 (556 chunks, 79 passes) 2 / 4 / 6 / 12
	// individual blocks give 6.85 / 6.57 / 6.35 / 6.29 ms per sweep
	const int want_ctas = 15 * (allo ? 2 : 3) * sms;
	int nblk = std::max(1, std::min(nsub_total, (want_ctas + g.nchunks - 1) / g.nchunks));
	if (g.nchunks >= (allo ? 2 : 3) * sms && nsub_total >= 2) nblk = std::min(nblk, nsub_total / 2);     // as in zq_configure
	if (const char *e = getenv("IG_TETRA_NBLK")) nblk = std::max(1, std::min(nsub_total, atoi(e)));       // tuning hook
	g.subs_per_blk = (nsub_total + nblk - 1) / nblk;
	g.nblk = (nsub_total + g.subs_per_blk - 1) / g.subs_per_blk;
	g.R = R;
	g.zq_smem = 0;
	return cudaSuccess;
}
static size_t smem_zs(const Geometry &g) { return (size_t)g.TL * g.A * g.KP * 4 + (size_t)2 * g.KP * TETRA_THREADS * 4 + (size_t)g.TL * 8; }
static size_t smem_tab(const Geometry &g, int Gmax) { return g.tab_stage ? (size_t)g.TL * g.K * Gmax * 4 + (size_t)((g.c2i_bytes + 15) / 16 * 16) : 0; }
static size_t smem_geno(const Geometry &g, int Gmax) { return (size_t)g.TL * g.A * g.KP * 4 * (1 + g.R) + (size_t)g.TL * 8 + smem_tab(g, Gmax); }
static size_t smem_geno_allo(const Geometry &g, int Gmax) { return (size_t)2 * g.TL * g.A * g.KP * 4 * (1 + g.R) + (size_t)g.TL * 8 + smem_tab(g, Gmax); }

}  // namespace ig

using namespace ig;

ig_status tetra_create(ig_ctx *c)
{
	if (c->cfg.autopoly != 1 && c->cfg.autopoly != 0) return fail(IG_ERR_ARG, "ploid 4: autopoly must be 1 (autotetraploid) or 0 (allotetraploid)");
	if (c->cfg.back_refl != 1) return fail(IG_ERR_UNSUPPORTED, "ploid 4 needs -e 1: with -e 0 the reference's genotype tables are log(0)");
	c->tetra = new TetraState();
	Geometry &g = c->geo;
	TetraState *t = c->tetra;
	t->allo = (c->cfg.autopoly == 0);
	t->Lq = (g.L + TT - 1) / TT * TT;
	t->LTq = t->Lq / TT;
	g.Lpad = t->Lq;
	g.LT = t->LTq;
	c->ns = g.K;
	return IG_OK;
}

void tetra_destroy(ig_ctx *c)
{
	TetraState *t = c->tetra;
	if (!t) return;
	cudaFree(t->Xq); cudaFree(t->Zq); cudaFree(t->Gq); cudaFree(t->cats); cudaFree(t->codes); cudaFree(t->c2i); cudaFree(t->loc_cat);
	cudaFree(t->exf); cudaFree(t->tabC); cudaFree(t->tabP); cudaFree(t->Sprop); cudaFree(t->dstat); cudaFree(t->accepted);
	cudaFree(t->dfix); cudaFree(t->lpart); cudaFree(t->P2); cudaFree(t->n2);
	delete t;
	c->tetra = nullptr;
}

template <typename T>
static cudaError_t dalloc0(T **p, size_t n)
{
	cudaError_t e = cudaMalloc((void **)p, (n ? n : 1) * sizeof(T));
	if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, (n ? n : 1) * sizeof(T), ig_alloc_stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ig_alloc_stream);       // see dalloc (ig_ctx.h)
	return e;
}

// x_dev: int16 [L][Nloc][4] distinct alleles ascending, -1 padding (include/instruct_b200.h)
ig_status tetra_load(ig_ctx *c, const int16_t *x_dev)
{
	Geometry &g = c->geo;
	TetraState *t = c->tetra;
	int amax = 1;
	for (int l = 0; l < g.L; l++) if (c->allelenum_h[l] > amax) amax = c->allelenum_h[l];
	if (amax > TETRA_MAX_A) return fail(IG_ERR_UNSUPPORTED, "ploid 4: allelenum_max %d > %d", amax, TETRA_MAX_A);
	// the code -> index tables hold one byte per genotype: 441 allotetraploid genotypes at 6 alleles do not fit
	g.A = amax < 2 ? 2 : amax;
	// catalogues, one per distinct allele count
	std::vector<int> codes;
	std::vector<uint16_t> c2i;
	std::vector<int32_t> loc_cat(t->Lq, -1);
	t->ncat = 0; t->Gmax = 1;
	for (int n = 1; n <= amax; n++) {
		bool used = false;
		for (int l = 0; l < g.L; l++) used |= (c->allelenum_h[l] == n);
		if (!used) continue;
		TetraCat &cat = t->cat_h[t->ncat];
		cat.n = n; cat.code_off = (int)codes.size(); cat.c2i_off = (int)c2i.size();
		std::vector<int> cd;
		if (t->allo) build_catalogue_allo(n, cd, cat.cls); else build_catalogue(n, cd, cat.cls);
		cat.total = (int)cd.size();
		if (cat.total > t->Gmax) t->Gmax = cat.total;
		c2i.resize(c2i.size() + (size_t)n * n * n * n, 0xFFFFu);
		for (int gi = 0; gi < cat.total; gi++) { codes.push_back(cd[gi]); c2i[cat.c2i_off + cd[gi]] = (uint16_t)gi; }
		for (int l = 0; l < g.L; l++) if (c->allelenum_h[l] == n) loc_cat[l] = t->ncat;
		t->ncat++;
	}
	CK(tetra_configure(g, c->cfg.device, t->allo, t->Gmax, (int)(c2i.size() * sizeof(uint16_t))));
	const size_t tiles = (size_t)t->LTq * g.Nloc * TT * 4;
	const size_t pn = (size_t)g.Lpad * g.A * g.KP;
	const size_t tn = (size_t)t->Lq * g.K * t->Gmax;
	CK(dalloc0(&t->Xq, tiles)); CK(dalloc0(&t->Zq, tiles)); CK(dalloc0(&t->Gq, tiles));
	CK(dalloc0(&t->cats, (size_t)t->ncat)); CK(dalloc0(&t->codes, codes.size())); CK(dalloc0(&t->c2i, c2i.size())); CK(dalloc0(&t->loc_cat, (size_t)t->Lq));
	CK(cudaMemcpyAsync(t->cats, t->cat_h, sizeof(TetraCat) * t->ncat, cudaMemcpyHostToDevice, c->stream)); CK(cudaStreamSynchronize(c->stream));
	CK(cudaMemcpyAsync(t->codes, codes.data(), codes.size() * 4, cudaMemcpyHostToDevice, c->stream)); CK(cudaStreamSynchronize(c->stream));
	CK(cudaMemcpyAsync(t->c2i, c2i.data(), c2i.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream)); CK(cudaStreamSynchronize(c->stream));
	CK(cudaMemcpyAsync(t->loc_cat, loc_cat.data(), loc_cat.size() * 4, cudaMemcpyHostToDevice, c->stream)); CK(cudaStreamSynchronize(c->stream));
	CK(dalloc0(&t->exf, tn)); CK(dalloc0(&t->tabC, tn)); CK(dalloc0(&t->tabP, tn));
	CK(dalloc0(&t->Sprop, (size_t)MAX_K)); CK(dalloc0(&t->dstat, (size_t)MAX_K)); CK(dalloc0(&t->accepted, (size_t)MAX_K));
	CK(dalloc0(&t->dfix, (size_t)MAX_K)); CK(dalloc0(&t->lpart, (size_t)g.nchunks * 2 * g.Nloc));
	CK(dalloc0(&c->P, pn)); CK(dalloc0(&c->n, pn));
	if (t->allo) { CK(dalloc0(&t->P2, pn)); CK(dalloc0(&t->n2, pn)); }
	if (c->cfg.print_freq) CK(dalloc0(&c->P64, (size_t)g.K * g.L * g.A));
	CK(dalloc0(&c->ind, (size_t)c->Npad * g.REC)); CK(dalloc0(&c->Qf, (size_t)g.Nloc * g.KP));
	CK(dalloc0(&c->S, (size_t)MAX_K)); CK(dalloc0(&c->state, (size_t)MAX_K)); CK(dalloc0(&c->state2, (size_t)MAX_K)); CK(dalloc0(&c->sc, 1));
	CK(dalloc0(&c->pcnt, (size_t)g.nchunks * g.Nloc * g.KP)); CK(dalloc0(&c->cnt, (size_t)g.Nloc * g.K));
	CK(dalloc0(&c->initd_dev, (size_t)MAX_K)); CK(dalloc0(&c->scratch, (size_t)64));
	CK(dalloc0(&c->mom.tot, 2)); CK(dalloc0(&c->mom.indvlkh, (size_t)g.N)); CK(dalloc0(&c->mom.qq, (size_t)g.N * g.K)); CK(dalloc0(&c->mom.qq2, (size_t)g.N * g.K));
	CK(dalloc0(&c->mom.self, (size_t)g.K)); CK(dalloc0(&c->mom.self2, (size_t)g.K)); CK(dalloc0(&c->mom.gen, (size_t)g.N)); CK(dalloc0(&c->mom.gen2, (size_t)g.N));
	CK(dalloc0(&c->mom.convg, (size_t)(c->cfg.ckrep > 0 ? c->cfg.ckrep : 1)));
	if (c->cfg.print_freq) { CK(dalloc0(&c->mom.freq, (size_t)g.K * g.L * g.A)); CK(dalloc0(&c->mom.freq2, (size_t)g.K * g.L * g.A)); }
	tile4_kernel<int16_t><<<nb((size_t)t->LTq * g.Nloc * TT, 256), 256, 0, c->stream>>>(x_dev, t->Xq, g.L, g.Nloc, t->LTq, (int16_t)-1);
	CK(cudaGetLastError());
	CK(cudaMemsetAsync(t->Gq, 0xFF, tiles, c->stream));
	c->launches++;
	CK(cudaStreamSynchronize(c->stream));
	c->loaded = true;
	return IG_OK;
}

// ---- phases ---------------------------------------------------------------------------
template <typename K>
static cudaError_t opt_smem(K kernel, size_t bytes) { return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); }

static ig_status tetra_pass_a(ig_ctx *c, int init)
{
	TetraState *t = c->tetra;
	const Geometry &g = c->geo;
	ZsArgs a{t->Zq, t->Gq, c->P, c->Qf, t->tabC, t->tabP, t->loc_cat, t->cats, t->c2i, c->pcnt, t->dfix, g, t->Gmax, init,
	         c->iter, c->key0, c->key1, 0x007fffffu, 0x3f800000u};
	dim3 grid(g.nchunks, g.nblk), block(TETRA_THREADS);
	const size_t sm = smem_zs(g);
	const bool timed = c->profile && !init && c->ev_used + 2 <= (int)c->ev.size();
	if (!init) CK(cudaMemsetAsync(t->dfix, 0, MAX_K * sizeof(unsigned long long), c->stream));
	if (timed) CK(cudaEventRecord(c->ev[c->ev_used], c->stream));
	switch (g.KP) {
	case 4: if (c->rounds == 10) { CK(opt_smem(tetra_zs_kernel<4, 10>, sm)); tetra_zs_kernel<4, 10><<<grid, block, sm, c->stream>>>(a); }
	         else { CK(opt_smem(tetra_zs_kernel<4, 7>, sm)); tetra_zs_kernel<4, 7><<<grid, block, sm, c->stream>>>(a); } break;
	case 8: if (c->rounds == 10) { CK(opt_smem(tetra_zs_kernel<8, 10>, sm)); tetra_zs_kernel<8, 10><<<grid, block, sm, c->stream>>>(a); }
	         else { CK(opt_smem(tetra_zs_kernel<8, 7>, sm)); tetra_zs_kernel<8, 7><<<grid, block, sm, c->stream>>>(a); } break;
	default: if (c->rounds == 10) { CK(opt_smem(tetra_zs_kernel<16, 10>, sm)); tetra_zs_kernel<16, 10><<<grid, block, sm, c->stream>>>(a); }
	         else { CK(opt_smem(tetra_zs_kernel<16, 7>, sm)); tetra_zs_kernel<16, 7><<<grid, block, sm, c->stream>>>(a); } break;
	}
	CK(cudaGetLastError());
	t->timing = timed;           // the closing event is recorded after PASS B: the two passes are one unit of work
	c->launches++;
	return IG_OK;
}

static ig_status tetra_pass_b_allo(ig_ctx *c, int init)
{
	TetraState *t = c->tetra;
	const Geometry &g = c->geo;
	GenoAlloArgs a{t->Xq, t->Zq, t->Gq, c->P, t->P2, c->Qf, t->tabC, t->loc_cat, t->cats, t->c2i, c->n, t->n2, t->lpart, g, t->Gmax, init,
	               c->iter, c->key0, c->key1};
	dim3 grid(g.nchunks, g.nblk), block(TETRA_THREADS);
	const size_t sm = smem_geno_allo(g, t->Gmax);
#define ALLO_LAUNCH1(KPV, RV, ST) do { CK(opt_smem(tetra_geno_allo_kernel<KPV, RV, ST>, sm)); tetra_geno_allo_kernel<KPV, RV, ST><<<grid, block, sm, c->stream>>>(a); } while (0)
#define ALLO_LAUNCH(KPV)                                                                                                   \
	do {                                                                                                               \
		if (c->rounds == 10) { if (g.tab_stage) ALLO_LAUNCH1(KPV, 10, true); else ALLO_LAUNCH1(KPV, 10, false); }     \
		else { if (g.tab_stage) ALLO_LAUNCH1(KPV, 7, true); else ALLO_LAUNCH1(KPV, 7, false); }                       \
	} while (0)
	switch (g.KP) {
	case 4: ALLO_LAUNCH(4); break;
	case 8: ALLO_LAUNCH(8); break;
	default: ALLO_LAUNCH(16); break;
	}
#undef ALLO_LAUNCH1
#undef ALLO_LAUNCH
	CK(cudaGetLastError());
	if (t->timing && !init) { CK(cudaEventRecord(c->ev[c->ev_used + 1], c->stream)); c->ev_used += 2; t->timing = false; }
	c->launches++;
	return IG_OK;
}

static ig_status tetra_pass_b(ig_ctx *c, int init)
{
	TetraState *t = c->tetra;
	const Geometry &g = c->geo;
	if (t->allo) return tetra_pass_b_allo(c, init);
	GenoArgs a{t->Xq, t->Zq, t->Gq, c->P, c->Qf, t->tabC, t->loc_cat, t->cats, t->c2i, c->n, t->lpart, g, t->Gmax, init, c->iter, c->key0, c->key1};
	dim3 grid(g.nchunks, g.nblk), block(TETRA_THREADS);
	const size_t sm = smem_geno(g, t->Gmax);
#define GENO_LAUNCH(KPV, RV)                                                                                          \
	do {                                                                                                          \
		if (g.tab_stage) { CK(opt_smem(tetra_geno_kernel<KPV, RV, true>, sm)); tetra_geno_kernel<KPV, RV, true><<<grid, block, sm, c->stream>>>(a); }   \
		else { CK(opt_smem(tetra_geno_kernel<KPV, RV, false>, sm)); tetra_geno_kernel<KPV, RV, false><<<grid, block, sm, c->stream>>>(a); }             \
	} while (0)
	switch (g.KP) {
	case 4: if (c->rounds == 10) GENO_LAUNCH(4, 10); else GENO_LAUNCH(4, 7); break;
	case 8: if (c->rounds == 10) GENO_LAUNCH(8, 10); else GENO_LAUNCH(8, 7); break;
	default: if (c->rounds == 10) GENO_LAUNCH(16, 10); else GENO_LAUNCH(16, 7); break;
	}
#undef GENO_LAUNCH
	CK(cudaGetLastError());
	if (t->timing && !init) { CK(cudaEventRecord(c->ev[c->ev_used + 1], c->stream)); c->ev_used += 2; t->timing = false; }
	c->launches++;
	return IG_OK;
}

static ig_status tetra_tables(ig_ctx *c, int do_cur, int do_prop)
{
	TetraState *t = c->tetra;
	const Geometry &g = c->geo;
	TabArgs a{c->P, c->allelenum, t->loc_cat, t->cats, t->codes, t->c2i, c->S, t->Sprop, t->exf, t->tabC, t->tabP, g.L, g.K, g.KP, g.A, t->Gmax, do_cur, do_prop,
	          t->P2, t->allo ? 1 : 0};
	tetra_tables_kernel<<<nb((size_t)g.L * g.K * ((do_cur && do_prop) ? 2 : 1), 64), 64, 0, c->stream>>>(a);
	CK(cudaGetLastError());
	c->launches++;
	return IG_OK;
}

static ig_status tetra_q(ig_ctx *c)
{
	const Geometry &g = c->geo;
	tetra_q_kernel<<<nb((size_t)g.Nloc, 32), 32 * CG, 0, c->stream>>>(c->pcnt, c->ind, c->Qf, c->cnt, c->sc, g, c->iter, c->key0, c->key1);
	CK(cudaGetLastError());
	c->launches++;
	return IG_OK;
}

static ig_status tetra_lkh(ig_ctx *c)
{
	tetra_indv_lkh_kernel<<<nb((size_t)c->geo.Nloc, 32), 32 * CG, 0, c->stream>>>(c->tetra->lpart, c->ind, c->geo);
	CK(cudaGetLastError());
	// (Q, indvlkh) of every shard: totallkh, the empty-cluster sums and the moments run redundantly everywhere
	ig_status st = ig_exchange_individuals(c);
	if (st != IG_OK) return st;
	tetra_lkh_kernel<<<1, RED1, 0, c->stream>>>(c->ind, c->sc, c->geo);
	CK(cudaGetLastError());
	c->launches += 2;
	return IG_OK;
}

static ig_status tetra_update_p(ig_ctx *c)
{
	ig_status st = ig_exchange_tally(c);        // int32 n[L][A][K] summed over the shards
	if (st != IG_OK) return st;
	PArgs a{c->n, c->P, c->P64, c->allelenum, c->geo, c->iter, c->key0, c->key1, nullptr, 1};
	CK(launch_p_dirichlet(a, c->stream));
	c->launches++;
	if (c->tetra->allo) {                          // update_P_allo, poly_geno.c:441-518: freq2 from the second subgenome's tally
		TetraState *t = c->tetra;
		if ((st = ig_allreduce_int32(c, t->n2, (size_t)c->geo.Lpad * c->geo.A * c->geo.KP)) != IG_OK) return st;
		PArgs b{t->n2, t->P2, nullptr, c->allelenum, c->geo, c->iter, c->key0, c->key1, nullptr, 1, 1};
		CK(launch_p_dirichlet(b, c->stream));
		c->launches++;
	}
	return IG_OK;
}

static ig_status tetra_s_begin(ig_ctx *c)
{
	TetraState *t = c->tetra;
	tetra_propose_kernel<<<1, 32, 0, c->stream>>>(c->S, t->Sprop, c->geo.K, c->iter, c->key0, c->key1);
	CK(cudaGetLastError());
	c->launches++;
	return tetra_tables(c, 1, 1);
}

static ig_status tetra_s_end(ig_ctx *c, int decide)
{
	TetraState *t = c->tetra;
	const Geometry &g = c->geo;
	ig_status st = ig_allreduce_int64(c, (int64_t *)t->dfix, (size_t)g.K);      // the shards' integer totals: an exact sum
	if (st != IG_OK) return st;
	tetra_accept_kernel<<<1, 32, 0, c->stream>>>(t->dfix, g.K, c->S, t->Sprop, t->dstat, t->accepted, c->sc, c->iter, c->key0, c->key1, decide);
	c->launches++;
	CK(cudaGetLastError());
	if (decide) {
		tetra_select_kernel<<<nb((size_t)g.L * g.K * t->Gmax, 256), 256, 0, c->stream>>>(t->tabC, t->tabP, t->accepted, g.L, g.K, t->Gmax);
		CK(cudaGetLastError());
		c->launches++;
	}
	c->launches++;
	return IG_OK;
}

ig_status tetra_one_sweep(ig_ctx *c)
{
	ig_status st;
	c->iter++;
	if ((st = tetra_update_p(c)) != IG_OK) return st;           // update_P_auto          poly_geno.c:101
	if ((st = tetra_s_begin(c)) != IG_OK) return st;            // calc_exfreq_auto, proposals and both tables  :102,:106
	if ((st = tetra_pass_a(c, 0)) != IG_OK) return st;          // S statistics + update_ZQ (Z half)            :106,:108
	if ((st = tetra_s_end(c, 1)) != IG_OK) return st;           // update_S_POP accepts
	if ((st = tetra_q(c)) != IG_OK) return st;                  // update_ZQ (Q half)
	if ((st = tetra_pass_b(c, 0)) != IG_OK) return st;          // update_geno + cal_lkd + tally                :110,:112
	return tetra_lkh(c);
}

ig_status tetra_chain_init(ig_ctx *c, int32_t chain_id, const float *initd)
{
	TetraState *t = c->tetra;
	const Geometry &g = c->geo;
	c->key0 = (uint32_t)c->cfg.seed ^ (0x9E3779B9u * (uint32_t)(chain_id + 1));
	c->key1 = (uint32_t)(c->cfg.seed >> 32) ^ (0x85EBCA6Bu * (uint32_t)(chain_id + 1));
	c->iter = 0;
	float init_h[MAX_K];
	for (int k = 0; k < MAX_K; k++) init_h[k] = (initd && k < g.K) ? initd[k] : 0.5f;
	if (!initd) {            // read_init without an -i file draws the starting rates from U(0,1) (initial.c:52-58)
		Stream st(0u, 2u, 0u, TAG_INIT, c->key0, c->key1);
		for (int k = 0; k < g.K; k++) init_h[k] = (float)st.uniform();
	}
	CK(cudaMemcpyAsync(c->initd_dev, init_h, sizeof(init_h), cudaMemcpyHostToDevice, c->stream));
	tetra_init_scalars_kernel<<<1, 32, 0, c->stream>>>(c->sc, c->S, c->initd_dev, g.K, c->key0, c->key1);
	CK(cudaGetLastError());
	const size_t pn = (size_t)g.Lpad * g.A * g.KP;
	CK(launch_fill_f32(c->P, 1.0f, pn, c->stream));
	CK(launch_fill_q_uniform(c->Qf, g, c->stream));
	CK(cudaMemsetAsync(c->n, 0, pn * sizeof(int32_t), c->stream));
	CK(cudaMemsetAsync(t->Zq, 0, (size_t)t->LTq * g.Nloc * TT * 4, c->stream));
	c->launches += 3;
	ig_status st;
	if ((st = tetra_pass_b(c, 1)) != IG_OK) return st;          // initial_geno, poly_geno.c:316 (uniform resolution)
	if ((st = tetra_pass_a(c, 1)) != IG_OK) return st;          // update_ZQ(init_flag = 1): uniform z, :87
	if ((st = tetra_q(c)) != IG_OK) return st;
	if (t->allo) CK(cudaMemsetAsync(t->n2, 0, (size_t)g.Lpad * g.A * g.KP * sizeof(int32_t), c->stream));
	CK(launch_tetra_tally(t->Zq, t->Gq, c->n, t->n2, g, t->LTq, c->stream));
	CK(cudaGetLastError());
	c->launches++;
	c->chain_ready = true;
	return IG_OK;
}

ig_status tetra_run_phase(ig_ctx *c, int32_t mask)
{
	ig_status st;
	if (mask & IG_PHASE_UPDATE_P) if ((st = tetra_update_p(c)) != IG_OK) return st;
	if (mask & IG_PHASE_UPDATE_S) if ((st = tetra_s_begin(c)) != IG_OK) return st;
	if (mask & IG_PHASE_ZQ) {
		if ((st = tetra_pass_a(c, 0)) != IG_OK) return st;
		if ((st = tetra_s_end(c, (mask & IG_PHASE_UPDATE_S) ? 1 : 0)) != IG_OK) return st;
		if ((st = tetra_q(c)) != IG_OK) return st;
	}
	if (mask & IG_PHASE_GENO) {
		if ((st = tetra_pass_b(c, 0)) != IG_OK) return st;
		if ((st = tetra_lkh(c)) != IG_OK) return st;
	}
	CK(cudaStreamSynchronize(c->stream));
	return IG_OK;
}

// state hooks: returns IG_ERR_UNSUPPORTED for ids that the common code handles
ig_status tetra_get_state(ig_ctx *c, int32_t id, void *host, size_t bytes, bool *handled)
{
	TetraState *t = c->tetra;
	const Geometry &g = c->geo;
	*handled = true;
	const size_t el = (size_t)g.L * g.Nloc * 4;
	switch (id) {
	case IG_STATE_X: case IG_STATE_Z: case IG_STATE_GENO: {
		const size_t want = el * (id == IG_STATE_X ? 2 : 1);
		if (bytes != want) return fail(IG_ERR_ARG, "X/Z/GENO: expected %zu bytes, got %zu", want, bytes);
		void *tmp = nullptr;
		CK(cudaMalloc(&tmp, want));
		if (id == IG_STATE_X) untile4_kernel<int16_t><<<nb((size_t)g.L * g.Nloc, 256), 256, 0, c->stream>>>(t->Xq, (int16_t *)tmp, g.L, g.Nloc);
		else untile4_kernel<int8_t><<<nb((size_t)g.L * g.Nloc, 256), 256, 0, c->stream>>>(id == IG_STATE_Z ? t->Zq : t->Gq, (int8_t *)tmp, g.L, g.Nloc);
		cudaError_t e = cudaGetLastError();
		if (e == cudaSuccess) e = cudaMemcpyAsync(host, tmp, want, cudaMemcpyDeviceToHost, c->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
		cudaFree(tmp);
		CK(e);
		return IG_OK;
	}
	case IG_STATE_MASK: {      // get_missing_tetra, data_interface.c:722-741: missing <=> no allele observed
		if (bytes != (size_t)g.L * g.Nloc) return fail(IG_ERR_ARG, "MASK: expected %zu bytes, got %zu", (size_t)g.L * g.Nloc, bytes);
		std::vector<int16_t> x(el);
		bool h2 = false;
		ig_status st = tetra_get_state(c, IG_STATE_X, x.data(), el * 2, &h2);
		if (st != IG_OK) return st;
		for (size_t q = 0; q < (size_t)g.L * g.Nloc; q++) ((uint8_t *)host)[q] = x[4 * q] < 0 ? 1 : 0;
		return IG_OK;
	}
	case IG_STATE_TABLES: case IG_STATE_TABLES_PROP: case IG_STATE_EXFREQ: {
		// float [K][L][Gmax], the reference's exfreq / genofreq layout
		const size_t n = (size_t)g.K * g.L * t->Gmax;
		if (bytes != n * 4) return fail(IG_ERR_ARG, "TABLES: expected %zu bytes, got %zu", n * 4, bytes);
		std::vector<float> h((size_t)t->Lq * g.K * t->Gmax);
		const float *src = id == IG_STATE_TABLES ? t->tabC : (id == IG_STATE_TABLES_PROP ? t->tabP : t->exf);
		CK(cudaMemcpy(h.data(), src, h.size() * 4, cudaMemcpyDeviceToHost));
		for (int k = 0; k < g.K; k++) for (int l = 0; l < g.L; l++) for (int q = 0; q < t->Gmax; q++)
			((float *)host)[((size_t)k * g.L + l) * t->Gmax + q] = h[((size_t)l * g.K + k) * t->Gmax + q];
		return IG_OK;
	}
	case IG_STATE_SPROP: case IG_STATE_DSTAT:
		if (bytes != (size_t)g.K * 8) return fail(IG_ERR_ARG, "SPROP/DSTAT: expected %zu bytes", (size_t)g.K * 8);
		CK(cudaMemcpy(host, id == IG_STATE_SPROP ? t->Sprop : t->dstat, bytes, cudaMemcpyDeviceToHost));
		return IG_OK;
	case IG_STATE_GMAX:
		if (bytes != 4) return fail(IG_ERR_ARG, "GMAX: expected 4 bytes");
		*(int32_t *)host = t->Gmax;
		return IG_OK;
	case IG_STATE_P2: case IG_STATE_TALLY2: {
		if (!t->allo) return fail(IG_ERR_ARG, "P2/TALLY2: allotetraploid model only");
		const size_t pn = (size_t)g.Lpad * g.A * g.KP, el = (size_t)g.K * g.L * g.A;
		if (bytes != el * (id == IG_STATE_P2 ? 8 : 4)) return fail(IG_ERR_ARG, "P2/TALLY2: expected %zu bytes, got %zu", el * (id == IG_STATE_P2 ? 8 : 4), bytes);
		std::vector<uint32_t> h(pn);
		CK(cudaMemcpy(h.data(), id == IG_STATE_P2 ? (const void *)t->P2 : (const void *)t->n2, pn * 4, cudaMemcpyDeviceToHost));
		for (int k = 0; k < g.K; k++) for (int l = 0; l < g.L; l++) for (int al = 0; al < g.A; al++) {
			const size_t src = ((size_t)l * g.A + al) * g.KP + k, dst = ((size_t)k * g.L + l) * g.A + al;
			if (id == IG_STATE_P2) { float f; memcpy(&f, &h[src], 4); ((double *)host)[dst] = (double)f; }
			else ((int32_t *)host)[dst] = (int32_t)h[src];
		}
		return IG_OK;
	}
	default:
		*handled = false;
		return IG_OK;
	}
}

ig_status tetra_set_state(ig_ctx *c, int32_t id, const void *host, size_t bytes, bool *handled)
{
	TetraState *t = c->tetra;
	const Geometry &g = c->geo;
	*handled = true;
	switch (id) {
	case IG_STATE_Z: case IG_STATE_GENO: {
		const size_t want = (size_t)g.L * g.Nloc * 4;
		if (bytes != want) return fail(IG_ERR_ARG, "Z/GENO: expected %zu bytes, got %zu", want, bytes);
		void *tmp = nullptr;
		CK(cudaMalloc(&tmp, want));
		cudaError_t e = cudaMemcpyAsync(tmp, host, want, cudaMemcpyHostToDevice, c->stream);
		if (e == cudaSuccess) {
			tile4_kernel<int8_t><<<nb((size_t)t->LTq * g.Nloc * TT, 256), 256, 0, c->stream>>>((const int8_t *)tmp, id == IG_STATE_Z ? t->Zq : t->Gq, g.L, g.Nloc, t->LTq,
			                                                                               (int8_t)(id == IG_STATE_Z ? 0 : -1));
			e = cudaGetLastError();
		}
		if (e == cudaSuccess && id == IG_STATE_GENO) {
			tetra_mask_geno_kernel<<<nb((size_t)t->LTq * g.Nloc * TT, 256), 256, 0, c->stream>>>(t->Xq, t->Gq, (size_t)t->LTq * g.Nloc * TT);
			e = cudaGetLastError();
		}
		// n always mirrors (z, geno)
		if (e == cudaSuccess) e = cudaMemsetAsync(c->n, 0, (size_t)g.Lpad * g.A * g.KP * 4, c->stream);
		if (e == cudaSuccess && t->allo) e = cudaMemsetAsync(t->n2, 0, (size_t)g.Lpad * g.A * g.KP * 4, c->stream);
		if (e == cudaSuccess) e = launch_tetra_tally(t->Zq, t->Gq, c->n, t->n2, g, t->LTq, c->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
		cudaFree(tmp);
		CK(e);
		return IG_OK;
	}
	case IG_STATE_SPROP:
		if (bytes != (size_t)g.K * 8) return fail(IG_ERR_ARG, "SPROP: expected %zu bytes", (size_t)g.K * 8);
		CK(cudaMemcpyAsync(t->Sprop, host, bytes, cudaMemcpyHostToDevice, c->stream)); CK(cudaStreamSynchronize(c->stream));
		return IG_OK;
	case IG_STATE_P2: {
		if (!t->allo) return fail(IG_ERR_ARG, "P2: allotetraploid model only");
		const size_t pn = (size_t)g.Lpad * g.A * g.KP;
		if (bytes != (size_t)g.K * g.L * g.A * 8) return fail(IG_ERR_ARG, "P2: expected %zu bytes, got %zu", (size_t)g.K * g.L * g.A * 8, bytes);
		std::vector<float> p(pn, 0.0f);
		for (int k = 0; k < g.K; k++) for (int l = 0; l < g.L; l++) for (int al = 0; al < g.A; al++)
			p[((size_t)l * g.A + al) * g.KP + k] = (float)((const double *)host)[((size_t)k * g.L + l) * g.A + al];
		CK(cudaMemcpyAsync(t->P2, p.data(), pn * 4, cudaMemcpyHostToDevice, c->stream)); CK(cudaStreamSynchronize(c->stream));
		return IG_OK;
	}
	case IG_STATE_TABLES: {
		// recompute exfreq and BOTH tables from the current P, S and S' (no proposal draw)
		ig_status st = tetra_tables(c, 1, 1);
		if (st != IG_OK) return st;
		CK(cudaStreamSynchronize(c->stream));
		return IG_OK;
	}
	default:
		*handled = false;
		return IG_OK;
	}
}
