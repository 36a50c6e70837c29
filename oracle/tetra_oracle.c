/*
 * oracle/tetra_oracle.c -- CPU restatement of the reference's autotetraploid sweep
 * (slowkoni/InStruct poly_geno.c, `-p 4 -ap 1`).  See tetra_oracle.h for status and pinning.
 *
 * TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path.
 *
 * The arithmetic (operation order, float/double mix, libm calls) follows the reference so that
 * tables and chains agree bit for bit; the data structures (flat packed arrays, one genotype
 * catalogue per distinct allele count) are this repo's own.  Two slips of the reference that
 * shape its numbers are reproduced and marked "as written".
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "tetra_oracle.h"

#define MIN2(X, Y) (((X) > (Y)) ? (Y) : (X))     /* mcmc.h:10 */

double orc_rng_u01(orc_rng *r);
void orc_rng_dirichlet(orc_rng *r, const double *alpha, int len, double *out, double add);
int orc_rng_bracket(orc_rng *r, double *vec, int len);
double orc_rng_three_state(orc_rng *r, int *new_state, int cur);
double orc_trans_prob(int a, int b);

/* genotype catalogue for loci with n alleles: codes are the four alleles read as a base-n
 * number, grouped by dosage class (auto_geno_num / auto_geno_list, poly_geno.c:1698-1800) */
typedef struct {
	int n;
	int cls[5];        /* mono iiii, simplex iiij, duplex iijj, tri iijk, quadri ijkl */
	int total;
	int *code;
} tet_cat;

struct tet_model {
	int N, L, K, Amax, Gmax, back_refl;
	int autopoly;      /* 1: autotetraploid (-ap 1); 0: allotetraploid (-ap 0): copies 0,1 and copies 2,3 are two subgenomes */
	const int16_t *x;
	const uint8_t *nd;
	const int32_t *allelenum;
	int8_t *geno, *z;
	double *qq, *qqnum, *freq;
	double *freq2;     /* allotetraploid: allele frequencies of the second subgenome (UPMCMC.freq2, mcmc.h:16) */
	double alpha, totallkh;
	double *self_rates, *indvlkh;
	int *state;
	int ncat;
	tet_cat *cat;
	int *cat_of;       /* [L] catalogue of each locus */
	float *exfreq, *genofreq;     /* [K][L][Gmax], natural logs */
	orc_rng rng;
};

static inline long XO(const tet_model *m, int l, int i) { return ((long)l * m->N + i) * 4; }
static inline long FO(const tet_model *m, int k, int l, int a) { return ((long)k * m->L + l) * m->Amax + a; }
static inline long TO(const tet_model *m, int k, int l) { return ((long)k * m->L + l) * m->Gmax; }

/* ---------------------------------------------------------------- catalogue ----------- */
static void build_catalogue(tet_cat *c, int n)
{
	int j, k, a, b, p = 0;
	c->n = n;
	c->cls[0] = n; c->cls[1] = n * (n - 1); c->cls[2] = n * (n - 1) / 2;
	c->cls[3] = n * (n - 1) * (n - 2) / 2; c->cls[4] = n * (n - 1) * (n - 2) * (n - 3) / 24;
	c->total = n + n * (n - 1) * 3 / 2 + n * (n - 1) * (n - 2) / 2 + n * (n - 1) * (n - 2) * (n - 3) / 24;
	c->code = (int *)malloc(sizeof(int) * (c->total > 0 ? c->total : 1));
	for (j = 0; j < n; j++) c->code[p++] = j * (n * n * n + n * n + n + 1);
	for (j = 0; j < n - 1; j++) for (k = j + 1; k < n; k++) {        /* jjjk then kkkj */
		c->code[p++] = j * (n * n * n + n * n + n) + k;
		c->code[p++] = n * (n * n + n + 1) * k + j;
	}
	for (j = 0; j < n - 1; j++) for (k = j + 1; k < n; k++) c->code[p++] = j * (n * n * n + n * n) + k * (n + 1);
	for (j = 0; j < n - 2; j++) for (k = j + 1; k < n - 1; k++) for (a = k + 1; a < n; a++) {
		c->code[p++] = j * (n * n * n + n * n) + k * n + a;
		c->code[p++] = k * (n * n * n + n * n) + j * n + a;
		c->code[p++] = a * (n * n * n + n * n) + j * n + k;
	}
	for (j = 0; j < n - 3; j++) for (k = j + 1; k < n - 2; k++) for (a = k + 1; a < n - 1; a++) for (b = a + 1; b < n; b++)
		c->code[p++] = j * n * n * n + k * n * n + n * a + b;
}

/* allo_geno_num / allo_geno_list, poly_geno.c:2031-2120: classes iikk | iikl (k<l) | ijkk (i<j) | ijkl
 * (i<j, k<l); cls[0..3] hold the four class sizes, cls[4] = 0 */
static void build_catalogue_allo(tet_cat *c, int n)
{
	int j, k, a, b, p = 0;
	c->n = n;
	c->cls[0] = n * n; c->cls[1] = n * (n - 1) / 2 * n; c->cls[2] = n * (n - 1) / 2 * n;
	c->cls[3] = n * (n - 1) * n * (n - 1) / 4; c->cls[4] = 0;
	c->total = n * n + n * (n - 1) * n + n * (n - 1) * n * (n - 1) / 4;
	c->code = (int *)malloc(sizeof(int) * (c->total > 0 ? c->total : 1));
	for (j = 0; j < n; j++) for (k = 0; k < n; k++) c->code[p++] = j * n * n * (n + 1) + k * (n + 1);
	for (j = 0; j < n; j++) for (k = 0; k < n - 1; k++) for (a = k + 1; a < n; a++) c->code[p++] = j * n * n * (n + 1) + n * k + a;
	for (j = 0; j < n - 1; j++) for (k = j + 1; k < n; k++) for (a = 0; a < n; a++) c->code[p++] = (j * n + k) * n * n + a * (n + 1);
	for (j = 0; j < n - 1; j++) for (k = j + 1; k < n; k++) for (a = 0; a < n - 1; a++) for (b = a + 1; b < n; b++)
		c->code[p++] = j * n * n * n + k * n * n + a * n + b;
}

/* find_id, poly_geno.c:2367 */
static int lookup(const tet_cat *c, int code)
{
	int g;
	for (g = 0; g < c->total; g++) if (c->code[g] == code) return g;
	fprintf(stderr, "tetra oracle: genotype code %d not in the catalogue of n=%d\n", code, c->n);
	exit(1);
}

static int member(int v, const int *set, int len) { int i; for (i = 0; i < len; i++) if (set[i] == v) return 1; return 0; }

tet_model *tet_new(int N, int L, int K, int back_refl, const int16_t *x, const uint8_t *nd, const int32_t *allelenum)
{
	return tet_new2(N, L, K, back_refl, 1, x, nd, allelenum);
}

tet_model *tet_new2(int N, int L, int K, int back_refl, int autopoly, const int16_t *x, const uint8_t *nd, const int32_t *allelenum)
{
	int l, j, i, k;
	tet_model *m = (tet_model *)calloc(1, sizeof(tet_model));
	m->N = N; m->L = L; m->K = K; m->back_refl = back_refl; m->x = x; m->nd = nd; m->allelenum = allelenum;
	m->autopoly = autopoly;
	m->Amax = 1;
	for (l = 0; l < L; l++) if (allelenum[l] > m->Amax) m->Amax = allelenum[l];
	/* one catalogue per distinct allele count, ascending (gen_allele_poly, poly_geno.c:1673) */
	m->cat = (tet_cat *)calloc(m->Amax + 1, sizeof(tet_cat));
	m->cat_of = (int *)calloc(L, sizeof(int));
	for (j = 1; j <= m->Amax; j++) {
		int used = 0;
		for (l = 0; l < L; l++) if (allelenum[l] == j) used = 1;
		if (!used) continue;
		if (autopoly) build_catalogue(&m->cat[m->ncat], j); else build_catalogue_allo(&m->cat[m->ncat], j);
		if (m->cat[m->ncat].total > m->Gmax) m->Gmax = m->cat[m->ncat].total;
		for (l = 0; l < L; l++) if (allelenum[l] == j) m->cat_of[l] = m->ncat;
		m->ncat++;
	}
	m->geno = (int8_t *)calloc((size_t)L * N * 4, 1);
	m->z = (int8_t *)calloc((size_t)L * N * 4, 1);
	m->qq = (double *)calloc((size_t)N * K, sizeof(double));
	m->qqnum = (double *)calloc((size_t)N * K, sizeof(double));
	m->freq = (double *)calloc((size_t)K * L * m->Amax, sizeof(double));
	m->freq2 = (double *)calloc((size_t)K * L * m->Amax, sizeof(double));
	m->self_rates = (double *)calloc(K, sizeof(double));
	m->state = (int *)calloc(K, sizeof(int));
	m->indvlkh = (double *)calloc(N, sizeof(double));
	m->exfreq = (float *)calloc((size_t)K * L * m->Gmax, sizeof(float));
	m->genofreq = (float *)calloc((size_t)K * L * m->Gmax, sizeof(float));
	m->alpha = 1.0;
	for (i = 0; i < N; i++) for (k = 0; k < K; k++) m->qq[(long)i * K + k] = 1.0 / K;
	m->rng.s1 = 13; m->rng.s2 = 4; m->rng.s3 = 1972;
	return m;
}

void tet_free(tet_model *m)
{
	int j;
	if (!m) return;
	for (j = 0; j < m->ncat; j++) free(m->cat[j].code);
	free(m->cat); free(m->cat_of); free(m->geno); free(m->z); free(m->qq); free(m->qqnum); free(m->freq); free(m->freq2);
	free(m->self_rates); free(m->state); free(m->indvlkh); free(m->exfreq); free(m->genofreq); free(m);
}

void tet_setseeds(tet_model *m, long a, long b, long c) { m->rng.s1 = a; m->rng.s2 = b; m->rng.s3 = c; }
int8_t *tet_z(tet_model *m) { return m->z; }
int8_t *tet_geno(tet_model *m) { return m->geno; }
double *tet_qq(tet_model *m) { return m->qq; }
double *tet_qqnum(tet_model *m) { return m->qqnum; }
double *tet_freq(tet_model *m) { return m->freq; }
double *tet_freq2(tet_model *m) { return m->freq2; }
double *tet_self(tet_model *m) { return m->self_rates; }
int *tet_state(tet_model *m) { return m->state; }
double *tet_alpha(tet_model *m) { return &m->alpha; }
double *tet_indvlkh(tet_model *m) { return m->indvlkh; }
double *tet_totallkh(tet_model *m) { return &m->totallkh; }
float *tet_exfreq(tet_model *m) { return m->exfreq; }
float *tet_genofreq(tet_model *m) { return m->genofreq; }
int tet_gmax(const tet_model *m) { return m->Gmax; }
int tet_amax(const tet_model *m) { return m->Amax; }
int tet_genolist(const tet_model *m, int l, int *codes)
{
	const tet_cat *c = &m->cat[m->cat_of[l]];
	memcpy(codes, c->code, sizeof(int) * c->total);
	return c->total;
}

/* ---------------------------------------------------------------- genotype classes ---- */
/* get_cat_auto, poly_geno.c:1313: 0 iiii, 1 iiij, 2 iijj, 3 iijk, 4 ijkl (first-seen order) */
static int dosage_class(const int8_t *g)
{
	int seen[4], ns = 1, i, c0 = 0;
	seen[0] = g[0];
	for (i = 1; i < 4; i++) if (!member(g[i], seen, ns)) seen[ns++] = g[i];
	if (ns == 1) return 0;
	if (ns == 2) { for (i = 0; i < 4; i++) if (g[i] == seen[0]) c0++; return c0 == 2 ? 2 : 1; }
	return ns == 3 ? 3 : 4;
}

/* get_cat_allo, poly_geno.c:1341-1372: 0 iikk, 1 iikl, 2 ijkk, 3 ijkl */
static int allo_class(const int8_t *g)
{
	const int c1 = (g[0] == g[1]) ? 1 : 2, c2 = (g[2] == g[3]) ? 1 : 2;
	if (c1 + c2 == 2) return 0;
	if (c1 + c2 == 3) return (c1 == 1 && c2 == 2) ? 1 : 2;
	return 3;
}

/* get_index_auto, poly_geno.c:1289 (genotypes here always follow the writing rules of
 * check_rule_auto :1346: two_/tri_allele_auto only emit canonical forms) */
int tet_geno_index(const tet_model *m, int l, const int8_t *g)
{
	const int n = m->allelenum[l];
	int code = g[0], i;
	for (i = 1; i < 4; i++) code = code * n + g[i];
	return lookup(&m->cat[m->cat_of[l]], code);
}

static inline int all_same4(const int8_t *z) { return z[0] == z[1] && z[1] == z[2] && z[2] == z[3]; }   /* chcksame()==0, mcmc.c:1658 */

/* ---------------------------------------------------------------- tables -------------- */
/* calc_exfreq_auto, poly_geno.c:1515-1577: log Hardy-Weinberg genotype frequencies, float */
static void calc_exfreq_allo(tet_model *m);
void tet_calc_exfreq(tet_model *m)
{
	int k, l, g, d[4], t, q;
	if (!m->autopoly) { calc_exfreq_allo(m); return; }
	for (k = 0; k < m->K; k++)
		for (l = 0; l < m->L; l++) {
			const int n = m->allelenum[l];
			const tet_cat *c = &m->cat[m->cat_of[l]];
			const double *f = m->freq + FO(m, k, l, 0);
			float *R = m->exfreq + TO(m, k, l);
			int lo = 0;
			for (g = lo; g < lo + c->cls[0]; g++) R[g] = (float)log(f[c->code[g] % n]) * (float)4;
			lo += c->cls[0];
			for (g = lo; g < lo + c->cls[1]; g++) {
				t = c->code[g]; d[0] = t % n; t /= n; d[1] = t % n;
				R[g] = (float)(log(4.0) + log(f[d[1]]) * (float)(4 - 1) + log(f[d[0]]));
			}
			lo += c->cls[1];
			for (g = lo; g < lo + c->cls[2]; g++) {
				t = c->code[g]; d[0] = t % n; t /= (n * n); d[1] = t % n;
				R[g] = (float)(log(6.0) + (log(f[d[1]]) + log(f[d[0]])) * (4 / 2));
			}
			lo += c->cls[2];
			for (g = lo; g < lo + c->cls[3]; g++) {
				t = c->code[g];
				for (q = 0; q < 3; q++) { d[q] = t % n; t /= n; }
				R[g] = (float)(log(12.0) + log(f[d[2]]) * (4 / 2) + log(f[d[0]]) + log(f[d[1]]));
			}
			lo += c->cls[3];
			for (g = lo; g < lo + c->cls[4]; g++) {
				t = c->code[g];
				for (q = 0; q < 4; q++) { d[q] = t % n; t /= n; }
				R[g] = (float)log(24.0);
				for (q = 0; q < 4; q++) R[g] += (float)log(f[d[q]]);
			}
		}
}

/* gaussj (poly_geno.c:2384, Numerical-Recipes Gauss-Jordan with full pivoting) for the 3x3
 * float system of the triallelic class; same pivot choice and elimination order */
static void solve3(float a[3][3], float b[3])
{
	int piv[3] = {0, 0, 0}, ir[3], ic[3], i, j, k, row = 0, col = 0;
	for (i = 0; i < 3; i++) {
		float big = 0.0f;
		for (j = 0; j < 3; j++)
			if (piv[j] != 1)
				for (k = 0; k < 3; k++)
					if (piv[k] == 0 && fabs(a[j][k]) >= big) { big = fabs(a[j][k]); row = j; col = k; }
		++piv[col];
		if (row != col) {
			float t;
			for (k = 0; k < 3; k++) { t = a[row][k]; a[row][k] = a[col][k]; a[col][k] = t; }
			t = b[row]; b[row] = b[col]; b[col] = t;
		}
		ir[i] = row; ic[i] = col;
		{
			float pivinv = 1.0 / a[col][col];
			a[col][col] = 1.0;
			for (k = 0; k < 3; k++) a[col][k] *= pivinv;
			b[col] *= pivinv;
		}
		for (j = 0; j < 3; j++)
			if (j != col) {
				float dum = a[j][col];
				a[j][col] = 0.0;
				for (k = 0; k < 3; k++) a[j][k] -= a[col][k] * dum;
				b[j] -= b[col] * dum;
			}
	}
	(void)ir; (void)ic;      /* the column un-scrambling only permutes the inverse, not b */
}

/* calc_val, poly_geno.c:2307: the ijkl code that contains the ascending triple d plus v */
static int quad_with(const int *d, int v, int n)
{
	if (v < d[0]) return v * n * n * n + d[0] * n * n + d[1] * n + d[2];
	if (v > d[0] && v < d[1]) return d[0] * n * n * n + v * n * n + d[1] * n + d[2];
	if (v > d[1] && v < d[2]) return d[0] * n * n * n + d[1] * n * n + v * n + d[2];
	if (v > d[2]) return d[0] * n * n * n + d[1] * n * n + d[2] * n + v;
	return 0;
}
/* calc_val2, poly_geno.c:2333: the ijkl code of {d[1] < d[0]} plus {v1 < v2} */
static int quad_with2(const int *d, int v1, int v2, int n)
{
	if (v2 < d[1]) return v1 * n * n * n + v2 * n * n + d[1] * n + d[0];
	if (v2 > d[1] && v2 < d[0] && v1 < d[1]) return v1 * n * n * n + d[1] * n * n + v2 * n + d[0];
	if (v1 > d[1] && v2 < d[0]) return d[1] * n * n * n + v1 * n * n + v2 * n + d[0];
	if (v1 > d[1] && v1 < d[0] && v2 > d[0]) return d[1] * n * n * n + v1 * n * n + d[0] * n + v2;
	if (v1 > d[0]) return d[1] * n * n * n + d[0] * n * n + v1 * n + v2;
	if (v1 < d[1] && v2 > d[0]) return v1 * n * n * n + d[1] * n * n + n * d[0] + v2;
	return 0;
}

/* auto_genfreq, poly_geno.c:1803-2028: log genotype frequencies under partial selfing, solved
 * class by class from the most to the least heterozygous: (I - sA) P = (1 - s) R */
static void genfreq_locus(const tet_model *m, float self, int k, int l, float *P)
{
	const int n = m->allelenum[l];
	const tet_cat *c = &m->cat[m->cat_of[l]];
	const float *R = m->exfreq + TO(m, k, l);
	int hi = c->total, i, j, q, num = 0, d[3];
	float temp;

	if (n >= 4)                                                        /* ijkl, :1817-1829 */
		for (i = hi - c->cls[4]; i < hi; i++) P[i] = log(1 - self) + R[i] - log(1 - self / 6);
	if (n >= 3) {                                                      /* iijk in triples, :1832-1879 */
		hi -= c->cls[4];
		for (i = 0; i < c->cls[3] / 3; i++) {
			const int base = hi - c->cls[3] + i * 3;
			float A[3][3], v[3];
			num = c->code[base];
			for (j = 2; j >= 0; j--) { d[j] = num % n; num /= n; }
			temp = 0;
			if (n >= 4)
				for (q = 0; q < n; q++)
					if (!member(q, d, 3)) { num = lookup(c, quad_with(d, q, n)); temp += exp(P[num]); }
			for (j = 0; j < 3; j++) {
				for (q = 0; q < 3; q++) A[j][q] = (j == q) ? 1 - self * 10.0 / 36.0 : -self / 9.0;
				v[j] = self / 18.0 * temp + (1.0 - self) * exp(R[base + j]);
			}
			temp = v[0];
			for (j = 0; j < 3; j++) v[j] /= temp;
			solve3(A, v);
			for (j = 0; j < 3; j++) P[base + j] = log(v[j]) + log(temp);
		}
	} else hi -= c->cls[4];
	hi -= c->cls[3];
	for (i = hi - c->cls[2]; i < hi; i++) {                            /* iijj, :1882-1931 */
		num = c->code[i];
		d[0] = num % n; num /= (n * n); d[1] = num % n;                /* d[0] > d[1] */
		temp = 0;
		if (n >= 3)
			for (j = 0; j < n; j++) {
				if (member(j, d, 2)) continue;
				if (d[0] < j) num = lookup(c, d[1] * n * n * (n + 1) + d[0] * n + j);
				else if (d[0] > j) num = lookup(c, d[1] * n * n * (n + 1) + j * n + d[0]);
				temp += exp(P[num]) / 9.0 * self;
				if (d[1] < j) num = lookup(c, d[0] * n * n * (n + 1) + d[1] * n + j);
				else if (d[1] > j) num = lookup(c, d[0] * n * n * (n + 1) + j * n + d[1]);
				temp += exp(P[num]) / 9.0 * self;
				num = lookup(c, j * n * n * (n + 1) + d[1] * n + d[0]);
				temp += exp(P[num]) / 36.0 * self;
				if (n >= 4)
					for (q = j + 1; q < n; q++)
						if (!member(q, d, 2)) { num = lookup(c, quad_with2(d, j, q, n)); temp += exp(P[num]) / 36.0 * self; }
			}
		P[i] = log((1 - self) * exp(R[i]) + temp) - log(1 - self / 2.0);
	}
	hi -= c->cls[2];
	for (i = hi - c->cls[1]; i < hi; i++) {                            /* iiij, :1934-1966 */
		num = c->code[i];
		d[0] = num % n; num /= n; d[1] = num % n;
		if (d[0] < d[1]) num = lookup(c, (d[0] * n * n + d[1]) * (n + 1));
		else if (d[0] > d[1]) num = lookup(c, (d[1] * n * n + d[0]) * (n + 1));
		temp = 8.0 / 36.0 * exp(P[num]) * self;
		if (n >= 3)
			for (j = 0; j < n; j++) {
				if (member(j, d, 2)) continue;
				if (d[0] < j) num = lookup(c, d[1] * n * n * (n + 1) + d[0] * n + j);
				else if (d[0] > j) num = lookup(c, d[1] * n * n * (n + 1) + j * n + d[0]);
				temp += exp(P[num]) / 9.0 * self;
			}
		P[i] = log((1 - self) * exp(R[i]) + temp) - log(1 - self / 2.0);
	}
	hi -= c->cls[1];
	for (i = hi - c->cls[0]; i < hi; i++) {                            /* iiii, :1969-2013 */
		num = c->code[i];
		d[0] = num % n;
		temp = 0;
		for (j = 0; j < n; j++) {
			if (j == d[0]) continue;
			num = lookup(c, d[0] * n * (n * n + n + 1) + j);
			temp += exp(P[num]) / 4.0 * self;
			/* as written (:1984-1989): the second branch repeats the first test, so for
			 * j < i the iijj term re-uses the iiij index found just above */
			if (d[0] < j) num = lookup(c, d[0] * n * n * (n + 1) + j * (n + 1));
			temp += exp(P[num]) / 36.0 * self;
			if (n >= 3)
				for (q = j + 1; q < n; q++)
					if (q != d[0]) {
						num = lookup(c, d[0] * n * n * (n + 1) + j * n + q);      /* i i j q, j < q */
						temp += exp(P[num]) / 36.0 * self;
					}
		}
		P[i] = log((1 - self) * exp(R[i]) + temp) - log(1 - self);
	}
}

/* calc_self_genofreq, poly_geno.c:1219-1233 (self_rate arrives as double, is passed as float) */
/* calc_exfreq_allo, poly_geno.c:1592-1671: log Hardy-Weinberg frequencies of the two-subgenome
 * genotypes; the float / double mix of the reference is kept */
static void calc_exfreq_allo(tet_model *m)
{
	int k, l, g, d[4], t, q;
	for (k = 0; k < m->K; k++)
		for (l = 0; l < m->L; l++) {
			const int n = m->allelenum[l];
			const tet_cat *c = &m->cat[m->cat_of[l]];
			const double *f = m->freq + FO(m, k, l, 0), *f2 = m->freq2 + FO(m, k, l, 0);
			float *R = m->exfreq + TO(m, k, l);
			int lo = 0;
			for (g = lo; g < lo + c->cls[0]; g++) {                     /* iikk */
				t = c->code[g]; d[0] = t % n; t /= (n * n); d[1] = t % n;
				R[g] = (float)((log(f[d[1]]) + log(f2[d[0]])) * (4 / 2));
			}
			lo += c->cls[0];
			for (g = lo; g < lo + c->cls[1]; g++) {                     /* iikl */
				t = c->code[g];
				for (q = 0; q < 3; q++) { d[q] = t % n; t /= n; }
				R[g] = (float)(log(2.0) + log(f[d[2]]) * (4 / 2) + log(f2[d[0]]) + log(f2[d[1]]));
			}
			lo += c->cls[1];
			for (g = lo; g < lo + c->cls[2]; g++) {                     /* ijkk */
				t = c->code[g]; t /= n;
				for (q = 0; q < 3; q++) { d[q] = t % n; t /= n; }
				R[g] = (float)(log(2.0) + log(f2[d[0]]) * (4 / 2) + log(f[d[2]]) + log(f[d[1]]));
			}
			lo += c->cls[2];
			for (g = lo; g < lo + c->cls[3]; g++) {                     /* ijkl */
				t = c->code[g];
				for (q = 0; q < 4; q++) { d[q] = t % n; t /= n; }
				R[g] = (float)log(4.0);
				for (q = 0; q < 2; q++) R[g] += (float)log(f2[d[q]]);
				for (q = 2; q < 4; q++) R[g] += (float)log(f[d[q]]);
			}
		}
}

/* allo_genfreq, poly_geno.c:2122-2305: (I - sA) P = (1 - s) R solved class by class from the
 * doubly heterozygous genotypes down; operation order and float / double mix as in the reference */
static void genfreq_locus_allo(const tet_model *m, float self, int k, int l, float *P)
{
	const int n = m->allelenum[l];
	const tet_cat *c = &m->cat[m->cat_of[l]];
	const float *R = m->exfreq + TO(m, k, l);
	int i, j, kk, tmp, num, d[3];
	float temp;
	tmp = c->total;
	for (i = tmp - c->cls[3]; i < tmp; i++)                             /* ijkl */
		P[i] = log(1 - self) + R[i] - log(1 - self / 4);
	tmp -= c->cls[3];
	for (i = tmp - c->cls[2]; i < tmp; i++) {                           /* ijkk */
		num = c->code[i];
		d[0] = num % n; num /= (n * n); d[1] = num % n; num /= n; d[2] = num % n;
		temp = 0;
		for (j = 0; j < n; j++)
			if (j != d[0]) {
				if (d[0] < j) num = lookup(c, d[2] * n * n * n + d[1] * n * n + d[0] * n + j);
				else num = lookup(c, d[2] * n * n * n + d[1] * n * n + j * n + d[0]);
				temp += exp(P[num]) * self / 8.0;
			}
		P[i] = log((1 - self) * exp(R[i]) + temp) - log(1 - self / 2.0);
	}
	tmp -= c->cls[2];
	for (i = tmp - c->cls[1]; i < tmp; i++) {                           /* iikl */
		num = c->code[i];
		d[0] = num % n; num /= n; d[1] = num % n; num /= n; d[2] = num % n;
		temp = 0;
		for (j = 0; j < n; j++)
			if (j != d[2]) {
				if (d[2] < j) num = lookup(c, d[2] * n * n * n + j * n * n + d[1] * n + d[0]);
				else num = lookup(c, j * n * n * n + d[2] * n * n + d[1] * n + d[0]);
				temp += exp(P[num]) * self / 8.0;
			}
		P[i] = log((1 - self) * exp(R[i]) + temp) - log(1 - self / 2.0);
	}
	tmp -= c->cls[1];
	for (i = tmp - c->cls[0]; i < tmp; i++) {                           /* iikk */
		num = c->code[i];
		d[0] = num % n; num /= (n * n); d[1] = num % n;
		temp = 0;
		for (j = 0; j < n; j++)                                         /* P_iikl */
			if (j != d[0]) {
				if (d[0] < j) num = lookup(c, d[1] * n * n * (n + 1) + d[0] * n + j);
				else num = lookup(c, d[1] * n * n * (n + 1) + j * n + d[0]);
				temp += exp(P[num]) * self / 4.0;
			}
		for (j = 0; j < n; j++)                                         /* P_ijkk */
			if (j != d[1]) {
				if (d[1] < j) num = lookup(c, d[1] * n * n * n + j * n * n + d[0] * (n + 1));
				else num = lookup(c, j * n * n * n + d[1] * n * n + d[0] * (n + 1));
				temp += exp(P[num]) * self / 4.0;
			}
		for (j = 0; j < n; j++)                                         /* P_ijkl */
			for (kk = 0; kk < n; kk++)
				if (j != d[1] && kk != d[0]) {
					const int a0 = d[1] < j ? d[1] : j, a1 = d[1] < j ? j : d[1];
					const int b0 = d[0] < kk ? d[0] : kk, b1 = d[0] < kk ? kk : d[0];
					num = lookup(c, a0 * n * n * n + a1 * n * n + b0 * n + b1);
					temp += exp(P[num]) * self / 16.0;
				}
		P[i] = log((1 - self) * exp(R[i]) + temp) - log(1 - self);
	}
}

void tet_calc_genofreq(tet_model *m, int k, double self, float *out)
{
	int l;
	for (l = 0; l < m->L; l++) {
		if (m->autopoly) genfreq_locus(m, (float)self, k, l, out + (long)l * m->Gmax);
		else genfreq_locus_allo(m, (float)self, k, l, out + (long)l * m->Gmax);
	}
}

/* ---------------------------------------------------------------- likelihood ---------- */
/* calc_genofq, poly_geno.c:1235-1286, with the table of population z[0] */
static double site_loglik(const tet_model *m, int l, int i, const int8_t *z, const float *tab_of_z0)
{
	const int8_t *g = m->geno + XO(m, l, i);
	double ld = 0;
	int c, q;
	if (m->nd[(long)l * m->N + i] == 0) return 0;
	if (all_same4(z)) return (double)tab_of_z0[tet_geno_index(m, l, g)];
	if (!m->autopoly) {                                    /* :1267-1283 */
		c = allo_class(g);
		for (q = 0; q < 2; q++) ld += log(m->freq[FO(m, z[q], l, g[q])]);
		for (q = 2; q < 4; q++) ld += log(m->freq2[FO(m, z[q], l, g[q])]);
		switch (c) {
		case 1: ld += log(2); break;
		case 2: ld += log(2); break;
		case 3: ld += log(4); break;
		}
		return ld;
	}
	c = dosage_class(g);
	for (q = 0; q < 4; q++) ld += log(m->freq[FO(m, z[q], l, g[q])]);
	switch (c) {
	case 1: ld += log(4); break;
	case 2: ld += log(6); break;
	case 3: ld += log(12); break;
	case 4: ld += log(24); break;
	}
	return ld;
}
double tet_site_loglik(const tet_model *m, int l, int i, const int8_t *z4)
{
	return site_loglik(m, l, i, z4, m->genofreq + TO(m, z4[0], l));
}

/* cal_lkd, poly_geno.c:715-736 */
double tet_cal_lkd(tet_model *m)
{
	double sum = 0;
	int i, l;
	for (i = 0; i < m->N; i++) {
		double ld = 0;
		for (l = 0; l < m->L; l++)
			if (m->nd[(long)l * m->N + i] != 0) ld += tet_site_loglik(m, l, i, m->z + XO(m, l, i));
		m->indvlkh[i] = ld;
		sum += ld;
	}
	return sum;
}

/* cal_lkd_props, poly_geno.c:645-711: the same sum with population k's table replaced */
double tet_cal_lkd_props(const tet_model *m, int k, const float *tab)
{
	double ld = 0;
	int i, l;
	for (i = 0; i < m->N; i++)
		for (l = 0; l < m->L; l++) {
			const int8_t *z = m->z + XO(m, l, i);
			if (m->nd[(long)l * m->N + i] == 0) continue;
			if (all_same4(z) && z[0] == k) ld += (double)tab[(long)l * m->Gmax + tet_geno_index(m, l, m->geno + XO(m, l, i))];
			else ld += tet_site_loglik(m, l, i, z);
		}
	return ld;
}

/* ---------------------------------------------------------------- tallies ------------- */
/* the count half of update_P_auto, poly_geno.c:390-424: over the LATENT genotype, every
 * non-missing locus (no allelenum > 1 guard) */
void tet_tally(const tet_model *m, int32_t *n)
{
	int i, l, q;
	memset(n, 0, sizeof(int32_t) * (size_t)m->K * m->L * m->Amax);
	for (l = 0; l < m->L; l++)
		for (i = 0; i < m->N; i++) {
			if (m->nd[(long)l * m->N + i] == 0) continue;
			for (q = 0; q < 4; q++) n[FO(m, m->z[XO(m, l, i) + q], l, m->geno[XO(m, l, i) + q])]++;
		}
}

/* the count half of update_P_allo, poly_geno.c:441-489: copies 0,1 and copies 2,3 are tallied apart */
void tet_tally_allo(const tet_model *m, int32_t *n1, int32_t *n2)
{
	int i, l, q;
	memset(n1, 0, sizeof(int32_t) * (size_t)m->K * m->L * m->Amax);
	memset(n2, 0, sizeof(int32_t) * (size_t)m->K * m->L * m->Amax);
	for (l = 0; l < m->L; l++)
		for (i = 0; i < m->N; i++) {
			if (m->nd[(long)l * m->N + i] == 0) continue;
			for (q = 0; q < 2; q++) n1[FO(m, m->z[XO(m, l, i) + q], l, m->geno[XO(m, l, i) + q])]++;
			for (q = 2; q < 4; q++) n2[FO(m, m->z[XO(m, l, i) + q], l, m->geno[XO(m, l, i) + q])]++;
		}
}

void tet_count_z(const tet_model *m, double *cnt)
{
	int i, l, q;
	for (i = 0; i < m->N * m->K; i++) cnt[i] = 0;
	for (i = 0; i < m->N; i++)
		for (l = 0; l < m->L; l++)
			if (m->nd[(long)l * m->N + i] != 0)
				for (q = 0; q < 4; q++) cnt[(long)i * m->K + m->z[XO(m, l, i) + q]] += 1.0;
}

/* ---------------------------------------------------------------- dosage resolution --- */
/* two_allele_auto :2440 / tri_allele_auto :2509: write resolution `pick` (1..3) of the observed
 * allele set into g */
static void write_resolution(const int16_t *a, int nd, int pick, int8_t *g)
{
	if (nd == 2) {
		if (pick == 1) { g[0] = a[0]; g[1] = a[0]; g[2] = a[0]; g[3] = a[1]; }
		if (pick == 2) { g[0] = a[1]; g[1] = a[1]; g[2] = a[1]; g[3] = a[0]; }
		if (pick == 3) { g[0] = a[0]; g[1] = a[0]; g[2] = a[1]; g[3] = a[1]; }
	} else {
		if (pick == 1) { g[0] = a[0]; g[1] = a[0]; g[2] = a[1]; g[3] = a[2]; }
		if (pick == 2) { g[0] = a[1]; g[1] = a[1]; g[2] = a[0]; g[3] = a[2]; }
		if (pick == 3) { g[0] = a[2]; g[1] = a[2]; g[2] = a[0]; g[3] = a[1]; }
	}
}

/* log weights of the three resolutions: choose_two_auto :854-897, choose_tri_auto :907-952 */
static void resolution_logw(const tet_model *m, int i, int l, double *w)
{
	const int16_t *a = m->x + XO(m, l, i);
	const int8_t *z = m->z + XO(m, l, i);
	const int nd = m->nd[(long)l * m->N + i], n = m->allelenum[l];
	const tet_cat *c = &m->cat[m->cat_of[l]];
	int code[3], t, j;
	if (all_same4(z)) {
		const float *tab = m->genofreq + TO(m, z[0], l);
		if (nd == 2) {
			code[0] = a[0] * n * (n * n + n + 1) + a[1];
			code[1] = a[1] * n * (n * n + n + 1) + a[0];
			code[2] = (a[0] * n * n + a[1]) * (n + 1);
		} else {
			code[0] = a[0] * n * n * (n + 1) + a[1] * n + a[2];
			code[1] = a[1] * n * n * (n + 1) + a[0] * n + a[2];
			code[2] = a[2] * n * n * (n + 1) + a[0] * n + a[1];
		}
		for (t = 0; t < 3; t++) w[t] = (double)tab[lookup(c, code[t])];
	} else {
		double f[3];
		for (t = 0; t < nd; t++) {
			f[t] = 0;
			for (j = 0; j < m->K; j++) f[t] += m->qq[(long)i * m->K + j] * m->freq[FO(m, j, l, a[t])];
		}
		if (nd == 2) {
			w[0] = log(4) + 3 * log(f[0]) + log(f[1]);
			w[1] = log(4) + 3 * log(f[1]) + log(f[0]);
			w[2] = log(6) + 2 * log(f[0]) + 2 * log(f[1]);
		} else {
			w[0] = 2 * log(f[0]) + log(f[1]) + log(f[2]);
			w[1] = 2 * log(f[1]) + log(f[0]) + log(f[2]);
			w[2] = 2 * log(f[2]) + log(f[1]) + log(f[0]);
		}
	}
}

/* ---- allotetraploid resolutions: two_allele_allo :2465 (7 ways), tri_allele_allo :2533 (12),
 *      tetra_allele_allo :2602 (6).  ALLO_RES[nd-2][pick-1][copy] = index into the observed allele set */
static const int8_t ALLO_RES2[7][4] = {{0,0,0,1},{0,1,0,0},{0,0,1,1},{1,1,0,0},{0,1,1,1},{1,1,0,1},{0,1,0,1}};
static const int8_t ALLO_RES3[12][4] = {{0,0,1,2},{1,2,0,0},{1,1,0,2},{0,2,1,1},{2,2,0,1},{0,1,2,2},
                                        {0,1,1,2},{1,2,0,1},{1,2,0,2},{0,2,1,2},{0,2,0,1},{0,1,0,2}};
static const int8_t ALLO_RES4[6][4] = {{0,1,2,3},{2,3,0,1},{0,2,1,3},{1,3,0,2},{0,3,1,2},{1,2,0,3}};
static int allo_nres(int nd) { return nd == 2 ? 7 : (nd == 3 ? 12 : 6); }
static const int8_t *allo_res(int nd, int pick) { return nd == 2 ? ALLO_RES2[pick - 1] : (nd == 3 ? ALLO_RES3[pick - 1] : ALLO_RES4[pick - 1]); }

static void write_resolution_allo(const int16_t *a, int nd, int pick, int8_t *g)
{
	const int8_t *r = allo_res(nd, pick);
	int q;
	for (q = 0; q < 4; q++) g[q] = (int8_t)a[r[q]];
}

/* log weights of the resolutions: choose_two_allo :962, choose_tri_allo :1043, choose_tetra_allo :1144.
 * Same-z genotypes read population z's table at the resolution's genotype code (the codes the reference
 * writes out case by case are exactly the base-n numbers of the resolved genotypes); otherwise the product
 * of the admixture-averaged frequencies, subgenome 1 with freq and subgenome 2 with freq2, times 2 per
 * heterozygous pair -- except that the reference leaves the factor out when only ONE pair is heterozygous
 * (cases 1,2,5,6 of the two-allele list, 1-6 of the three-allele list) and in all six four-allele cases. */
static void resolution_logw_allo(const tet_model *m, int i, int l, double *w)
{
	const int16_t *a = m->x + XO(m, l, i);
	const int8_t *z = m->z + XO(m, l, i);
	const int nd = m->nd[(long)l * m->N + i], n = m->allelenum[l], nr = allo_nres(nd);
	const tet_cat *c = &m->cat[m->cat_of[l]];
	int t, j;
	if (all_same4(z)) {
		const float *tab = m->genofreq + TO(m, z[0], l);
		for (t = 0; t < nr; t++) {
			const int8_t *r = allo_res(nd, t + 1);
			const int code = ((a[r[0]] * n + a[r[1]]) * n + a[r[2]]) * n + a[r[3]];
			w[t] = (double)tab[lookup(c, code)];
		}
	} else {
		double f[4], f2[4];
		for (t = 0; t < nd; t++) {
			f[t] = 0; f2[t] = 0;
			for (j = 0; j < m->K; j++) {
				f[t] += m->qq[(long)i * m->K + j] * m->freq[FO(m, j, l, a[t])];
				f2[t] += m->qq[(long)i * m->K + j] * m->freq2[FO(m, j, l, a[t])];
			}
		}
		if (nd == 2) {
			w[0] = 2 * log(f[0]) + log(f2[0]) + log(f2[1]);
			w[1] = log(f[0]) + log(f[1]) + 2 * log(f2[0]);
			w[2] = 2 * log(f[0]) + 2 * log(f2[1]);
			w[3] = 2 * log(f[1]) + 2 * log(f2[0]);
			w[4] = log(f[0]) + log(f[1]) + 2 * log(f2[1]);
			w[5] = 2 * log(f[1]) + log(f2[0]) + log(f2[1]);
			w[6] = log(2) + log(f[0]) + log(f[1]) + log(f2[0]) + log(f2[1]);
		} else if (nd == 3) {
			w[0] = 2 * log(f[0]) + log(f2[1]) + log(f2[2]);
			w[1] = log(f[1]) + log(f[2]) + 2 * log(f2[0]);
			w[2] = 2 * log(f[1]) + log(f2[0]) + log(f2[2]);
			w[3] = log(f[0]) + log(f[2]) + 2 * log(f2[1]);
			w[4] = 2 * log(f[2]) + log(f2[0]) + log(f2[1]);
			w[5] = log(f[0]) + log(f[1]) + 2 * log(f2[2]);
			w[6] = log(2) + log(f[0]) + log(f[1]) + log(f2[1]) + log(f2[2]);
			w[7] = log(2) + log(f[1]) + log(f[2]) + log(f2[0]) + log(f2[1]);
			w[8] = log(2) + log(f[1]) + log(f[2]) + log(f2[0]) + log(f2[2]);
			w[9] = log(2) + log(f[0]) + log(f[2]) + log(f2[1]) + log(f2[2]);
			w[10] = log(2) + log(f[0]) + log(f[2]) + log(f2[0]) + log(f2[1]);
			w[11] = log(2) + log(f[0]) + log(f[1]) + log(f2[0]) + log(f2[2]);
		} else {
			w[0] = log(f[0]) + log(f[1]) + log(f2[2]) + log(f2[3]);
			w[1] = log(f[2]) + log(f[3]) + log(f2[0]) + log(f2[1]);
			w[2] = log(f[0]) + log(f[2]) + log(f2[1]) + log(f2[3]);
			w[3] = log(f[1]) + log(f[3]) + log(f2[0]) + log(f2[2]);
			w[4] = log(f[0]) + log(f[3]) + log(f2[1]) + log(f2[2]);
			w[5] = log(f[1]) + log(f[2]) + log(f2[0]) + log(f2[3]);
		}
	}
}

static int draw_resolution_allo(tet_model *m, int i, int l)
{
	double w[12], tm;
	const int nr = allo_nres(m->nd[(long)l * m->N + i]);
	int t;
	resolution_logw_allo(m, i, l, w);
	tm = w[0];
	for (t = 0; t < nr; t++) w[t] = exp(w[t] - tm);
	for (t = 1; t < nr; t++) w[t] += w[t - 1];
	return orc_rng_bracket(&m->rng, w, nr) + 1;
}

/* exact conditional of the allotetraploid resolution (tests); returns the number of resolutions */
int tet_geno_conditional_allo(const tet_model *m, int i, int l, double *prob)
{
	double w[12], s = 0, w0;
	const int nr = allo_nres(m->nd[(long)l * m->N + i]);
	int t;
	resolution_logw_allo(m, i, l, w);
	w0 = w[0];
	for (t = 0; t < nr; t++) { w[t] = exp(w[t] - w0); s += w[t]; }
	for (t = 0; t < nr; t++) prob[t] = w[t] / s;
	return nr;
}

void tet_geno_conditional(const tet_model *m, int i, int l, double *prob)
{
	double w[3], s = 0;
	int t;
	resolution_logw(m, i, l, w);
	for (t = 2; t >= 0; t--) w[t] = exp(w[t] - w[0]);
	for (t = 0; t < 3; t++) s += w[t];
	for (t = 0; t < 3; t++) prob[t] = w[t] / s;
}

static int draw_resolution(tet_model *m, int i, int l)
{
	double w[3], tm;
	int t;
	resolution_logw(m, i, l, w);
	tm = w[0];
	for (t = 0; t < 3; t++) w[t] = exp(w[t] - tm);
	for (t = 1; t < 3; t++) w[t] += w[t - 1];
	return orc_rng_bracket(&m->rng, w, 3) + 1;
}

/* choose_unif, poly_geno.c:842-852 */
static int draw_uniform_pick(tet_model *m, int n)
{
	double w[12];
	int j;
	for (j = 0; j < n; j++) w[j] = (double)(j + 1) / (double)n;
	return orc_rng_bracket(&m->rng, w, n) + 1;
}

/* initial_geno (:316-377) and update_geno (:520-580), autopolyploid branch */
static void resolve_all(tet_model *m, int initial)
{
	int i, l, q;
	for (i = 0; i < m->N; i++)
		for (l = 0; l < m->L; l++) {
			const int nd = m->nd[(long)l * m->N + i];
			const int16_t *a = m->x + XO(m, l, i);
			int8_t *g = m->geno + XO(m, l, i);
			if (nd == 0) continue;
			if (!m->autopoly) {                                 /* initial_geno :320-340, update_geno :524-548 */
				if (nd == 1) for (q = 0; q < 4; q++) g[q] = (int8_t)a[0];
				else write_resolution_allo(a, nd, initial ? draw_uniform_pick(m, allo_nres(nd)) : draw_resolution_allo(m, i, l), g);
				if (!initial) {                                 /* change_geno_allo :1475: each pair ascending */
					if (g[0] > g[1]) { int8_t t = g[0]; g[0] = g[1]; g[1] = t; }
					if (g[2] > g[3]) { int8_t t = g[2]; g[2] = g[3]; g[3] = t; }
				}
				continue;
			}
			if (nd == 1) for (q = 0; q < 4; q++) g[q] = (int8_t)a[0];
			else if (nd == 4) for (q = 0; q < 4; q++) g[q] = (int8_t)a[q];
			else write_resolution(a, nd, initial ? draw_uniform_pick(m, 3) : draw_resolution(m, i, l), g);
		}
}
void tet_initial_geno(tet_model *m) { resolve_all(m, 1); }
void tet_update_geno(tet_model *m) { resolve_all(m, 0); }

/* ---------------------------------------------------------------- conditional updates - */
/* update_P_auto, poly_geno.c:390-438 */
void tet_update_P(tet_model *m)
{
	int32_t *n = (int32_t *)malloc(sizeof(int32_t) * (size_t)m->K * m->L * m->Amax);
	double *cnt = (double *)malloc(sizeof(double) * m->Amax);
	int k, l, a;
	if (!m->autopoly) {                                        /* update_P_allo, poly_geno.c:441-518 */
		int32_t *n2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)m->K * m->L * m->Amax);
		tet_tally_allo(m, n, n2);
		for (k = 0; k < m->K; k++)
			for (l = 0; l < m->L; l++) {
				for (a = 0; a < m->allelenum[l]; a++) cnt[a] = (double)n[FO(m, k, l, a)];
				orc_rng_dirichlet(&m->rng, cnt, m->allelenum[l], m->freq + FO(m, k, l, 0), 1.0);
				for (a = 0; a < m->allelenum[l]; a++) cnt[a] = (double)n2[FO(m, k, l, a)];
				orc_rng_dirichlet(&m->rng, cnt, m->allelenum[l], m->freq2 + FO(m, k, l, 0), 1.0);
			}
		free(n); free(n2); free(cnt);
		return;
	}
	tet_tally(m, n);
	for (k = 0; k < m->K; k++)
		for (l = 0; l < m->L; l++) {
			for (a = 0; a < m->allelenum[l]; a++) cnt[a] = (double)n[FO(m, k, l, a)];
			orc_rng_dirichlet(&m->rng, cnt, m->allelenum[l], m->freq + FO(m, k, l, 0), 1.0);
		}
	free(n); free(cnt);
}

/* update_S_POP, poly_geno.c:584-643 */
void tet_update_S(tet_model *m)
{
	const double delta0 = 0.05;
	float *tab = (float *)malloc(sizeof(float) * (size_t)m->L * m->Gmax);
	int j, st = 0;
	for (j = 0; j < m->K; j++) tet_calc_genofreq(m, j, m->self_rates[j], m->genofreq + TO(m, j, 0));
	for (j = 0; j < m->K; j++) {
		double prop = 0, mh;
		if (m->back_refl == 1) {
			prop = orc_rng_u01(&m->rng) * 2 * delta0 - delta0;
			prop += m->self_rates[j];
			if (prop <= 0.000) prop = 0.000 - prop;
			else if (prop >= 1.000) prop = 1.000 - (prop - 1.000);
		} else prop = orc_rng_three_state(&m->rng, &st, m->state[j]);
		tet_calc_genofreq(m, j, prop, tab);
		mh = tet_cal_lkd_props(m, j, tab) - tet_cal_lkd(m);
		/* as written (:627): the Hastings factor multiplies the LOG ratio */
		if (m->back_refl == 0) mh *= orc_trans_prob(m->state[j], st) / orc_trans_prob(st, m->state[j]);
		if (orc_rng_u01(&m->rng) < exp(MIN2(0, mh))) {
			m->self_rates[j] = prop;
			if (m->back_refl == 0) m->state[j] = st;
			memcpy(m->genofreq + TO(m, j, 0), tab, sizeof(float) * (size_t)m->L * m->Gmax);
		}
	}
	free(tab);
}

void tet_z_conditional(const tet_model *m, int i, int l, int c, double *prob)
{
	double s = 0;
	int k;
	for (k = 0; k < m->K; k++) { prob[k] = m->qq[(long)i * m->K + k] * m->freq[FO(m, k, l, m->geno[XO(m, l, i) + c])]; s += prob[k]; }
	for (k = 0; k < m->K; k++) prob[k] /= s;
}

/* update_ZQ, poly_geno.c:750-836 */
void tet_update_ZQ(tet_model *m, int init_flag)
{
	double *w = (double *)malloc(sizeof(double) * m->K);
	int i, l, q, k;
	for (i = 0; i < m->N; i++) {
		for (l = 0; l < m->L; l++) {
			if (m->nd[(long)l * m->N + i] == 0) continue;
			for (q = 0; q < 4; q++) {
				for (k = 0; k < m->K; k++) {
					if (init_flag == 1) w[k] = (double)(k + 1) / m->K;
					else {
						w[k] = m->qq[(long)i * m->K + k] * m->freq[FO(m, k, l, m->geno[XO(m, l, i) + q])];
						if (k >= 1) w[k] += w[k - 1];
					}
				}
				m->z[XO(m, l, i) + q] = (int8_t)orc_rng_bracket(&m->rng, w, m->K);
			}
		}
		for (k = 0; k < m->K; k++) m->qqnum[(long)i * m->K + k] = 0.0;
		for (l = 0; l < m->L; l++)
			if (m->nd[(long)l * m->N + i] != 0)
				for (q = 0; q < 4; q++) m->qqnum[(long)i * m->K + m->z[XO(m, l, i) + q]] += 1.0;
		for (k = 0; k < m->K; k++) w[k] = m->qqnum[(long)i * m->K + k];
		orc_rng_dirichlet(&m->rng, w, m->K, m->qq + (long)i * m->K, m->alpha);
	}
	free(w);
}

/* one pass of the for(step) body, poly_geno.c:97-112 */
static void one_sweep(tet_model *m)
{
	tet_update_P(m);
	tet_calc_exfreq(m);
	tet_update_S(m);
	tet_update_ZQ(m, 0);
	tet_update_geno(m);
	m->totallkh = tet_cal_lkd(m);
}
void tet_sweeps(tet_model *m, int n) { while (n-- > 0) one_sweep(m); }

/* ---------------------------------------------------------------- chain ---------------- */
orc_chain *tet_chain_new(const tet_model *m, int ckrep)
{
	orc_chain *c = (orc_chain *)calloc(1, sizeof(orc_chain));
	c->indvlkh = (double *)calloc(m->N, sizeof(double));
	c->qq = (double *)calloc((size_t)m->N * m->K, sizeof(double));
	c->qq2 = (double *)calloc((size_t)m->N * m->K, sizeof(double));
	c->self_rates = (double *)calloc(m->K, sizeof(double));
	c->self_rates2 = (double *)calloc(m->K, sizeof(double));
	c->gen = (double *)calloc(1, sizeof(double));
	c->gen2 = (double *)calloc(1, sizeof(double));
	c->convg = (double *)calloc(ckrep > 0 ? ckrep : 1, sizeof(double));
	return c;
}

/* store_chn, mcmc.c:1320-1456 (ploid 4 stores totallkh, indvlkh, qq, self_rates) */
static inline double run_mean(double mean, double xv, long step)
{
	if (mean != 0) return mean * ((step + xv / mean) / (1 + step));
	return xv / (1 + step);
}
static void store(const tet_model *m, orc_chain *c)
{
	int i, k;
	c->totallkh = run_mean(c->totallkh, m->totallkh, c->step);
	c->totallkh2 = run_mean(c->totallkh2, m->totallkh * m->totallkh, c->step);
	for (i = 0; i < m->N; i++) c->indvlkh[i] = run_mean(c->indvlkh[i], m->indvlkh[i], c->step);
	for (i = 0; i < m->N * m->K; i++) {
		c->qq[i] = run_mean(c->qq[i], m->qq[i], c->step);
		c->qq2[i] = run_mean(c->qq2[i], m->qq[i] * m->qq[i], c->step);
	}
	for (k = 0; k < m->K; k++) {
		c->self_rates[k] = run_mean(c->self_rates[k], m->self_rates[k], c->step);
		c->self_rates2[k] = run_mean(c->self_rates2[k], m->self_rates[k] * m->self_rates[k], c->step);
	}
	c->step++;
}

/* mcmc_POP_tetra_selfing, poly_geno.c:75-140 */
int tet_run_chain(tet_model *m, long update, long burnin, int thinning, int ckrep, int nstep_check_empty,
                  const float *initd, orc_chain *out)
{
	long step, cnt_step = 0;
	int i, k;
	m->alpha = orc_rng_u01(&m->rng) * 10;                    /* initial_chn, :379-388 */
	out->steps = (int)((update - burnin) / thinning);
	tet_initial_geno(m);
	for (k = 0; k < m->K; k++) {
		m->self_rates[k] = initd ? initd[k] : 0.5f;
		if (m->back_refl == 0) m->state[k] = orc_dt_stat(m->self_rates[k]);
	}
	tet_update_ZQ(m, 1);
	for (step = 0; step < update; step++) {
		one_sweep(m);
		if (step == burnin - 1) {                               /* allocate_chn + initialize_chn, mcmc.c:588-738 */
			out->step = 0; out->flag_empty_cluster = 0;
			out->totallkh = out->totallkh2 = 1.0;
			for (i = 0; i < m->N; i++) out->indvlkh[i] = 1.0;
			for (i = 0; i < m->N * m->K; i++) { out->qq[i] = 1.0; out->qq2[i] = 1.0; }
			for (k = 0; k < m->K; k++) { out->self_rates[k] = 1.0; out->self_rates2[k] = 1.0; }
		}
		if (step >= burnin && (step + 1 - burnin) % thinning == 0) {
			store(m, out);
			if (cnt_step < ckrep) out->convg[cnt_step] = m->totallkh;
			cnt_step++;
		}
		if (cnt_step == nstep_check_empty) {
			int empty = 0;
			for (k = 0; k < m->K; k++) {                        /* check_empty_cluster, mcmc.c:1944-1974 */
				double s = 0;
				for (i = 0; i < m->N; i++) s += m->qq[(long)i * m->K + k];
				if (s < 0.01) empty = 1;
			}
			out->flag_empty_cluster = empty;
			if (empty) return 1;
		}
	}
	return 0;
}
