/*
 * oracle/instruct_oracle.c -- CPU restatement of the reference's diploid per-sweep hot path
 * (modes 2 and 3 of slowkoni/InStruct).  See instruct_oracle.h for status and pinning.
 *
 * TEST INFRASTRUCTURE ONLY -- never imported, linked or executed by the product path.
 *
 * Every function cites the reference lines it restates.  The arithmetic (operation order,
 * libm calls, truncated constants such as E=2.71828182 and PI=3.141592654) follows the
 * reference so that chains started from the same Wichmann-Hill seeds agree bit for bit;
 * the data structures (flat packed arrays, index-linked cluster pool) are this repo's own.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "instruct_oracle.h"

#define ORC_E  2.71828182       /* random.c:7 */
#define ORC_PI 3.141592654      /* random.c:8, mcmc.c:26 */
/* mcmc.h:10 -- note MIN2(1, NaN) == 1, so a NaN ratio accepts */
#define MIN2(X, Y) (((X) > (Y)) ? (Y) : (X))

struct orc_model {
	int N, L, K, ploid, Amax;
	int mode, prior_flag, back_refl, type_freq;
	double alpha_dpm;
	const int16_t *x;
	const int32_t *allelenum;
	int8_t *z;
	int *zz;              /* mode 0: UPMCMC.zz, the cluster of each individual (mcmc.c:539) */
	double *qq, *qqnum, *freq;
	double alpha;
	double *self_rates;
	int *state;
	int *gen;
	double *indvlkh;
	double totallkh;
	orc_rng rng;
	/* DP prior: value-sorted singly linked list held in a pool (DPMM.h:10-23) */
	int dp_head, dp_free, dp_cnt;
	double *dp_value;
	int *dp_num, *dp_next;
	int *dp_of;           /* cluster slot of each individual */
	double *dp_sval;      /* indv_array[j].value */
	double *scratch;      /* K+N+2 doubles */
};

/* ------------------------------------------------------------------ helpers ---------- */
static inline long XO(const orc_model *m, int l, int i, int c) { return ((long)l * m->N + i) * m->ploid + c; }
static inline long FO(const orc_model *m, int k, int l, int a) { return ((long)k * m->L + l) * m->Amax + a; }

/* usable genotype: no copy missing and the locus polymorphic (mcmc.c:817,1137,1180,1737;
 * missindx rule data_interface.c:820-832) */
static inline int usable(const orc_model *m, int i, int l)
{
	int c;
	if (m->allelenum[l] <= 1) return 0;
	for (c = 0; c < m->ploid; c++) if (m->x[XO(m, l, i, c)] < 0) return 0;
	return 1;
}

orc_model *orc_new(int N, int L, int K, int ploid, int mode, int prior_flag, int back_refl,
                   int type_freq, double alpha_dpm, const int16_t *x, const int32_t *allelenum)
{
	int l, i, k, ns;
	orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
	m->N = N; m->L = L; m->K = K; m->ploid = ploid; m->mode = mode; m->prior_flag = prior_flag;
	m->back_refl = back_refl; m->type_freq = type_freq; m->alpha_dpm = alpha_dpm;
	m->x = x; m->allelenum = allelenum;
	m->Amax = 1;
	for (l = 0; l < L; l++) if (allelenum[l] > m->Amax) m->Amax = allelenum[l];
	m->z = (int8_t *)calloc((size_t)L * N * ploid, 1);
	m->zz = (int *)calloc(N, sizeof(int));
	m->qq = (double *)calloc((size_t)N * K, sizeof(double));
	m->qqnum = (double *)calloc((size_t)N * K, sizeof(double));
	m->freq = (double *)calloc((size_t)K * L * m->Amax, sizeof(double));
	ns = (mode == 3 || mode == 5) ? N : K;   /* modes 4/5: self_rates holds UPMCMC.inbreed (mcmc.c:521-522) */
	m->self_rates = (double *)calloc(ns, sizeof(double));
	m->state = (int *)calloc(K, sizeof(int));
	m->gen = (int *)calloc(N, sizeof(int));
	m->indvlkh = (double *)calloc(N, sizeof(double));
	m->alpha = 1.0;
	for (i = 0; i < N; i++) { m->gen[i] = 1; for (k = 0; k < K; k++) m->qq[(long)i * K + k] = 1.0 / K; }
	m->rng.s1 = 13; m->rng.s2 = 4; m->rng.s3 = 1972;       /* random.c:10-12 */
	m->dp_head = -1; m->dp_free = -1; m->dp_cnt = 0;
	m->dp_value = (double *)calloc(N + 1, sizeof(double));
	m->dp_num = (int *)calloc(N + 1, sizeof(int));
	m->dp_next = (int *)calloc(N + 1, sizeof(int));
	m->dp_of = (int *)calloc(N + 1, sizeof(int));
	m->dp_sval = (double *)calloc(N + 1, sizeof(double));
	m->scratch = (double *)calloc((size_t)K + N + 4, sizeof(double));
	return m;
}

void orc_free(orc_model *m)
{
	if (!m) return;
	free(m->z); free(m->zz); free(m->qq); free(m->qqnum); free(m->freq); free(m->self_rates); free(m->state);
	free(m->gen); free(m->indvlkh); free(m->dp_value); free(m->dp_num); free(m->dp_next);
	free(m->dp_of); free(m->dp_sval); free(m->scratch); free(m);
}

int8_t *orc_z(orc_model *m) { return m->z; }
int *orc_zz(orc_model *m) { return m->zz; }
double *orc_qq(orc_model *m) { return m->qq; }
double *orc_qqnum(orc_model *m) { return m->qqnum; }
double *orc_freq(orc_model *m) { return m->freq; }
double *orc_self(orc_model *m) { return m->self_rates; }
int *orc_state(orc_model *m) { return m->state; }
int *orc_gen(orc_model *m) { return m->gen; }
double *orc_indvlkh(orc_model *m) { return m->indvlkh; }
double *orc_alpha(orc_model *m) { return &m->alpha; }
double *orc_totallkh(orc_model *m) { return &m->totallkh; }
int orc_amax(orc_model *m) { return m->Amax; }
void orc_setseeds(orc_model *m, long a, long b, long c) { m->rng.s1 = a; m->rng.s2 = b; m->rng.s3 = c; }
void orc_getseeds(orc_model *m, long *o) { o[0] = m->rng.s1; o[1] = m->rng.s2; o[2] = m->rng.s3; }

/* ------------------------------------------------------------------ RNG (random.c) --- */

/* ran1()/wichmann(), random.c:19-47: three 16-bit LCGs, fractional part of the sum */
static double u01(orc_rng *r)
{
	r->s1 = (171 * r->s1) % 30269;
	r->s2 = (172 * r->s2) % 30307;
	r->s3 = (170 * r->s3) % 30323;
	return fmod(r->s1 / 30269.0 + r->s2 / 30307.0 + r->s3 / 30323.0, 1.0);
}

/* rexp, random.c:121-130 */
static double draw_exp(orc_rng *r, double lambda)
{
	double u = u01(r);
	return -(1 / lambda) * log(u);
}

/* rgamma1, random.c:167-193: shape < 1, one accept/reject attempt, -1 on reject */
static double gamma_small_try(orc_rng *r, double a)
{
	double u0 = u01(r), u1 = u01(r), v, t;
	if (u0 > ORC_E / (a + ORC_E)) {
		v = -log((a + ORC_E) * (1 - u0) / (a * ORC_E));
		return (u1 > pow(v, a - 1)) ? -1 : v;
	}
	t = (a + ORC_E) * u0 / ORC_E;
	v = pow(t, 1 / a);
	return (u1 > exp(-v)) ? -1 : v;
}

/* rgamma2, random.c:195-231: shape > 1 (Cheng-Feast style ratio method), -1 on reject */
static double gamma_large_try(orc_rng *r, double a)
{
	double c1 = a - 1, c2 = (a - 1 / (6 * a)) / c1, c3 = 2 / c1, c4 = c3 + 2, c5 = 1 / sqrt(a);
	double u1, u2, w;
	do {
		u1 = u01(r);
		u2 = u01(r);
		if (a > 2.5) u1 = u2 + c5 * (1 - 1.86 * u1);
	} while (u1 >= 1 || u1 <= 0);
	w = c2 * u2 / u1;
	if (c3 * u1 + w + 1 / w > c4)
		if (c3 * log(u1) - log(w) + w >= 1) return -1;
	return c1 * w;
}

/* rgamma, random.c:233-250 */
static double draw_gamma(orc_rng *r, double a, double b)
{
	double v = 0;
	if (a < 1) do { v = gamma_small_try(r, a) / b; } while (v < 0);
	if (a == 1) v = draw_exp(r, 1) / b;
	if (a > 1) do { v = gamma_large_try(r, a) / b; } while (v < 0);
	return v;
}

/* rdirich, random.c:264-280: out = normalised Gamma(alpha[k]+add, 1) */
static void draw_dirichlet(orc_rng *r, const double *alpha, int len, double *out, double add)
{
	double sum = 0;
	int k;
	for (k = 0; k < len; k++) { out[k] = draw_gamma(r, alpha[k] + add, 1.0); sum += out[k]; }
	for (k = 0; k < len; k++) out[k] /= sum;
}

/* rbeta, random.c:252-261 */
static double draw_beta(orc_rng *r, double a, double b)
{
	double g = draw_gamma(r, a, 1.0);
	return g / (g + draw_gamma(r, b, 1.0));
}

/* rstd_normal + rnormal, random.c:283-307 (Box-Muller, cosine branch only) */
static double draw_normal(orc_rng *r, double mu, double sd)
{
	double u1 = u01(r), u2 = u01(r);
	double theta = 2 * ORC_PI * u1, rad = sqrt(2 * (-log(u2)));
	return mu + sd * (rad * cos(theta));
}

/* rgeom, random.c:311-321 */
static int draw_geom(orc_rng *r, double p)
{
	double u = u01(r);
	return (int)(log(u) / log(1 - p)) + 1;
}

/* disc_unif, random.c:403-430: draws FIRST, then normalises the running sums in place
 * (the last element divides itself to 1), then returns the bracket holding the draw. */
static int draw_bracket(orc_rng *r, double *vec, int len)
{
	int i, pick = 0;
	double u = u01(r);
	for (i = 0; i < len; i++) vec[i] /= vec[len - 1];
	if (u < 0.0 || u > vec[len - 1]) { fprintf(stderr, "oracle: draw outside interval\n"); exit(1); }
	if (!(u <= vec[0] && u >= 0.0))
		for (i = 1; i < len; i++) if (u > vec[i - 1] && u <= vec[i]) pick = i;
	return pick;
}

/* the same samplers on a bare stream, for oracle/tetra_oracle.c */
double orc_rng_u01(orc_rng *r) { return u01(r); }
void orc_rng_dirichlet(orc_rng *r, const double *alpha, int len, double *out, double add) { draw_dirichlet(r, alpha, len, out, add); }
int orc_rng_bracket(orc_rng *r, double *vec, int len) { return draw_bracket(r, vec, len); }

double orc_ran1(orc_model *m) { return u01(&m->rng); }
double orc_rgamma(orc_model *m, double a, double b) { return draw_gamma(&m->rng, a, b); }
double orc_rbeta(orc_model *m, double a, double b) { return draw_beta(&m->rng, a, b); }
double orc_rnormal(orc_model *m, double mu, double sd) { return draw_normal(&m->rng, mu, sd); }
int orc_rgeom(orc_model *m, double p) { return draw_geom(&m->rng, p); }
int orc_disc_unif(orc_model *m, double *vec, int len) { return draw_bracket(&m->rng, vec, len); }

/* ------------------------------------------------------------------ pure pieces ------ */

void orc_missing_mask(const orc_model *m, uint8_t *mask)
{
	/* missindx[i][l] = 1 if ANY copy is missing, data_interface.c:822-835 */
	int l, i, c;
	for (l = 0; l < m->L; l++) for (i = 0; i < m->N; i++) {
		uint8_t miss = 0;
		for (c = 0; c < m->ploid; c++) if (m->x[XO(m, l, i, c)] < 0) miss = 1;
		mask[(long)l * m->N + i] = miss;
	}
}

/* tally half of update_P, mcmc.c:810-845: n[k][l][a] over usable genotypes */
void orc_tally_range(const orc_model *m, int i0, int i1, int32_t *n)
{
	int l, i, c;
	memset(n, 0, (size_t)m->K * m->L * m->Amax * sizeof(int32_t));
	for (l = 0; l < m->L; l++)
		for (i = i0; i < i1; i++)
			if (usable(m, i, l))
				for (c = 0; c < m->ploid; c++)
					n[FO(m, m->mode == 0 ? m->zz[i] : m->z[XO(m, l, i, c)], l, m->x[XO(m, l, i, c)])]++;   /* mcmc.c:825-831: mode 0 counts by zz */
}
void orc_tally(const orc_model *m, int32_t *n) { orc_tally_range(m, 0, m->N, n); }

/* per-individual ancestry counts, mcmc.c:1176-1194 */
void orc_count_z(const orc_model *m, double *cnt)
{
	int i, l, c, k;
	for (i = 0; i < m->N; i++) {
		for (k = 0; k < m->K; k++) cnt[(long)i * m->K + k] = 0.0;
		for (l = 0; l < m->L; l++)
			if (usable(m, i, l))
				for (c = 0; c < m->ploid; c++) cnt[(long)i * m->K + m->z[XO(m, l, i, c)]] += 1.0;
	}
}

/* genofreq, mcmc.c:1683-1703 (diploid) */
double orc_genofreq(int a0, int a1, double f0, double f1, int gen)
{
	double res, t;
	int g;
	if (a0 == a1) {
		res = pow(f0, 2.0);
		t = 2 * f0 * (1 - f0);
		for (g = 1; g < gen; g++) { t /= 2; res += t / 2; }
		return res;
	}
	return 2 * f0 * f1 * pow(0.5, (double)(gen - 1));
}

/* log_ld_indv, mcmc.c:1726-1773 */
double orc_log_ld_indv(const orc_model *m, int gen, int i)
{
	double ll = 0, f[2];
	int l, c, k;
	for (l = 0; l < m->L; l++) {
		int a0, a1, z0, z1;
		if (!usable(m, i, l)) continue;
		a0 = m->x[XO(m, l, i, 0)]; a1 = m->x[XO(m, l, i, 1)];
		z0 = m->z[XO(m, l, i, 0)]; z1 = m->z[XO(m, l, i, 1)];
		if (m->type_freq == 0) {                       /* :1739-1749 expectation over Q */
			for (c = 0; c < 2; c++) {
				int a = c ? a1 : a0;
				f[c] = 0;
				for (k = 0; k < m->K; k++) f[c] += m->freq[FO(m, k, l, a)] * m->qq[(long)i * m->K + k];
			}
			ll += log(orc_genofreq(a0, a1, f[0], f[1], gen));
		}
		if (m->type_freq == 1) {                       /* :1750-1768 conditional on Z */
			f[0] = m->freq[FO(m, z0, l, a0)]; f[1] = m->freq[FO(m, z1, l, a1)];
			if (z0 == z1) ll += log(orc_genofreq(a0, a1, f[0], f[1], gen));
			else {
				ll += log(f[0]);
				ll += log(f[1]);
				if (a0 != a1) ll += log(2);
			}
		}
	}
	return ll;
}

/* proposal, mcmc.c:1630-1648 */
double orc_proposal(const orc_model *m, const double *S)
{
	double ld = 0;
	int i, k;
	for (i = 0; i < m->N; i++) {
		double s = 0;
		for (k = 0; k < m->K; k++) s += m->qq[(long)i * m->K + k] * S[k];
		ld += log(pow(s, m->gen[i] - 1) * (1 - s));
	}
	return ld;
}

/* genofreq_inbreedcoff, mcmc.c:1707-1723 (diploid) */
double orc_genofreq_F(int a0, int a1, double f0, double f1, double F)
{
	if (a0 == a1) return pow(f0, 2.0) * (1 - F) + f0 * F;
	return 2 * f0 * f1 * (1 - F);
}

/* log_ld_F_pop (mcmc.c:1776-1809, by_pop = 1: F = inbreed[z0]) and log_ld_F_indv
 * (mcmc.c:1812-1845, by_pop = 0: F = inbreed[0]) */
double orc_log_ld_F(const orc_model *m, const double *inbreed, int by_pop, int i)
{
	double ll = 0;
	int l;
	for (l = 0; l < m->L; l++) {
		int a0, a1, z0, z1;
		double f0, f1;
		if (!usable(m, i, l)) continue;
		a0 = m->x[XO(m, l, i, 0)]; a1 = m->x[XO(m, l, i, 1)];
		z0 = m->z[XO(m, l, i, 0)]; z1 = m->z[XO(m, l, i, 1)];
		f0 = m->freq[FO(m, z0, l, a0)]; f1 = m->freq[FO(m, z1, l, a1)];
		if (z0 == z1) ll += log(orc_genofreq_F(a0, a1, f0, f1, by_pop ? inbreed[z0] : inbreed[0]));
		else {
			ll += log(f0);
			ll += log(f1);
			if (a0 != a1) ll += log(2);
		}
	}
	return ll;
}

/* log_ld_F_total, mcmc.c:1849-1866 */
double orc_log_ld_F_total(const orc_model *m, const double *inbreed)
{
	double ld = 0;
	int i;
	if (m->mode == 4) for (i = 0; i < m->N; i++) ld += orc_log_ld_F(m, inbreed, 1, i);
	if (m->mode == 5) for (i = 0; i < m->N; i++) ld += orc_log_ld_F(m, inbreed + i, 0, i);
	return ld;
}

/* dgeom, mcmc.c:1596-1604 */
double orc_dgeom(double s, int g) { return pow(s, (double)(g - 1)) * (1 - s); }

/* dt_stat, mcmc.c:1524-1546; returns -1 where the reference exits */
int orc_dt_stat(double v)
{
	const double eps = 0.001;
	if (v <= 0.0 + eps && v >= 0.0 - eps) return 0;
	if (v >= 1.0 - eps && v <= 1.0 + eps) return 2;
	if (v >= 0.0 + eps && v < 1.0 - eps) return 1;
	return -1;
}

/* update_alpha's ratio, mcmc.c:1254-1260, as written (running product of pow ratios) */
double orc_alpha_ratio_product(const orc_model *m, double ralpha)
{
	double r = 1.0;
	long j;
	for (j = 0; j < (long)m->N * m->K; j++)
		r *= pow(m->qq[j], ralpha + m->qqnum[j]) / pow(m->qq[j], m->qqnum[j] + m->alpha);
	return r;
}
/* ... and its logarithm, (alpha' - alpha) * sum log q, which is what the device evaluates */
double orc_alpha_logratio(const orc_model *m, double ralpha)
{
	double s = 0;
	long j;
	for (j = 0; j < (long)m->N * m->K; j++) s += log(m->qq[j]);
	return (ralpha - m->alpha) * s;
}

/* check_empty_cluster, mcmc.c:1944-1974 */
int orc_check_empty_cluster(const orc_model *m)
{
	int k, i;
	for (k = 0; k < m->K; k++) {
		double s = 0;
		for (i = 0; i < m->N; i++) s += m->qq[(long)i * m->K + k];
		if (s < 0.01) return 1;
	}
	return 0;
}

/* exact conditional P(z_ilc = k) used by update_ZQ, mcmc.c:1141-1149 (for chi-square tests) */
void orc_z_conditional(const orc_model *m, int i, int l, int c, double *prob)
{
	double tot = 0;
	int k, a = m->x[XO(m, l, i, c)];
	for (k = 0; k < m->K; k++) { prob[k] = m->qq[(long)i * m->K + k] * m->freq[FO(m, k, l, a)]; tot += prob[k]; }
	for (k = 0; k < m->K; k++) prob[k] /= tot;
}

/* ------------------------------------------------------------------ updates ---------- */

/* update_P, mcmc.c:799-861 */
void orc_update_P(orc_model *m)
{
	int32_t *n = (int32_t *)malloc((size_t)m->K * m->L * m->Amax * sizeof(int32_t));
	double *cnt = (double *)malloc(m->Amax * sizeof(double));
	int k, l, a;
	orc_tally(m, n);
	for (k = 0; k < m->K; k++)
		for (l = 0; l < m->L; l++)
			if (m->allelenum[l] > 1) {
				for (a = 0; a < m->allelenum[l]; a++) cnt[a] = (double)n[FO(m, k, l, a)];
				draw_dirichlet(&m->rng, cnt, m->allelenum[l], &m->freq[FO(m, k, l, 0)], 1.0);
			}
	free(n); free(cnt);
}

/* adpt_indp, mcmc.c:1461-1520: 3-state adaptive independence proposal (-e 0) */
static double propose_three_state(orc_rng *r, int *new_state, int cur)
{
	double t;
	if (cur == 0) {
		if (u01(r) < 0.50) { *new_state = 0; return 0.0; }
		*new_state = 1; return u01(r);
	}
	if (cur == 2) {
		if (u01(r) < 0.5) { *new_state = 2; return 1.0; }
		*new_state = 1; return u01(r);
	}
	t = u01(r);
	if (t <= 0.05) { *new_state = 0; return 0.0; }
	if (t >= 0.95) { *new_state = 2; return 1.0; }
	*new_state = 1; return u01(r);
}
/* q(), mcmc.c:1566-1593 */
static double trans_prob(int a, int b)
{
	if (a == 0) return (b == 0 || b == 1) ? 0.5 : 0.0;
	if (a == 2) return (b == 2 || b == 1) ? 0.5 : 0.0;
	if (a == 1) return (b == 1) ? 0.90 : 0.05;
	return 0.0;
}

double orc_rng_three_state(orc_rng *r, int *new_state, int cur) { return propose_three_state(r, new_state, cur); }
double orc_trans_prob(int a, int b) { return trans_prob(a, b); }

/* update_S_POP, mcmc.c:913-983 */
void orc_update_S_POP(orc_model *m)
{
	const double delta0 = 0.05;
	double *tmp = m->scratch;
	int j, i, st = 0;
	for (j = 0; j < m->K; j++) {
		double mh;
		for (i = 0; i < m->K; i++) tmp[i] = m->self_rates[i];
		if (m->back_refl == 1) {
			tmp[j] = u01(&m->rng) * 2 * delta0 - delta0;
			tmp[j] += m->self_rates[j];
			if (tmp[j] <= 0.0) tmp[j] = 0.0 - tmp[j];
			else if (tmp[j] >= 1.0) tmp[j] = 1.0 - (tmp[j] - 1.0);
		} else {
			tmp[j] = propose_three_state(&m->rng, &st, m->state[j]);
		}
		mh = exp(orc_proposal(m, tmp) - orc_proposal(m, m->self_rates));
		if (m->back_refl == 0)       /* hastings_stat, mcmc.c:1550-1563: only term j differs from 1 */
			mh *= trans_prob(m->state[j], st) / trans_prob(st, m->state[j]);
		if (u01(&m->rng) < MIN2(1, mh)) {
			m->self_rates[j] = tmp[j];
			if (m->back_refl == 0) m->state[j] = st;
		}
	}
}

/* update_inbreedcoff_POP, mcmc.c:986-1050.  The difference of log-likelihoods is used as it
 * stands there: multiplied (not added to) by the Hastings ratio under -e 0 and accepted when
 * u < exp(min(1, .)), which for u < 1 is the usual min(1, ratio) rule. */
void orc_update_F_POP(orc_model *m)
{
	const double delta0 = 0.05;
	double *tmp = m->scratch;
	int j, i, st = 0;
	for (j = 0; j < m->K; j++) {
		double mh;
		for (i = 0; i < m->K; i++) tmp[i] = m->self_rates[i];
		if (m->back_refl == 1) {
			tmp[j] = u01(&m->rng) * 2 * delta0 - delta0;
			tmp[j] += m->self_rates[j];
			if (tmp[j] <= 0.000) tmp[j] = 0.000 - tmp[j];
			else if (tmp[j] >= 1.000) tmp[j] = 1.000 - (tmp[j] - 1.000);
		} else {
			tmp[j] = propose_three_state(&m->rng, &st, m->state[j]);
		}
		mh = orc_log_ld_F_total(m, tmp) - orc_log_ld_F_total(m, m->self_rates);
		if (m->back_refl == 0) mh *= trans_prob(m->state[j], st) / trans_prob(st, m->state[j]);
		if (u01(&m->rng) < exp(MIN2(1, mh))) {
			m->self_rates[j] = tmp[j];
			if (m->back_refl == 0) m->state[j] = st;
		}
	}
}

/* update_F_IND, mcmc.c:888-910 (two independent reflections, unlike update_inbreedcoff_POP) */
void orc_update_F_IND(orc_model *m)
{
	const double delta0 = 0.05;
	int j;
	for (j = 0; j < m->N; j++) {
		double t = u01(&m->rng) * 2 * delta0 - delta0, mh;
		t += m->self_rates[j];
		if (t <= 0.0) t = 0.0 - t;
		if (t >= 1.0) t = 1.0 - (t - 1);
		mh = exp(orc_log_ld_F(m, &t, 0, j) - orc_log_ld_F(m, m->self_rates + j, 0, j));
		m->self_rates[j] = (u01(&m->rng) < MIN2(1, mh)) ? t : m->self_rates[j];
	}
}

/* update_S_IND, mcmc.c:864-886 */
void orc_update_S_IND(orc_model *m)
{
	const double delta0 = 0.05;
	int j;
	for (j = 0; j < m->N; j++) {
		double t = u01(&m->rng) * 2 * delta0 - delta0, mh;
		t += m->self_rates[j];
		if (t <= 0.0) t = 0.0 - t;
		if (t >= 1.0) t = 1.0 - (t - 1);
		mh = exp(log(orc_dgeom(t, m->gen[j])) - log(orc_dgeom(m->self_rates[j], m->gen[j])));
		m->self_rates[j] = (u01(&m->rng) < MIN2(1, mh)) ? t : m->self_rates[j];
	}
}

/* update_G, mcmc.c:1053-1091 */
void orc_update_G(orc_model *m)
{
	int i, k;
	for (i = 0; i < m->N; i++) {
		double s = 0, mh;
		int st, g = 0;
		if (m->mode == 2) for (k = 0; k < m->K; k++) s += m->qq[(long)i * m->K + k] * m->self_rates[k];
		if (m->mode == 3) s = m->self_rates[i];
		st = orc_dt_stat(s);
		if (st == 1) {
			g = draw_geom(&m->rng, 1 - s);
			if (g < 1) g = 1;
			if (g > 50) g = 50;
		} else if (st == 0) g = 1;
		else if (st == 2) g = 50;
		else { fprintf(stderr, "oracle: selfing rate %f outside [0,1]\n", s); exit(1); }
		mh = exp(orc_log_ld_indv(m, g, i) - orc_log_ld_indv(m, m->gen[i], i));
		if (u01(&m->rng) < MIN2(1, mh)) m->gen[i] = g;
	}
}

/* update_ZQ, mcmc.c:1122-1203 */
void orc_update_ZQ(orc_model *m, int init_flag)
{
	double *w = m->scratch;
	int i, l, c, k;
	for (i = 0; i < m->N; i++) {
		double *q = &m->qq[(long)i * m->K], *cnt = &m->qqnum[(long)i * m->K];
		for (l = 0; l < m->L; l++) {
			if (!usable(m, i, l)) continue;
			for (c = 0; c < m->ploid; c++) {
				int a = m->x[XO(m, l, i, c)];
				for (k = 0; k < m->K; k++) {
					if (init_flag == 1) w[k] = (double)(k + 1) / m->K;
					else {
						w[k] = q[k] * m->freq[FO(m, k, l, a)];
						if (k >= 1) w[k] += w[k - 1];
					}
				}
				m->z[XO(m, l, i, c)] = (int8_t)draw_bracket(&m->rng, w, m->K);
			}
		}
		for (k = 0; k < m->K; k++) cnt[k] = 0.0;
		for (l = 0; l < m->L; l++)
			if (usable(m, i, l))
				for (c = 0; c < m->ploid; c++) cnt[m->z[XO(m, l, i, c)]] += 1.0;
		for (k = 0; k < m->K; k++) w[k] = cnt[k];
		draw_dirichlet(&m->rng, w, m->K, q, m->alpha);
	}
}

/* update_alpha, mcmc.c:1244-1263 */
void orc_update_alpha(orc_model *m)
{
	double ralpha = draw_normal(&m->rng, m->alpha, 1.0);
	if (ralpha > 0) {
		double mh = orc_alpha_ratio_product(m, ralpha);
		m->alpha = (u01(&m->rng) < MIN2(1, mh)) ? ralpha : m->alpha;
	}
}

/* cal_lkh, mcmc.c:1916-1942 (modes 2 and 3) */
/* log_ld_noselfing_indv, mcmc.c:1869-1891 (mode 1: no selfing term) */
double orc_log_ld_noselfing(const orc_model *m, int i)
{
	double ll = 0;
	int l, c;
	for (l = 0; l < m->L; l++) {
		if (!usable(m, i, l)) continue;
		for (c = 0; c < m->ploid; c++) ll += log(m->freq[FO(m, m->z[XO(m, l, i, c)], l, m->x[XO(m, l, i, c)])]);
		if (m->x[XO(m, l, i, 0)] != m->x[XO(m, l, i, 1)]) ll += log(2);
	}
	return ll;
}

/* log_ld_indv_K, mcmc.c:1893-1913: individual i wholly in cluster k */
double orc_log_ld_indv_K(const orc_model *m, int i, int k)
{
	double ll = 0;
	int l, c;
	for (l = 0; l < m->L; l++) {
		if (!usable(m, i, l)) continue;
		for (c = 0; c < m->ploid; c++) ll += log(m->freq[FO(m, k, l, m->x[XO(m, l, i, c)])]);
		if (m->x[XO(m, l, i, 0)] != m->x[XO(m, l, i, 1)]) ll += log(2);
	}
	return ll;
}

/* update_Z, mcmc.c:1094-1119 (mode 0): whole-individual assignment, weights relative to cluster 0 */
void orc_update_Z(orc_model *m, int init_flag)
{
	double *tmp = m->scratch, temp = 0;
	int i, k;
	for (i = 0; i < m->N; i++) {
		for (k = 0; k < m->K; k++) {
			if (init_flag == 1) tmp[k] = (double)(k + 1) / m->K;
			else {
				tmp[k] = orc_log_ld_indv_K(m, i, k);
				if (k == 0) temp = tmp[k];
				tmp[k] = exp(tmp[k] - temp);
				if (k >= 1) tmp[k] += tmp[k - 1];
			}
		}
		m->zz[i] = draw_bracket(&m->rng, tmp, m->K);
	}
}

/* cal_lkh, mcmc.c:1916-1942 */
void orc_cal_lkh(orc_model *m)
{
	int i;
	m->totallkh = 0;
	for (i = 0; i < m->N; i++) {
		if (m->mode == 0) m->indvlkh[i] = orc_log_ld_indv_K(m, i, m->zz[i]);
		else if (m->mode == 4) m->indvlkh[i] = orc_log_ld_F(m, m->self_rates, 1, i);
		else if (m->mode == 5) m->indvlkh[i] = orc_log_ld_F(m, m->self_rates + i, 0, i);
		else m->indvlkh[i] = (m->mode == 1) ? orc_log_ld_noselfing(m, i) : orc_log_ld_indv(m, m->gen[i], i);
		m->totallkh += m->indvlkh[i];
	}
}

/* ------------------------------------------------------------------ DP prior (DPMM.c) - */


/* The reference keeps a value-sorted singly linked list of NODEs (DPMM.h:10-16), each with
 * a member count and a member index array.  Only the counts, the values and the list order
 * influence the sampler, so the restatement keeps a slot pool with index links. */
static void dp_reset(orc_model *m)
{
	int s;
	for (s = 0; s <= m->N; s++) m->dp_next[s] = s + 1;
	m->dp_next[m->N] = -1;
	m->dp_free = 0; m->dp_head = -1; m->dp_cnt = 0;
}
static int dp_alloc(orc_model *m)
{
	int s = m->dp_free;
	m->dp_free = m->dp_next[s];
	return s;
}

/* creat + find, DPMM.c:202-263: a new singleton cluster goes in front of the head when its
 * value is <= the head's, otherwise after the LAST slot whose value is <= the new value. */
static int dp_create(orc_model *m, double v)
{
	int s = dp_alloc(m), p, q;
	m->dp_value[s] = v; m->dp_num[s] = 1;
	if (m->dp_head < 0) { m->dp_next[s] = -1; m->dp_head = s; return s; }
	if (v <= m->dp_value[m->dp_head]) { m->dp_next[s] = m->dp_head; m->dp_head = s; return s; }
	p = q = m->dp_head;
	while (p >= 0 && m->dp_value[p] <= v) { q = p; p = m->dp_next[p]; }
	m->dp_next[s] = m->dp_next[q];
	m->dp_next[q] = s;
	return s;
}

/* delete, DPMM.c:280-321: drop individual j from its cluster; unlink a cluster that emptied */
static int dp_remove(orc_model *m, int j)
{
	int s = m->dp_of[j], p, q, removed = 0;
	m->dp_num[s]--;
	if (m->dp_num[s] > 0 || m->dp_head < 0) return 0;
	if (m->dp_head == s) { m->dp_head = m->dp_next[s]; removed = 1; }
	else {
		q = m->dp_head; p = m->dp_next[q];
		while (p >= 0 && p != s) { q = p; p = m->dp_next[p]; }
		if (p == s) { m->dp_next[q] = m->dp_next[s]; removed = 1; }
	}
	if (removed) { m->dp_next[s] = m->dp_free; m->dp_free = s; }
	return removed;
}

static int dp_nth(const orc_model *m, int n)   /* n-th slot in list order, 0-based */
{
	int p = m->dp_head;
	while (n-- > 0 && p >= 0) p = m->dp_next[p];
	return p;
}

/* init_DP, DPMM.c:124-161: a draw from the Chinese-restaurant prior, atoms ~ U(0,1) */
void orc_init_DP(orc_model *m)
{
	double *pr = m->scratch;      /* needs up to N+1 entries */
	int j, i, p, pick;
	double a = m->alpha_dpm;
	dp_reset(m);
	for (j = 0; j < m->N; j++) {
		pr[0] = a / (a + (double)j);
		for (i = 1, p = m->dp_head; i <= m->dp_cnt && p >= 0; i++, p = m->dp_next[p])
			pr[i] = pr[i - 1] + (double)m->dp_num[p] / (a + (double)j);
		pick = draw_bracket(&m->rng, pr, m->dp_cnt + 1);
		if (pick == 0) {
			m->dp_sval[j] = u01(&m->rng);
			m->dp_of[j] = dp_create(m, m->dp_sval[j]);
			m->dp_cnt++;
		} else {
			p = dp_nth(m, pick - 1);
			m->dp_sval[j] = m->dp_value[p];
			m->dp_num[p]++;
			m->dp_of[j] = p;
		}
	}
	for (j = 0; j < m->N; j++) m->self_rates[j] = m->dp_sval[j];     /* mcmc.c:322-323 */
}

/* update_DP + gen_post_prob + sample_poster (mode-3 branches), DPMM.c:165-199,361-377,392-398 */
void orc_update_DP(orc_model *m)
{
	double *pr = m->scratch;
	int j, i, p, pick;
	for (j = 0; j < m->N; j++) {
		int g = m->gen[j];
		m->dp_cnt -= dp_remove(m, j);
		pr[0] = m->alpha_dpm / (g + 1) / g;                 /* new cluster: alpha * B(g, 2) */
		for (i = 1, p = m->dp_head; i <= m->dp_cnt && p >= 0; i++, p = m->dp_next[p])
			pr[i] = pr[i - 1] + m->dp_num[p] * orc_dgeom(m->dp_value[p], g);
		pick = draw_bracket(&m->rng, pr, m->dp_cnt + 1);
		if (pick == 0) {
			m->dp_sval[j] = draw_beta(&m->rng, (double)g, 2);
			m->dp_of[j] = dp_create(m, m->dp_sval[j]);
			m->dp_cnt++;
		} else {
			p = dp_nth(m, pick - 1);
			m->dp_sval[j] = m->dp_value[p];
			m->dp_num[p]++;
			m->dp_of[j] = p;
		}
	}
	for (j = 0; j < m->N; j++) m->self_rates[j] = m->dp_sval[j];     /* mcmc.c:340-341 */
}
int orc_dp_nclusters(const orc_model *m) { return m->dp_cnt; }

/* TEST HOOK (no reference counterpart): rebuild the cluster list from self_rates[] -- individuals with the same value
 * share a cluster -- so that one orc_update_DP scan can start from an injected clustering.  Built with the same
 * dp_create as the sampler, so the list order is the value order the reference's find/insert keeps (DPMM.c:202-263). */
void orc_dp_from_values(orc_model *m)
{
	int j, p;
	dp_reset(m);
	for (j = 0; j < m->N; j++) {
		double v = m->self_rates[j];
		for (p = m->dp_head; p >= 0; p = m->dp_next[p]) if (m->dp_value[p] == v) break;
		if (p >= 0) { m->dp_num[p]++; m->dp_of[j] = p; }
		else { m->dp_of[j] = dp_create(m, v); m->dp_cnt++; }
		m->dp_sval[j] = v;
	}
}

/* ------------------------------------------------------------------ sweeps / chains --- */

/* one sweep in the reference's order: mcmc.c:152-155 (mode 1), :210-215 (mode 2), :336-348 (mode 3) */
static void one_sweep(orc_model *m)
{
	orc_update_P(m);
	if (m->mode == 0) {                               /* mcmc_POP_no_admixture, mcmc.c:111-113 */
		orc_update_Z(m, 0);
		orc_cal_lkh(m);
		return;
	}
	if (m->mode == 1) {
		orc_update_ZQ(m, 0);
		orc_update_alpha(m);
		orc_cal_lkh(m);
		return;
	}
	if (m->mode == 4 || m->mode == 5) {               /* mcmc.c:266-270, :424-437 (uniform prior) */
		if (m->mode == 4) orc_update_F_POP(m); else orc_update_F_IND(m);
		orc_update_ZQ(m, 0);
		orc_update_alpha(m);
		orc_cal_lkh(m);
		return;
	}
	if (m->mode == 2) orc_update_S_POP(m);
	if (m->mode == 3) {
		if (m->prior_flag == 1) orc_update_DP(m);
		if (m->prior_flag == 0) orc_update_S_IND(m);
	}
	orc_update_G(m);
	orc_update_ZQ(m, 0);
	orc_update_alpha(m);
	orc_cal_lkh(m);
}
void orc_sweeps(orc_model *m, int n) { while (n-- > 0) one_sweep(m); }

orc_chain *orc_chain_new(const orc_model *m, int ckrep)
{
	orc_chain *c = (orc_chain *)calloc(1, sizeof(orc_chain));
	int ns = (m->mode == 3 || m->mode == 5) ? m->N : m->K;
	c->indvlkh = (double *)calloc(m->N, sizeof(double));
	c->qq = (double *)calloc((size_t)m->N * m->K, sizeof(double));
	c->qq2 = (double *)calloc((size_t)m->N * m->K, sizeof(double));
	c->self_rates = (double *)calloc(ns, sizeof(double));
	c->self_rates2 = (double *)calloc(ns, sizeof(double));
	c->gen = (double *)calloc(m->N, sizeof(double));
	c->gen2 = (double *)calloc(m->N, sizeof(double));
	c->convg = (double *)calloc(ckrep > 0 ? ckrep : 1, sizeof(double));
	return c;
}
void orc_chain_free(orc_chain *c)
{
	if (!c) return;
	free(c->indvlkh); free(c->qq); free(c->qq2); free(c->self_rates); free(c->self_rates2);
	free(c->gen); free(c->gen2); free(c->convg); free(c);
}

/* initialize_chn, mcmc.c:644-738: every running moment starts at 1 with step = 0 */
static void chain_reset(const orc_model *m, orc_chain *c)
{
	int ns = (m->mode == 3 || m->mode == 5) ? m->N : m->K;
	long j;
	c->step = 0; c->totallkh = 1; c->totallkh2 = 1;
	for (j = 0; j < m->N; j++) { c->indvlkh[j] = 1; c->gen[j] = 1; c->gen2[j] = 1; }
	for (j = 0; j < (long)m->N * m->K; j++) { c->qq[j] = (m->mode == 0) ? 0 : 1; c->qq2[j] = 1; }   /* CHAIN.z starts at 0, mcmc.c:671 */
	for (j = 0; j < ns; j++) { c->self_rates[j] = 1; c->self_rates2[j] = 1; }
}

/* the running-moment update of store_chn, mcmc.c:1327-1332 and its copies: the mean after
 * step+1 samples, written as mean*((step + x/mean)/(1+step)), with a 0-mean fallback */
static inline double run_mean(double mean, double xv, long step)
{
	if (mean != 0) return mean * ((step + xv / mean) / (1 + step));
	return xv / (1 + step);
}

/* store_chn, mcmc.c:1320-1456 (modes 2/3 members) */
void orc_store_chn(const orc_model *m, orc_chain *c)
{
	int ns = (m->mode == 3 || m->mode == 5) ? m->N : m->K;
	long j;
	c->totallkh = run_mean(c->totallkh, m->totallkh, c->step);
	c->totallkh2 = run_mean(c->totallkh2, m->totallkh * m->totallkh, c->step);
	for (j = 0; j < m->N; j++) c->indvlkh[j] = run_mean(c->indvlkh[j], m->indvlkh[j], c->step);
	if (m->mode == 0) {                               /* :1356-1362: CHAIN.z[i][zz[i]] += 1, kept here in qq */
		for (j = 0; j < m->N; j++) c->qq[j * m->K + m->zz[j]] += 1;
		c->step++;
		return;
	}
	for (j = 0; j < (long)m->N * m->K; j++) {
		c->qq[j] = run_mean(c->qq[j], m->qq[j], c->step);
		c->qq2[j] = run_mean(c->qq2[j], m->qq[j] * m->qq[j], c->step);
	}
	if (m->mode == 1) { c->step++; return; }          /* mode 1 stores neither selfing rates nor generations */
	for (j = 0; j < ns; j++) {
		c->self_rates[j] = run_mean(c->self_rates[j], m->self_rates[j], c->step);
		c->self_rates2[j] = run_mean(c->self_rates2[j], m->self_rates[j] * m->self_rates[j], c->step);
	}
	if (m->mode == 4 || m->mode == 5) { c->step++; return; }   /* :1388-1409 store inbreed in the same way; no generations */
	for (j = 0; j < m->N; j++) {
		/* :1428-1433 -- the fallbacks there use integer division; gen >= 1 keeps the mean non-zero */
		if (c->gen[j] != 0) c->gen[j] = c->gen[j] * ((c->step + m->gen[j] / c->gen[j]) / (1 + c->step));
		else c->gen[j] = m->gen[j] / (1 + c->step);
		if (c->gen2[j] != 0) c->gen2[j] = c->gen2[j] * ((c->step + m->gen[j] * m->gen[j] / c->gen2[j]) / (1 + c->step));
		else c->gen2[j] = m->gen[j] * m->gen[j] / (1 + c->step);
	}
	c->step++;
}

/* mcmc_POP_selfing (mcmc.c:182-239) and mcmc_INDV_selfing (mcmc.c:297-383).
 * Returns flag_empty_cluster.  initd is float like INIT.initd (initial.h:11). */
int orc_run_chain(orc_model *m, long update, long burnin, int thinning, int ckrep,
                  int nstep_check_empty, const float *initd, orc_chain *out)
{
	long step, cnt_step = 0;
	int i;
	out->flag_empty_cluster = 0;
	out->steps = (int)((update - burnin) / thinning);            /* mcmc.c:485 */
	if (m->mode == 0) {
		/* mcmc_POP_no_admixture, mcmc.c:90-131: no alpha, no empty-cluster check */
		orc_update_Z(m, 1);
		for (step = 0; step < update; step++) {
			one_sweep(m);
			if (step == burnin - 1) chain_reset(m, out);
			if (step >= burnin && (step + 1 - burnin) % thinning == 0) {
				orc_store_chn(m, out);
				if (cnt_step < ckrep) out->convg[cnt_step] = m->totallkh;
				cnt_step++;
			}
		}
		return 0;
	}
	m->alpha = u01(&m->rng) * 10;                                /* mcmc.c:479 */
	if (m->mode == 1) {
		/* mcmc_POP_admixture, mcmc.c:135-180: nothing else to initialise */
	} else if (m->mode == 2) {
		for (i = 0; i < m->N; i++) {                             /* mcmc.c:196-199 */
			double p = u01(&m->rng);
			m->gen[i] = draw_geom(&m->rng, p);
			if (m->gen[i] > 50) m->gen[i] = 50;
		}
		for (i = 0; i < m->K; i++) {                             /* mcmc.c:200-205 */
			m->self_rates[i] = initd[i];
			if (m->back_refl == 0) m->state[i] = orc_dt_stat(m->self_rates[i]);
		}
	} else if (m->mode == 4) {
		for (i = 0; i < m->K; i++) {                             /* mcmc.c:256-260 */
			m->self_rates[i] = initd[i];
			if (m->back_refl == 0) m->state[i] = orc_dt_stat(m->self_rates[i]);
		}
	} else if (m->mode == 5) {
		for (i = 0; i < m->N; i++) m->self_rates[i] = u01(&m->rng);          /* mcmc.c:416-419, uniform prior only */
	} else {
		if (m->prior_flag == 1) orc_init_DP(m);                  /* mcmc.c:318-324 */
		else for (i = 0; i < m->N; i++) m->self_rates[i] = u01(&m->rng);   /* :326-327 */
		/* :329-331 -- the >50 cap sits outside the loop in the reference, so the initial
		 * generations are effectively uncapped; restated as such */
		for (i = 0; i < m->N; i++) m->gen[i] = draw_geom(&m->rng, 1 - m->self_rates[i]);
	}
	orc_update_ZQ(m, 1);
	for (step = 0; step < update; step++) {
		one_sweep(m);
		if (step == burnin - 1) chain_reset(m, out);
		if (step >= burnin && (step + 1 - burnin) % thinning == 0) {
			orc_store_chn(m, out);
			if (cnt_step < ckrep) out->convg[cnt_step] = m->totallkh;
			cnt_step++;
		}
		if (cnt_step == nstep_check_empty) {
			out->flag_empty_cluster = orc_check_empty_cluster(m);
			if (out->flag_empty_cluster == 1) break;
		}
	}
	return out->flag_empty_cluster;
}

/* GelmanRubin exactly as written, check_converg.c:100-153.  chain_converg calls it with
 * totrep = ckrep although the trace holds n_chain*ckrep values (check_converg.c:67), so it
 * compares n_chain consecutive segments of chain 0. */
double orc_gelman_rubin_ref(const double *vec, int numchains, int totrep)
{
	int per = totrep / numchains, i, j;
	double psi = 0, W = 0, B = 0, V;
	double *mu = (double *)calloc(numchains, sizeof(double));
	for (i = 0; i < numchains; i++) {
		for (j = 0; j < per; j++) mu[i] += vec[i * per + j];
		mu[i] = mu[i] / per;
		psi = psi + mu[i];
	}
	psi = psi / numchains;
	for (i = 0; i < numchains; i++) {
		double s = 0;
		for (j = 0; j < per; j++) s += (vec[i * per + j] - mu[i]) * (vec[i * per + j] - mu[i]);
		s = s / (per - 1);
		W += s;
	}
	W = W / numchains;
	for (i = 0; i < numchains; i++) B += (mu[i] - psi) * (mu[i] - psi);
	B = (B * per) / (numchains - 1);
	V = (W * (per - 1)) / per + B / per;
	free(mu);
	return V / W;
}
/* the statistic the comment block above GelmanRubin describes: m chains of n draws each */
double orc_gelman_rubin(const double *vec, int numchains, int n)
{
	return orc_gelman_rubin_ref(vec, numchains, numchains * n);
}
