/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Reaches the `static` hot-path functions of the UNMODIFIED reference by including its
 * translation unit in place:  #include "mcmc.c"  is resolved by the -I flag in
 * oracle/Makefile to /root/reference/mcmc.c -- no reference source is copied into this
 * repository, and the built object lands in oracle/_ref/ (git-ignored).
 *
 * What it is for:
 *   - pin the in-repo C restatement (oracle/instruct_oracle.c) against the reference's own
 *     arithmetic on injected state: tallies (update_P, mcmc.c:810-845) bit-exact,
 *     log_ld_indv (mcmc.c:1726) / proposal (mcmc.c:1630) / cal_lkh (mcmc.c:1916) to 1e-12,
 *     and whole chains through the reference's own driver mcmc_updating() (mcmc.c:63);
 *   - time the reference's CPU sweep for bench.py's cpu_baseline / --impl reference.
 *
 * The tally inside update_P is not observable from outside (it lives in a local
 * i3tensor that is freed before return), so the Makefile compiles the reference's
 * random.c with -Drdirich=ref_real_rdirich and this file supplies `rdirich`, which
 * records the count vectors the reference passes in and then forwards to the real one.
 *
 * All entry points take flat, row-major C arrays so Python (ctypes) can drive them.
 */
/* mcmc_INDV_selfing reads and may write generation[N], one element past its ivector
 * (mcmc.c:329-331), which corrupts the heap of the test process.  The
 * harness gives every ivector allocated from this TU eight ints of slack; the layout that
 * free_ivector (nrutil.c) expects is unchanged.  Build-recipe fix, source untouched. */
#define ivector refh_slack_ivector
#include "mcmc.c"
#undef ivector
int *refh_slack_ivector(long nl, long nh)
{
	int *v = (int *)malloc((size_t)((nh - nl + 1 + 1 + 8) * sizeof(int)));
	if (!v) nrerror("allocation failure in refh_slack_ivector()");
	return v - nl + 1;
}

void ref_real_rdirich(double *alpha, int length, double **rand, double add);

/* ---- rdirich capture -------------------------------------------------------------- */
static double *cap_buf = NULL;   /* concatenated alpha vectors, in call order */
static long cap_len = 0, cap_max = 0;
static int cap_on = 0;

void rdirich(double *alpha, int length, double **rand, double add)
{
	int k;
	if (cap_on) {
		if (cap_len + length > cap_max) {
			cap_max = 2 * (cap_len + length) + 1024;
			cap_buf = (double *)realloc(cap_buf, cap_max * sizeof(double));
		}
		for (k = 0; k < length; k++) cap_buf[cap_len++] = alpha[k];
	}
	ref_real_rdirich(alpha, length, rand, add);
}

/* ---- harness object --------------------------------------------------------------- */
typedef struct {
	SEQDATA data;
	UPMCMC *ptr;
	double **qqnum;
	NODE *head;
	SF *indv_array;
	int cnt_node;
} REFH;

/* provided by ref_harness_data.c (which includes the reference's data_interface.c) */
void refd_get_missing(SEQDATA *data);

REFH *refh_new(int N, int L, int K, int ploid, int mode, int prior_flag, int back_refl,
               int type_freq, double alpha_dpm, const int *x /*[N][L][ploid]*/,
               const int *allelenum /*[L]*/)
{
	int i, j, k, amax = 0;
	REFH *h = (REFH *)calloc(1, sizeof(REFH));
	SEQDATA *d = &h->data;
	memset(d, 0, sizeof(SEQDATA));
	d->ploid = ploid; d->popnum = K; d->locinum = L; d->totalsize = N;
	d->mode = mode; d->prior_flag = prior_flag; d->back_refl = back_refl;
	d->type_freq = type_freq; d->alpha_dpm = alpha_dpm;
	d->nstep_check_empty_cluster = 1 << 30; d->print_iter = 0; d->print_freq = 0;
	d->missingnum = -9; d->missingdata = "-9"; d->autopoly = 1;
	d->seqdata = i3tensor(0, N - 1, 0, L - 1, 0, ploid - 1);
	d->allelenum = refh_slack_ivector(0, L - 1);
	for (j = 0; j < L; j++) { d->allelenum[j] = allelenum[j]; if (allelenum[j] > amax) amax = allelenum[j]; }
	d->allelenum_max = amax;
	for (i = 0; i < N; i++)
		for (j = 0; j < L; j++)
			for (k = 0; k < ploid; k++)
				d->seqdata[i][j][k] = x[((long)i * L + j) * ploid + k];
	refd_get_missing(d);      /* the reference's own get_missing(), data_interface.c:812 */
	allocate_node(&h->ptr, *d);
	h->qqnum = dmatrix(0, N - 1, 0, K - 1);
	h->ptr->alpha = 1.0;
	for (i = 0; i < N; i++) for (k = 0; k < K; k++) { h->qqnum[i][k] = 0; if (mode != 0) h->ptr->qq[i][k] = 1.0 / K; }
	if (mode == 0) for (i = 0; i < N; i++) h->ptr->zz[i] = 0;
	for (k = 0; k < K; k++) for (j = 0; j < L; j++) for (i = 0; i < amax; i++) h->ptr->freq[k][j][i] = 0.0;
	return h;
}

void refh_set_flags(REFH *h, int nstep_check, int print_iter, int print_freq)
{
	h->data.nstep_check_empty_cluster = nstep_check;
	h->data.print_iter = print_iter;
	h->data.print_freq = print_freq;
}

int refh_amax(REFH *h) { return h->data.allelenum_max; }

void refh_get_missindx(REFH *h, int *out /*[N][L]*/)
{
	int i, j;
	for (i = 0; i < h->data.totalsize; i++)
		for (j = 0; j < h->data.locinum; j++)
			out[(long)i * h->data.locinum + j] = h->data.missindx[i][j];
}

/* state accessors: dir 0 = get (reference -> flat), 1 = set (flat -> reference) */
void refh_z(REFH *h, int *z /*[N][L][ploid]*/, int dir)
{
	int i, j, k; SEQDATA *d = &h->data;
	for (i = 0; i < d->totalsize; i++) for (j = 0; j < d->locinum; j++) for (k = 0; k < d->ploid; k++) {
		long o = ((long)i * d->locinum + j) * d->ploid + k;
		if (dir) h->ptr->z[i][j][k] = z[o]; else z[o] = h->ptr->z[i][j][k];
	}
}
void refh_zz(REFH *h, int *zz /*[N]*/, int dir)
{
	int i; for (i = 0; i < h->data.totalsize; i++) { if (dir) h->ptr->zz[i] = zz[i]; else zz[i] = h->ptr->zz[i]; }
}
void refh_update_Z(REFH *h, int init_flag) { update_Z(&h->ptr, h->data, init_flag); }
double refh_log_ld_indv_K(REFH *h, int i, int k) { return log_ld_indv_K(h->ptr, h->data, i, k); }
void refh_qq(REFH *h, double *qq /*[N][K]*/, int dir)
{
	int i, k; SEQDATA *d = &h->data;
	for (i = 0; i < d->totalsize; i++) for (k = 0; k < d->popnum; k++) {
		if (dir) h->ptr->qq[i][k] = qq[(long)i * d->popnum + k]; else qq[(long)i * d->popnum + k] = h->ptr->qq[i][k];
	}
}
void refh_qqnum(REFH *h, double *q /*[N][K]*/, int dir)
{
	int i, k; SEQDATA *d = &h->data;
	for (i = 0; i < d->totalsize; i++) for (k = 0; k < d->popnum; k++) {
		if (dir) h->qqnum[i][k] = q[(long)i * d->popnum + k]; else q[(long)i * d->popnum + k] = h->qqnum[i][k];
	}
}
void refh_freq(REFH *h, double *f /*[K][L][Amax]*/, int dir)
{
	int i, j, k; SEQDATA *d = &h->data; int A = d->allelenum_max;
	for (k = 0; k < d->popnum; k++) for (j = 0; j < d->locinum; j++) for (i = 0; i < A; i++) {
		long o = ((long)k * d->locinum + j) * A + i;
		if (dir) h->ptr->freq[k][j][i] = f[o]; else f[o] = h->ptr->freq[k][j][i];
	}
}
void refh_gen(REFH *h, int *g, int dir)
{
	int i; for (i = 0; i < h->data.totalsize; i++) { if (dir) h->ptr->generation[i] = g[i]; else g[i] = h->ptr->generation[i]; }
}
static int n_self(REFH *h) { return (h->data.mode == 3 || h->data.mode == 5) ? h->data.totalsize : h->data.popnum; }
/* modes 4/5 keep their inbreeding coefficients in UPMCMC.inbreed (mcmc.c:521-522) */
static double *self_vec(REFH *h) { return (h->data.mode == 4 || h->data.mode == 5) ? h->ptr->inbreed : h->ptr->self_rates; }
void refh_self(REFH *h, double *s, int dir)
{
	int i; double *v = self_vec(h);
	for (i = 0; i < n_self(h); i++) { if (dir) v[i] = s[i]; else s[i] = v[i]; }
}
void refh_state(REFH *h, int *s, int dir)
{
	int i; for (i = 0; i < h->data.popnum; i++) { if (dir) h->ptr->state[i] = s[i]; else s[i] = h->ptr->state[i]; }
}
void refh_alpha(REFH *h, double *a, int dir) { if (dir) h->ptr->alpha = *a; else *a = h->ptr->alpha; }
void refh_lkh(REFH *h, double *indv /*[N]*/, double *total)
{
	int i; for (i = 0; i < h->data.totalsize; i++) indv[i] = h->ptr->indvlkh[i];
	*total = h->ptr->totallkh;
}

/* ---- RNG ---------------------------------------------------------------------------- */
void refh_setseeds(int a, int b, int c) { setseeds(a, b, c); }
double refh_ran1(void) { return ran1(); }

/* ---- single conditional updates on the injected state -------------------------------- */

/* update_P (mcmc.c:799): returns the tally n[k][l][a] the reference computed (captured from
 * the rdirich arguments, which are exactly (double)seqpop[i][j][k], mcmc.c:852-854). */
void refh_update_P(REFH *h, int *tally /*[K][L][Amax], may be NULL*/)
{
	int i, j, k; SEQDATA *d = &h->data; long p = 0;
	cap_on = (tally != NULL); cap_len = 0;
	update_P(&h->ptr, *d);
	cap_on = 0;
	if (!tally) return;
	for (i = 0; i < d->popnum; i++) for (j = 0; j < d->locinum; j++) for (k = 0; k < d->allelenum_max; k++)
		tally[((long)i * d->locinum + j) * d->allelenum_max + k] = 0;
	for (i = 0; i < d->popnum; i++)
		for (j = 0; j < d->locinum; j++)
			if (d->allelenum[j] > 1)
				for (k = 0; k < d->allelenum[j]; k++)
					tally[((long)i * d->locinum + j) * d->allelenum_max + k] = (int)cap_buf[p++];
}
void refh_update_ZQ(REFH *h, int init_flag) { update_ZQ(&h->ptr, h->data, init_flag, &h->qqnum); }
void refh_update_G(REFH *h) { update_G(h->data, &h->ptr); }
void refh_update_S_POP(REFH *h) { update_S_POP(h->data, &h->ptr); }
void refh_update_S_IND(REFH *h) { update_S_IND(h->data.totalsize, &h->ptr); }
void refh_update_F_POP(REFH *h) { update_inbreedcoff_POP(h->data, &h->ptr); }
void refh_update_F_IND(REFH *h) { update_F_IND(h->data.totalsize, &h->ptr, h->data); }
double refh_log_ld_F(REFH *h, double *inbreed, int by_pop, int i)
{
	return by_pop ? log_ld_F_pop(inbreed, h->ptr, i, h->data) : log_ld_F_indv(inbreed[0], h->ptr, i, h->data);
}
double refh_log_ld_F_total(REFH *h, double *inbreed) { return log_ld_F_total(inbreed, h->ptr, h->data); }
void refh_update_alpha(REFH *h) { update_alpha(&h->ptr, h->data, h->qqnum); }
void refh_cal_lkh(REFH *h) { cal_lkh(&h->ptr, h->data); }
double refh_log_ld_indv(REFH *h, int gen, int i) { return log_ld_indv(gen, h->ptr, i, h->data); }
double refh_proposal(REFH *h, double *S) { return proposal(S, h->ptr->generation, h->ptr->qq, h->data.totalsize, h->data.popnum); }
double refh_dgeom(double s, int g) { return dgeom(s, g); }
double refh_genofreq(int a0, int a1, double f0, double f1, int gen)
{
	int ld[2]; double fr[2]; ld[0] = a0; ld[1] = a1; fr[0] = f0; fr[1] = f1;
	return genofreq(ld, fr, gen, 2);
}
int refh_dt_stat(double s) { return dt_stat(s); }
int refh_check_empty_cluster(REFH *h) { return check_empty_cluster(h->ptr, h->data); }
int refh_rgeom(double p) { return rgeom(p); }
int refh_disc_unif(double *vec, int len) { return disc_unif(vec, len); }
double refh_rgamma(double a, double b) { return rgamma(a, b); }
double refh_rbeta(double a, double b) { return rbeta(a, b); }
double refh_rnormal(double m, double s) { return rnormal(m, s); }

/* ---- DP prior (DPMM.c:124,165); DPMM.c is compiled with -Divector=refh_plain_ivector so
 *      that the realloc() in insert() (DPMM.c:271) acts on a pointer malloc() returned
 *       -- a build-recipe fix, the source is untouched. ------------- */
int *refh_plain_ivector(long nl, long nh) { return (int *)malloc((size_t)(nh - nl + 1) * sizeof(int)) - nl; }

void refh_init_DP(REFH *h)
{
	int i;
	if (!h->indv_array) h->indv_array = (SF *)malloc(h->data.totalsize * sizeof(SF));
	h->head = NULL; h->cnt_node = 0;
	init_DP(&h->head, h->data.alpha_dpm, &h->indv_array, &h->cnt_node, h->data.totalsize);
	for (i = 0; i < h->data.totalsize; i++) h->ptr->self_rates[i] = h->indv_array[i].value;
}
void refh_update_DP(REFH *h)
{
	int j;
	update_DP(&h->head, h->data.alpha_dpm, &h->indv_array, &h->cnt_node, h->data.totalsize, h->data, h->ptr);
	for (j = 0; j < h->data.totalsize; j++) h->ptr->self_rates[j] = h->indv_array[j].value;
}
int refh_dp_nclusters(REFH *h) { return h->cnt_node; }

/* ---- n sweeps in the reference's own order (mcmc.c:208-215 mode 2, :334-348 mode 3) --- */
void refh_sweeps(REFH *h, int n)
{
	int s;
	for (s = 0; s < n; s++) {
		update_P(&h->ptr, h->data);
		if (h->data.mode == 2) update_S_POP(h->data, &h->ptr);
		if (h->data.mode == 3) {
			if (h->data.prior_flag == 1) refh_update_DP(h);
			if (h->data.prior_flag == 0) update_S_IND(h->data.totalsize, &h->ptr);
		}
		if (h->data.mode == 0) { update_Z(&h->ptr, h->data, 0); cal_lkh(&h->ptr, h->data); continue; }
		if (h->data.mode == 4) update_inbreedcoff_POP(h->data, &h->ptr);
		if (h->data.mode == 5) update_F_IND(h->data.totalsize, &h->ptr, h->data);
		if (h->data.mode == 2 || h->data.mode == 3) update_G(h->data, &h->ptr);
		update_ZQ(&h->ptr, h->data, 0, &h->qqnum);
		update_alpha(&h->ptr, h->data, h->qqnum);
		cal_lkh(&h->ptr, h->data);
	}
}

/* ---- a whole chain through the reference's own driver mcmc_updating() (mcmc.c:63) ------
 * out_* receive the CHAIN running moments (mcmc.h:29-53); NULL pointers are skipped. */
int refh_mcmc_updating(REFH *h, long update, long burnin, int thinning, int ckrep,
                       const float *initd /*[K]*/, double *out_tot /*[2]*/, double *out_indvlkh,
                       double *out_qq, double *out_qq2, double *out_self, double *out_self2,
                       double *out_gen, double *out_gen2, double *out_convg /*[ckrep]*/)
{
	INIT init; CONVG cvg; CHAIN c; int i, k, ns, flag;
	SEQDATA *d = &h->data;
	memset(&init, 0, sizeof(init));
	init.chainnum = 1; init.update = update; init.burnin = burnin; init.thinning = thinning; init.popnum = d->popnum;
	init.initd = matrix(0, 0, 0, d->popnum - 1);
	for (k = 0; k < d->popnum; k++) init.initd[0][k] = initd ? initd[k] : 0.5f;
	init.name_len = refh_slack_ivector(0, 0); init.chn_name = cmatrix(0, 0, 0, 99);
	strcpy(init.chn_name[0], "Chain#1"); init.name_len[0] = 8;
	cvg.n_chain = 1; cvg.ckrep = ckrep; cvg.convgfilename = NULL;
	cvg.convg_ld = dvector(0, ckrep > 0 ? ckrep - 1 : 0);
	c = mcmc_updating(*d, init, 0, &cvg);
	flag = (d->mode == 0) ? 0 : c.flag_empty_cluster;     /* mcmc_POP_no_admixture never sets the flag */
	if (flag == 1) return 1;
	ns = (d->mode == 3 || d->mode == 5) ? d->totalsize : d->popnum;
	if (out_tot) { out_tot[0] = c.totallkh; out_tot[1] = c.totallkh2; }
	for (i = 0; i < d->totalsize; i++) {
		if (out_indvlkh) out_indvlkh[i] = c.indvlkh[i];
		for (k = 0; k < d->popnum; k++) {
			if (d->mode == 0) {              /* CHAIN.z: retained samples of individual i in cluster k (mcmc.c:1356-1362) */
				if (out_qq) out_qq[(long)i * d->popnum + k] = (double)c.z[i][k];
				if (out_qq2) out_qq2[(long)i * d->popnum + k] = 1.0;
				continue;
			}
			if (out_qq) out_qq[(long)i * d->popnum + k] = c.qq[i][k];
			if (out_qq2) out_qq2[(long)i * d->popnum + k] = c.qq2[i][k];
		}
		if (d->mode == 2 || d->mode == 3) {
			if (out_gen) out_gen[i] = c.gen[i];
			if (out_gen2) out_gen2[i] = c.gen2[i];
		}
	}
	if (d->mode == 2 || d->mode == 3)
		for (i = 0; i < ns; i++) {
			if (out_self) out_self[i] = c.self_rates[i];
			if (out_self2) out_self2[i] = c.self_rates2[i];
		}
	if (d->mode == 4 || d->mode == 5)
		for (i = 0; i < ns; i++) {
			if (out_self) out_self[i] = c.inbreed[i];
			if (out_self2) out_self2[i] = c.inbreed2[i];
		}
	if (out_convg) for (i = 0; i < ckrep; i++) out_convg[i] = cvg.convg_ld[i];
	free_chain(&c, *d);
	return 0;
}

/* ---- the reference's result writer on given CHAIN moments (result_analysis.c:34), for the
 *      byte-parity test of instruct_b200/host/writer.c.  Labels are "ind<i>", pre-defined
 *      populations "pop<p>". ------------------------------------------------------------- */
double refh_chain_stat(REFH *h, const char *outfile, int label, int popdata, int distr_fmt, const int *popindx,
                       int pop_count, const int *missvec, const char *chn_name, double tot, double tot2,
                       const double *indvlkh, const double *qq, const double *qq2, const double *self,
                       const double *self2, const double *gen, const double *gen2)
{
	SEQDATA d = h->data;
	CHAIN c;
	int i, k, ns = (d.mode == 3 || d.mode == 5) ? d.totalsize : d.popnum;
	double dic;
	d.label = label; d.popdata = popdata; d.distr_fmt = distr_fmt; d.print_freq = 0; d.pop_count = pop_count;
	d.indvname = cmatrix(0, d.totalsize, 0, 99);      /* one spare row: print_S_INDV reads indvname[N] (App. B #2) */
	for (i = 0; i <= d.totalsize; i++) sprintf(d.indvname[i], "ind%d", i);
	d.poptype = (char **)malloc((pop_count + 1) * sizeof(char *));
	for (i = 0; i < pop_count; i++) { d.poptype[i] = (char *)malloc(32); sprintf(d.poptype[i], "pop%d", i); }
	d.popindx = refh_slack_ivector(0, d.totalsize - 1);
	d.missvec = refh_slack_ivector(0, d.totalsize - 1);
	for (i = 0; i < d.totalsize; i++) { d.popindx[i] = popindx[i]; d.missvec[i] = missvec[i]; }
	memset(&c, 0, sizeof(c));
	c.name_len = (int)strlen(chn_name) + 1;
	c.chn_name = cvector(0, c.name_len - 1);
	memcpy(c.chn_name, chn_name, c.name_len);
	c.steps = 10; c.totallkh = tot; c.totallkh2 = tot2;
	c.indvlkh = dvector(0, d.totalsize - 1);
	c.qq = dmatrix(0, d.totalsize - 1, 0, d.popnum - 1); c.qq2 = dmatrix(0, d.totalsize - 1, 0, d.popnum - 1);
	c.self_rates = dvector(0, ns); c.self_rates2 = dvector(0, ns);
	c.gen = dvector(0, d.totalsize - 1); c.gen2 = dvector(0, d.totalsize - 1);
	for (i = 0; i < d.totalsize; i++) {
		c.indvlkh[i] = indvlkh[i]; c.gen[i] = gen[i]; c.gen2[i] = gen2[i];
		for (k = 0; k < d.popnum; k++) { c.qq[i][k] = qq[(long)i * d.popnum + k]; c.qq2[i][k] = qq2[(long)i * d.popnum + k]; }
	}
	for (i = 0; i < ns; i++) { c.self_rates[i] = self[i]; c.self_rates2[i] = self2[i]; }
	c.self_rates[ns] = 0; c.self_rates2[ns] = 0;
	if (d.mode == 0) {                                          /* mode 0 prints CHAIN.z / steps (result_analysis.c:153-192) */
		c.z = lmatrix(0, d.totalsize - 1, 0, d.popnum - 1);
		for (i = 0; i < d.totalsize; i++) for (k = 0; k < d.popnum; k++) c.z[i][k] = (long)(qq[(long)i * d.popnum + k] * c.steps + 0.5);
	}
	c.inbreed = c.self_rates; c.inbreed2 = c.self_rates2;      /* modes 4/5 print CHAIN.inbreed (result_analysis.c:114-148) */
	dic = chain_stat((char *)outfile, c, d, 0);
	return dic;
}
