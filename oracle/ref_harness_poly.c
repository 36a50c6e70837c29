/*
 * oracle/ref_harness_poly.c -- TEST INFRASTRUCTURE ONLY.
 * Third harness translation unit: includes the reference's poly_geno.c in place (resolved
 * through -I/root/reference; nothing is copied) to reach the static autotetraploid
 * functions (poly_geno.c:75-140 driver and the helpers listed in SURVEY.md section 8 a16).
 */
#include "poly_geno.c"

int refp_auto_geno_num(int A, int cat)
{
	/* number of genotypes per dosage class for A alleles, poly_geno.c:1708-1713 */
	int n[5]; n[0] = A; n[1] = A * (A - 1); n[2] = A * (A - 1) / 2; n[3] = A * (A - 1) * (A - 2) / 2;
	n[4] = A * (A - 1) * (A - 2) * (A - 3) / 24; return n[cat];
}

/* ---- autotetraploid harness (TEST INFRASTRUCTURE ONLY) -------------------------------
 * Builds the reference's own SEQDATA / POLY for ploid 4 from flat arrays and exposes
 *   - the genotype-frequency tables (calc_exfreq_auto poly_geno.c:1515, auto_genfreq :1803)
 *     on injected allele frequencies and selfing rates,
 *   - a whole chain through the reference's own driver mcmc_updating() -> mcmc_POP_tetra_selfing
 *     (poly_geno.c:75-140),
 * so that oracle/tetra_oracle.c can be pinned bit for bit. */
typedef struct {
	SEQDATA data;
	POLY *poly;
} REFP;

REFP *refp_new2(int N, int L, int K, int back_refl, int autopoly, const int *x, const int *alleleid, const int *allelenum);
REFP *refp_new(int N, int L, int K, int back_refl, const int *x /*[N][L][4]*/, const int *alleleid /*[N][L]*/,
               const int *allelenum /*[L]*/)
{
	return refp_new2(N, L, K, back_refl, 1, x, alleleid, allelenum);
}

/* autopoly = 0: the allotetraploid model (-ap 0) */
REFP *refp_new2(int N, int L, int K, int back_refl, int autopoly, const int *x /*[N][L][4]*/, const int *alleleid /*[N][L]*/,
                const int *allelenum /*[L]*/)
{
	int i, j, k, amax = 0;
	REFP *h = (REFP *)calloc(1, sizeof(REFP));
	SEQDATA *d = &h->data;
	d->ploid = 4; d->popnum = K; d->locinum = L; d->totalsize = N;
	d->mode = 2; d->prior_flag = 0; d->back_refl = back_refl; d->type_freq = 1; d->autopoly = autopoly;
	d->nstep_check_empty_cluster = 1 << 30; d->print_iter = 0; d->print_freq = 0;
	d->missingnum = -9; d->missingdata = "-9";
	d->seqdata = i3tensor(0, N - 1, 0, L - 1, 0, 3);
	d->alleleid = imatrix(0, N - 1, 0, L - 1);
	d->allelenum = ivector(0, L - 1);
	for (j = 0; j < L; j++) { d->allelenum[j] = allelenum[j]; if (allelenum[j] > amax) amax = allelenum[j]; }
	d->allelenum_max = amax;
	for (i = 0; i < N; i++)
		for (j = 0; j < L; j++) {
			d->alleleid[i][j] = alleleid[(long)i * L + j];
			for (k = 0; k < 4; k++) d->seqdata[i][j][k] = x[((long)i * L + j) * 4 + k];
		}
	/* get_missing_tetra (data_interface.c:722-741) lives in the other harness TU's include;
	 * its rule is one line, restated here: missing <=> no allele observed */
	d->missvec = ivector(0, N - 1);
	d->missindx = imatrix(0, N - 1, 0, L - 1);
	for (i = 0; i < N; i++) {
		d->missvec[i] = 0;
		for (j = 0; j < L; j++) { d->missindx[i][j] = (d->alleleid[i][j] == 0); d->missvec[i] += d->missindx[i][j]; }
	}
	h->poly = (POLY *)malloc(sizeof(POLY));
	gen_polyinfo(h->poly, *d);
	return h;
}

int refp_gmax(REFP *h)
{
	int j, g = 0;
	for (j = 0; j < h->poly->num_allele; j++) if (h->poly->genonum[j][0] > g) g = h->poly->genonum[j][0];
	return g;
}

/* genotype list of locus l: codes in table order; returns the count */
int refp_genolist(REFP *h, int l, int *codes)
{
	int id = find_id(h->data.allelenum[l], h->poly->allele_poly, h->poly->num_allele), g;
	for (g = 0; g < h->poly->genonum[id][0]; g++) codes[g] = h->poly->genolist[id][g];
	return h->poly->genonum[id][0];
}

/* tables on injected state: exfreq / genofreq as float [K][L][Gmax] */
void refp_tables(REFP *h, const double *freq /*[K][L][Amax]*/, const double *S /*[K]*/, float *exfreq, float *genofreq, int Gmax)
{
	SEQDATA *d = &h->data;
	UPMCMC up;
	int k, l, a, g, A = d->allelenum_max;
	up.freq = d3tensor(0, d->popnum - 1, 0, d->locinum - 1, 0, A - 1);
	for (k = 0; k < d->popnum; k++) for (l = 0; l < d->locinum; l++) for (a = 0; a < A; a++)
		up.freq[k][l][a] = freq[((long)k * d->locinum + l) * A + a];
	if (d->autopoly == 1) calc_exfreq_auto(&up, *d, h->poly);
	else {
		/* allotetraploid: the second subgenome's frequencies follow the first in the same array, [2][K][L][Amax] */
		up.freq2 = d3tensor(0, d->popnum - 1, 0, d->locinum - 1, 0, A - 1);
		for (k = 0; k < d->popnum; k++) for (l = 0; l < d->locinum; l++) for (a = 0; a < A; a++)
			up.freq2[k][l][a] = freq[(((long)d->popnum + k) * d->locinum + l) * A + a];
		calc_exfreq_allo(&up, *d, h->poly);
		free_d3tensor(up.freq2, 0, d->popnum - 1, 0, d->locinum - 1, 0, A - 1);
	}
	for (k = 0; k < d->popnum; k++) calc_self_genofreq(S[k], h->poly->genofreq[k], *d, h->poly, k);
	for (k = 0; k < d->popnum; k++) for (l = 0; l < d->locinum; l++) {
		int id = find_id(d->allelenum[l], h->poly->allele_poly, h->poly->num_allele);
		for (g = 0; g < h->poly->genonum[id][0]; g++) {
			exfreq[((long)k * d->locinum + l) * Gmax + g] = h->poly->exfreq[k][l][g];
			genofreq[((long)k * d->locinum + l) * Gmax + g] = h->poly->genofreq[k][l][g];
		}
	}
	free_d3tensor(up.freq, 0, d->popnum - 1, 0, d->locinum - 1, 0, A - 1);
}

/* a whole chain through the reference's own driver; outputs are the CHAIN running moments */
int refp_mcmc_updating(REFP *h, long update, long burnin, int thinning, int ckrep, const float *initd,
                       double *out_tot /*[2]*/, double *out_indvlkh, double *out_qq, double *out_qq2,
                       double *out_self, double *out_self2, double *out_convg)
{
	INIT init; CONVG cvg; CHAIN c; int i, k;
	SEQDATA *d = &h->data;
	memset(&init, 0, sizeof(init));
	init.chainnum = 1; init.update = update; init.burnin = burnin; init.thinning = thinning; init.popnum = d->popnum;
	init.initd = matrix(0, 0, 0, d->popnum - 1);
	for (k = 0; k < d->popnum; k++) init.initd[0][k] = initd ? initd[k] : 0.5f;
	init.name_len = ivector(0, 0); init.chn_name = cmatrix(0, 0, 0, 99);
	strcpy(init.chn_name[0], "Chain#1"); init.name_len[0] = 8;
	cvg.n_chain = 1; cvg.ckrep = ckrep; cvg.convgfilename = NULL;
	cvg.convg_ld = dvector(0, ckrep > 0 ? ckrep - 1 : 0);
	c = mcmc_updating(*d, init, 0, &cvg);
	if (c.flag_empty_cluster == 1) return 1;
	if (out_tot) { out_tot[0] = c.totallkh; out_tot[1] = c.totallkh2; }
	for (i = 0; i < d->totalsize; i++) {
		if (out_indvlkh) out_indvlkh[i] = c.indvlkh[i];
		for (k = 0; k < d->popnum; k++) {
			if (out_qq) out_qq[(long)i * d->popnum + k] = c.qq[i][k];
			if (out_qq2) out_qq2[(long)i * d->popnum + k] = c.qq2[i][k];
		}
	}
	for (k = 0; k < d->popnum; k++) {
		if (out_self) out_self[k] = c.self_rates[k];
		if (out_self2) out_self2[k] = c.self_rates2[k];
	}
	if (out_convg) for (i = 0; i < ckrep; i++) out_convg[i] = cvg.convg_ld[i];
	return 0;
}
