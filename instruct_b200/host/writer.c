/*
 * writer.c -- the `-o` result file, byte-compatible with the reference.
 *
 * Restates the OUTPUT FORMAT of printinfo() (InStruct.c:450-531), chain_stat() and its
 * print_* helpers (result_analysis.c:34-412) and chain_converg() (check_converg.c:44-91).
 * The format strings are the contract (north_star: "keeps ... the output file format"); the
 * code around them is this repository's own and works on the flat arrays of
 * ig_chain_result / gs_store instead of SEQDATA / CHAIN.
 *
 * Documented differences from the reference's bytes:
 *   - mode 3 selfing-rate table lists individuals 0..N-1; the reference loops 1..N, drops
 *     individual 0 and reads one element past the arrays;
 *   - the Gelman-Rubin line reports the across-chain statistic; the reference's as-written
 *     value (segments of chain 0 only, App. B #3) is available with --ref-compat-gr.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "writer.h"

/* ascending order of v[0..n), as indexx() (quantile.c:20) delivers it: out[r] = index */
static void order_ascending(int n, const double *v, int *out)
{
	int i, j;
	for (i = 0; i < n; i++) {
		for (j = i; j > 0 && v[out[j - 1]] > v[i]; j--) out[j] = out[j - 1];
		out[j] = i;
	}
}

int wr_banner(const char *path, int argc, char **argv, const wr_run *r)
{
	int i;
	FILE *f = fopen(path, "w");
	if (!f) return 1;
	fprintf(f, "\n");
	for (i = 0; i < 100; i++) fprintf(f, "=");
	fprintf(f, "\n\tInStruct by Gao, Williamson and Bustamante (2007)\n");
	fprintf(f, "\t\t  Code by Hong Gao\n");
	fprintf(f, "\t\tVersion 1.0 (May. 2007)\n");
	for (i = 0; i < 100; i++) fprintf(f, "=");
	fprintf(f, "\n\n\n\nCommand line arguments:\n    ");
	for (i = 0; i < argc; i++) fprintf(f, "%s ", argv[i]);
	fprintf(f, "\n\n");
	fprintf(f, "Data File:   %s\n", r->datafilename);
	if (r->initialfilename) fprintf(f, "Initial File:   %s\n", r->initialfilename);
	fprintf(f, "Output File:   %s\n\n", path);
	fprintf(f, "\nRun parameters:\n");
	fprintf(f, "    Chain Number=%d\n", r->chainnum);
	fprintf(f, "    MCMC Iterations Number=%ld\n", r->update);
	fprintf(f, "    Burn-in=%ld\n", r->burnin);
	fprintf(f, "    Thinning=%d\n", r->thinning);
	fprintf(f, "    Ploid=%d\n", r->ploid);
	if (r->ploid > 2) {
		if (r->autopoly == 1) fprintf(f, "Autopolyploid assumed\n");
		else if (r->autopoly == 0) fprintf(f, "Allopolyploid assumed\n");
	}
	fprintf(f, "    Missing Data=%s\n", r->missingdata);
	fprintf(f, "    Population size=%d\n", r->totalsize);
	fprintf(f, "    Number of loci=%d\n", r->locinum);
	fprintf(f, "    Population number assumed=%d\n", r->popnum);
	fprintf(f, "    Significance level for Posterior Credible Interval=%f\n", r->siglevel);
	fprintf(f, "    Mode = ");
	if (r->ploid == 2) {
		static const char *what[6] = {
		    "Make inference of population structure only without admixture.\n",
		    "Make inference of population structure only with admixture.\n",
		    "Make inference of population structure and the selfing rates for subpopulations.\n",
		    "Make inference of population structure and the selfing rates for individuals.\n",
		    "Make inference of population structure and the inbreeding coefficients for subpopulations.\n",
		    "Make inference of population structure and the inbreeding coefficients for individuals.\n"};
		if (r->mode >= 0 && r->mode <= 5) fprintf(f, "%s", what[r->mode]);
	} else if (r->ploid == 4)
		fprintf(f, "Make inference of population structure and the selfing rates for subpopulations.\n");
	if (r->inf_K == 1) fprintf(f, "\nMake inference of the number of subpopulations.\n");
	if (r->mode == 3 || r->mode == 5) {
		if (r->prior_flag == 0) fprintf(f, "The Uniform prior is used for selfing rates.\n\n");
		if (r->prior_flag == 1) fprintf(f, "The Dirichlet Process prior is used for selfing rates and the scaling parameter is %f.\n\n", r->alpha_dpm);
	}
	if (r->back_refl == 0) fprintf(f, "The proposal method for selfing rates is adaptive independence sampler.\n");
	if (r->back_refl == 1) fprintf(f, "The proposal method for selfing rates is back-reflection.\n");
	if (r->print_freq == 1) fprintf(f, "The posterior allele frequencies will also be summarized and written to output file.\n");
	if (r->GR_flag == 1) fprintf(f, "The %d stored iteration results after burn-in will be used to calculate the GR statistic.\n", r->ckrep);
	if (r->distr_fmt == 1) fprintf(f, "The output of Q are generated in the Distruct format.\n");
	fclose(f);
	return 0;
}

/* chain_stat, result_analysis.c:34-72.  Returns the DIC through *dic. */
int wr_chain(const char *path, const wr_run *r, const wr_data *d, const wr_chain_t *c, double *dic_out)
{
	int i, j, k, N = r->totalsize, K = r->popnum;
	int *ord = (int *)malloc((size_t)(K + 1) * sizeof(int));
	double dev = 0, dic;
	FILE *f = fopen(path, "a+");
	if (!f) { free(ord); return 1; }
	for (k = 0; k < K; k++) ord[k] = k;

	/* ---- print_lkh_to_file :389-412.  name_len counts the terminating NUL (initial.c:65), and
	 *      the reference writes it to the file (App. B #6); reproduced. */
	fprintf(f, "\n\n\n");
	fwrite(c->chn_name, 1, (size_t)c->name_len, f);
	fprintf(f, ":\n");
	fprintf(f, "\nThe log Likelihood:\n");
	fprintf(f, "    Posterior Mean = %.3f\n", c->totallkh);
	fprintf(stdout, "    Posterior Mean = %.3f\n", c->totallkh);
	fprintf(f, "    Posterior Variance = %.3f\n", c->totallkh2 - c->totallkh * c->totallkh);
	for (j = 0; j < N; j++) dev += c->indvlkh[j];
	dic = -4 * c->totallkh + 2 * dev;
	fprintf(f, "\nThe Deviance information criterion of this model is %f.\n", dic);
	if (dic_out) *dic_out = dic;

	/* ---- selfing rates */
	if ((r->ploid == 2 && r->mode == 2) || r->ploid == 4) {          /* print_S_POP_to_file :74-95 */
		order_ascending(K, c->self_rates, ord);
		fprintf(f, "\nThe Posterior distribution of Selfing Rates:\n");
		fprintf(f, "\t\tMean\tVar\n");
		for (j = 0; j < K; j++) {
			double m = c->self_rates[ord[j]];
			fprintf(f, "Cluster %d\t%.3f\t%.3f\n", j + 1, m, c->self_rates2[ord[j]] - m * m);
		}
	}
	if (r->ploid == 2 && r->mode == 4) {                            /* print_F_POP_to_file :114-133 */
		order_ascending(K, c->self_rates, ord);
		fprintf(f, "\nThe Posterior distribution of Inbreeding Coefficients:\n");
		fprintf(f, "\t\tMean\tVar\n");
		for (j = 0; j < K; j++) {
			double m = c->self_rates[ord[j]];
			fprintf(f, "Cluster %d\t%.3f\t%.3f\n", j + 1, m, c->self_rates2[ord[j]] - m * m);
		}
	}
	if (r->ploid == 2 && r->mode == 5) {                            /* print_F_INDV_to_file :135-148 */
		fprintf(f, "\nThe Posterior distribution of Inbreeding Coefficients:\n");
		fprintf(f, "\t\tMean\tVar\n");
		for (j = 0; j < N; j++) {
			fprintf(f, "Indv %d\t\t", j);
			if (r->label == 1) fprintf(f, "%s\t", d->indvname[j]);
			fprintf(f, "%.3f\t%.3f\n", c->self_rates[j], c->self_rates2[j] - c->self_rates[j] * c->self_rates[j]);
		}
	}
	if (r->ploid == 2 && r->mode == 3) {                            /* print_S_INDV_to_file :97-111 */
		fprintf(f, "\nThe Posterior distribution of Selfing Rates:\n");
		if (r->label == 1) fprintf(f, "\t");
		fprintf(f, "\t\tMean\tVar\n");
		for (j = 0; j < N; j++) {
			fprintf(f, "Indv %d\t\t", j);
			if (r->label == 1) fprintf(f, "%s\t", d->indvname[j]);
			fprintf(f, "%.3f\t%.3f\n", c->self_rates[j], c->self_rates2[j] - c->self_rates[j] * c->self_rates[j]);
		}
	}
	if (r->ploid == 2 && (r->mode == 2 || r->mode == 3)) {          /* print_gen_to_file :373-387 */
		fprintf(f, "\nThe Posterior distribution of Generations:\n");
		fprintf(f, "\t\tMean\tVariance\n");
		for (j = 0; j < N; j++) {
			fprintf(f, "Indv %d\t\t", j);
			if (r->label == 1) fprintf(f, "%s\t", d->indvname[j]);
			fprintf(f, "%.3f\t%.3f\n", c->gen[j], c->gen2[j] - c->gen[j] * c->gen[j]);
		}
	}

	/* ---- print_Z_to_file :153-192 (mode 0): c->qq holds CHAIN.z / steps, the share of the retained
	 *      samples in which the individual sat in each cluster */
	if (r->ploid == 2 && r->mode == 0) {
		fprintf(f, "\nInferred Classification of individuals:\n");
		fprintf(f, "\nIndv\t");
		if (r->label == 1) fprintf(f, "Label\t");
		fprintf(f, "(Miss)\t");
		if (r->popdata == 1) fprintf(f, "Pop : ");
		for (j = 0; j < K; j++) fprintf(f, "Prob in Cluster %d\t", j + 1);
		fprintf(f, "\n");
		for (j = 0; j < N; j++) {
			fprintf(f, "%d\t", j + 1);
			if (r->label == 1) fprintf(f, "%s\t", d->indvname[j]);
			fprintf(f, "(%d)\t", d->missvec[j]);
			if (r->popdata == 1) fprintf(f, "%d : ", d->popindx[j]);
			for (i = 0; i < K; i++) fprintf(f, "\t%f", c->qq[(size_t)j * K + i]);
			fprintf(f, "\n");
		}
	} else
	/* ---- print_Q_to_file :194-311 */
	{
		int pc = (r->popdata == 0) ? 1 : d->pop_count;
		double *acc = (double *)calloc((size_t)pc * K, sizeof(double));
		int *cnt = (int *)calloc((size_t)pc, sizeof(int));
		fprintf(f, "\nInferred ancestry of individuals:\n");
		fprintf(f, "\nIndv\t");
		if (r->label == 1) fprintf(f, "Label\t");
		fprintf(f, "(Miss)\tPop : ");
		if (r->distr_fmt == 0) for (j = 0; j < K; j++) fprintf(f, "Cluster %d:Mean\tVar\t\t", j + 1);
		else if (r->distr_fmt == 1) for (j = 0; j < K; j++) fprintf(f, "\tCluster %d", j + 1);
		fprintf(f, "\n");
		for (j = 0; j < N; j++) {
			int p = (r->popdata == 1) ? d->popindx[j] : 0;
			fprintf(f, "%d\t", j + 1);
			if (r->label == 1) fprintf(f, "%s\t", d->indvname[j]);
			fprintf(f, "(%d)\t", d->missvec[j]);
			if (r->popdata == 1) fprintf(f, "%d : ", d->popindx[j]);
			else fprintf(f, "1 : ");
			for (k = 0; k < K; k++) {
				double q = c->qq[(size_t)j * K + k];
				acc[(size_t)p * K + k] += q;
				if (r->distr_fmt == 0) fprintf(f, "\t%.3f\t%.3f\t", q, c->qq2[(size_t)j * K + k] - q * q);
				else if (r->distr_fmt == 1) fprintf(f, "\t%.3f", q);
			}
			cnt[p]++;
			fprintf(f, "\n");
		}
		fprintf(f, "\n\n\nThe index and name of pre-defined populations:\n");
		if (r->popdata == 1) for (i = 0; i < pc; i++) fprintf(f, "%d %s\n", i, d->poptype[i]);
		else fprintf(f, "1\n");
		fprintf(f, "\n\nProportion of membership of each pre-defined population in each of the %d clusters\n", K);
		fprintf(f, "Given Pop\tInferred Clusters\t\tNumber ofIndividuals\n");
		fprintf(f, "    \t\t");
		for (i = 0; i < K; i++) fprintf(f, "%d    ", i + 1);
		fprintf(f, "\n");
		for (i = 0; i < pc; i++) {
			fprintf(f, "%d:\t", i);
			/* population-selfing modes list the clusters in the order of the selfing-rate table (:299) */
			for (j = 0; j < K; j++) {
				int col = ((r->ploid == 2 && (r->mode == 2 || r->mode == 4)) || r->ploid == 4) ? ord[j] : j;
				fprintf(f, "%.3f ", acc[(size_t)i * K + col] / cnt[i]);
			}
			fprintf(f, "\t%d\n", cnt[i]);
		}
		fprintf(f, "\n");
		free(acc);
		free(cnt);
	}

	/* ---- print_P_to_file :313-371 (diploid, -pf 1) */
	if (r->print_freq == 1 && r->ploid == 2 && c->freq) {
		int A = d->allelenum_max, L = r->locinum;
		fprintf(f, "\n\n\nEstimated allele frequencies:\n");
		fprintf(f, "\nLocus_ID\t");
		if (r->markername_flag == 1) fprintf(f, "Marker Name\t");
		fprintf(f, "Alleletype\t");
		for (j = 0; j < K; j++) fprintf(f, "Cluster %d:Mean\tVar\t\t", j + 1);
		fprintf(f, "\n");
		for (j = 0; j < L; j++) {
			for (i = 0; i < d->allelenum[j]; i++) {
				if (i == 0) {
					fprintf(f, "%d\t", j + 1);
					if (r->markername_flag == 1) fprintf(f, "%s\t", d->marker_names[j]);
				} else {
					fprintf(f, "\t");
					if (r->markername_flag == 1) fprintf(f, "\t");
				}
				fprintf(f, "%s\t", d->alleletype[j][i]);
				for (k = 0; k < K; k++) {
					int kk = (r->mode == 2 || r->mode == 4) ? ord[k] : k;
					double m = c->freq[((size_t)kk * L + j) * A + i];
					fprintf(f, "\t%.3f\t%.3f\t", m, c->freq2[((size_t)kk * L + j) * A + i] - m * m);
				}
				fprintf(f, "\n");
			}
			fprintf(f, "\n");
		}
	}
	fclose(f);
	free(ord);
	return 0;
}

/* Gelman-Rubin on [n_chain][n] traces: R = V/W, V = W (n-1)/n + B/n (check_converg.c:100-153) */
double wr_gelman_rubin(const double *tr, int m, int n)
{
	int i, j;
	double psi = 0, W = 0, B = 0;
	double *mu = (double *)calloc((size_t)m, sizeof(double));
	for (i = 0; i < m; i++) {
		for (j = 0; j < n; j++) mu[i] += tr[(size_t)i * n + j];
		mu[i] /= n;
		psi += mu[i];
	}
	psi /= m;
	for (i = 0; i < m; i++) {
		double s = 0;
		for (j = 0; j < n; j++) s += (tr[(size_t)i * n + j] - mu[i]) * (tr[(size_t)i * n + j] - mu[i]);
		W += s / (n - 1);
	}
	W /= m;
	for (i = 0; i < m; i++) B += (mu[i] - psi) * (mu[i] - psi);
	B = B * n / (m - 1);
	free(mu);
	return (W * (n - 1) / n + B / n) / W;
}

/* chain_converg, check_converg.c:44-91 */
int wr_convergence(const char *path, const double *convg_ld, int n_chain, int ckrep, const char *convgfile, int ref_compat)
{
	int k, flag = 0;
	FILE *f = fopen(path, "a+");
	if (!f) return -1;
	if (n_chain == 1) fprintf(f, "There is only one MCMC. No need to check the convergence.\n");
	else {
		/* ref_compat: what the reference computes -- n_chain segments of chain 0's trace (App. B #3) */
		double gr = ref_compat ? wr_gelman_rubin(convg_ld, n_chain, ckrep / n_chain) : wr_gelman_rubin(convg_ld, n_chain, ckrep);
		fprintf(stdout, "The Gelman-Rubin statistics of log-likelihood is %f\n", gr);
		fprintf(f, "\n\nThe Gelman-Rubin statistics for the convergence of log-likelihood is %f.\n", gr);
		if (gr > 1.1) flag = 1;
	}
	fclose(f);
	if (convgfile) {
		if ((f = fopen(convgfile, "w")) == NULL) return -1;
		fprintf(f, "Values of log-likelihood:\n");
		for (k = 0; k < ckrep * n_chain; k++) fprintf(f, k == 0 ? "%f " : " %f ", convg_ld[k]);
		fprintf(f, "\n");
		fclose(f);
	}
	return flag;
}

int wr_rate_convergence(const char *path, const double *tr, const double *const *qq, int n_chain, int ckrep, int K, int N, const char *what)
{
	int c, a, b, i, j, flag = 0;
	int *perm = (int *)calloc((size_t)n_chain * K, sizeof(int));          /* perm[c][a] = cluster of chain c matched to cluster a of chain 0 */
	double *ov = (double *)calloc((size_t)K * K, sizeof(double)), *one = (double *)calloc((size_t)n_chain * ckrep, sizeof(double));
	FILE *f = path ? fopen(path, "a+") : NULL;
	for (a = 0; a < K; a++) perm[a] = a;
	for (c = 1; c < n_chain; c++) {
		char *ua = (char *)calloc((size_t)K, 1), *ub = (char *)calloc((size_t)K, 1);
		for (a = 0; a < K * K; a++) ov[a] = 0;
		for (i = 0; i < N; i++)
			for (a = 0; a < K; a++)
				for (b = 0; b < K; b++) ov[a * K + b] += qq[0][(size_t)i * K + a] * qq[c][(size_t)i * K + b];
		for (j = 0; j < K; j++) {                                         /* greedy assignment, best remaining pair first */
			int ba = -1, bb = -1;
			for (a = 0; a < K; a++) if (!ua[a]) for (b = 0; b < K; b++) if (!ub[b] && (ba < 0 || ov[a * K + b] > ov[ba * K + bb])) { ba = a; bb = b; }
			perm[c * K + ba] = bb; ua[ba] = 1; ub[bb] = 1;
		}
		free(ua); free(ub);
	}
	for (a = 0; a < K; a++) {
		double gr;
		for (c = 0; c < n_chain; c++)
			for (j = 0; j < ckrep; j++) one[(size_t)c * ckrep + j] = tr[((size_t)c * ckrep + j) * K + perm[c * K + a]];
		gr = wr_gelman_rubin(one, n_chain, ckrep);
		fprintf(stdout, "The Gelman-Rubin statistics of the %s of cluster %d is %f\n", what, a + 1, gr);
		if (f) fprintf(f, "The Gelman-Rubin statistics for the convergence of the %s of cluster %d is %f.\n", what, a + 1, gr);
		if (gr > 1.1) flag = 1;
	}
	if (f) fclose(f);
	free(perm); free(ov); free(one);
	return flag;
}
