/* writer.h -- result-file writer of the inbreed host program (see writer.c). */
#ifndef WRITER_H
#define WRITER_H

typedef struct wr_run {            /* the run parameters printinfo() echoes (InStruct.c:450-531) */
	const char *datafilename, *initialfilename, *missingdata;
	int chainnum, thinning, ploid, autopoly, totalsize, locinum, popnum, mode, inf_K, prior_flag, back_refl;
	int print_freq, GR_flag, ckrep, distr_fmt, label, popdata, markername_flag;
	long update, burnin;
	double siglevel, alpha_dpm;
} wr_run;

typedef struct wr_data {           /* what the tables need from the genotype store */
	char **indvname, **poptype, **marker_names;
	char ***alleletype;
	const int *popindx, *missvec, *allelenum;
	int pop_count, allelenum_max;
} wr_data;

typedef struct wr_chain_t {        /* CHAIN (mcmc.h:29-53) as flat arrays */
	const char *chn_name;
	int name_len;                  /* strlen + 1, like INIT.name_len (initial.c:65) */
	double totallkh, totallkh2;
	const double *indvlkh, *qq, *qq2, *self_rates, *self_rates2, *gen, *gen2, *freq, *freq2;
} wr_chain_t;

int wr_banner(const char *path, int argc, char **argv, const wr_run *r);
int wr_chain(const char *path, const wr_run *r, const wr_data *d, const wr_chain_t *c, double *dic_out);
double wr_gelman_rubin(const double *traces, int n_chain, int n);
int wr_convergence(const char *path, const double *convg_ld, int n_chain, int ckrep, const char *convgfile, int ref_compat);

/* Gelman-Rubin per population rate on the first ckrep retained draws of every chain (tr: [n_chain][ckrep][K]).  Clusters are
 * matched across chains first (label switching): cluster b of chain c goes with the cluster a of chain 0 it shares most
 * posterior membership with, sum_i qq_0[i][a] * qq_c[i][b], greedily from the best pair down.  Printed to stdout, and
 * appended to `path` when it is not NULL.  Returns 1 when some statistic exceeds 1.1. */
int wr_rate_convergence(const char *path, const double *tr, const double *const *qq, int n_chain, int ckrep, int K, int N, const char *what);

#endif
