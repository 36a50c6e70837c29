/*
 * mcmc_gpu.c -- the reference-side binding of INTEGRATION.md, as a real file.
 *
 * Compiled TOGETHER WITH the unmodified sources of slowkoni/InStruct (never copied into this repo: oracle/Makefile
 * compiles them where they lie), it replaces mcmc_updating() (mcmc.h:56, mcmc.c:63-87; sole call sites
 * InStruct.c:184,565) by a call into libinstruct_b200.so.  mcmc.c stays in the link for allocate / free helpers
 * (free_chain, mcmc.c:740) with its own mcmc_updating renamed on the command line (-Dmcmc_updating=mcmc_updating_cpu).
 * The result is the reference PROGRAM -- its flag parser, text reader, chain_stat / chain_converg writer, untouched --
 * running every chain on the GPU: `oracle/_ref/InStruct_b200` (test infrastructure; the product's own host program
 * is instruct_b200/host/inbreed).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "instruct_b200.h"
#include "nrutil.h"
#include "random.h"
#include "data_interface.h"
#include "initial.h"
#include "check_converg.h"
#include "mcmc.h"

extern int GR_flag;     /* InStruct.c:32: with -g 0 the caller never runs allocate_convg (InStruct.c:181), so *cvg is uninitialised stack */

/* SEQDATA.seqdata[i][l][c] (int; -9 = missing, data_interface.c:497,538; ploid 4: ascending distinct alleles padded with
 * -1, data_interface.c:617-650)  ->  int16 [L][N][ploid], any negative value = not an allele */
static int16_t *pack_store(SEQDATA d)
{
	int16_t *x = (int16_t *)malloc((size_t)d.locinum * d.totalsize * d.ploid * sizeof *x);
	int l, i, c;
	if (!x) nrerror("allocation failure in pack_store()");
	for (l = 0; l < d.locinum; l++)
		for (i = 0; i < d.totalsize; i++)
			for (c = 0; c < d.ploid; c++)
				x[((size_t)l * d.totalsize + i) * d.ploid + c] = (int16_t)d.seqdata[i][l][c];
	return x;
}

CHAIN mcmc_updating(SEQDATA data, INIT initial, int chn, CONVG *cvg)
{
	CHAIN chain;
	ig_config cfg;
	ig_chain_result r;
	ig_status st;
	int16_t *x;
	double *z_share = NULL;
	const int N = data.totalsize, K = data.popnum;
	const int dip = data.ploid == 2, mode = data.mode;
	int j;
	const int have_cvg = (GR_flag == 1 && cvg != NULL);
	double u_hi, u_lo;

	memset(&chain, 0, sizeof chain);
	memset(&cfg, 0, sizeof cfg);
	memset(&r, 0, sizeof r);
	/* ---- what initial_chn (mcmc.c:476-486) and allocate_chn (mcmc.c:588-642) would have allocated, so that
	 *      chain_stat / free_chain keep working */
	chain.name_len = initial.name_len[chn];
	chain.chn_name = (char *)cvector(0, chain.name_len - 1);
	for (j = 0; j < chain.name_len; j++) chain.chn_name[j] = initial.chn_name[chn][j];
	fprintf(stdout, "\n\n%s Starts:\n", chain.chn_name);
	chain.indvlkh = dvector(0, N - 1);
	r.indvlkh = chain.indvlkh;
	if ((dip && mode != 0) || data.ploid == 4) {
		chain.qq = dmatrix(0, N - 1, 0, K - 1);                 /* one contiguous block behind the row pointers (nrutil.c:92-112) */
		chain.qq2 = dmatrix(0, N - 1, 0, K - 1);
		r.qq = &chain.qq[0][0];
		r.qq2 = &chain.qq2[0][0];
	} else {
		chain.z = lmatrix(0, N - 1, 0, K - 1);                  /* mode 0: counts of retained samples per (individual, cluster) */
		z_share = (double *)calloc((size_t)N * K, sizeof(double));
		r.qq = z_share;
	}
	if (dip && (mode == 2 || mode == 3)) {
		const int ns = mode == 3 ? N : K;
		chain.self_rates = dvector(0, ns - 1); chain.self_rates2 = dvector(0, ns - 1);
		r.self_rates = chain.self_rates; r.self_rates2 = chain.self_rates2;
		chain.gen = dvector(0, N - 1); chain.gen2 = dvector(0, N - 1);
		r.gen = chain.gen; r.gen2 = chain.gen2;
	} else if (dip && (mode == 4 || mode == 5)) {                   /* inbreeding coefficients travel in the self_rates slots */
		const int ns = mode == 5 ? N : K;
		chain.inbreed = dvector(0, ns - 1); chain.inbreed2 = dvector(0, ns - 1);
		r.self_rates = chain.inbreed; r.self_rates2 = chain.inbreed2;
	} else if (data.ploid == 4) {
		chain.self_rates = dvector(0, K - 1); chain.self_rates2 = dvector(0, K - 1);
		r.self_rates = chain.self_rates; r.self_rates2 = chain.self_rates2;
	}
	if (data.print_freq == 1 && dip) {
		chain.freq = d3tensor(0, K - 1, 0, data.locinum - 1, 0, data.allelenum_max - 1);
		chain.freq2 = d3tensor(0, K - 1, 0, data.locinum - 1, 0, data.allelenum_max - 1);
		r.freq = &chain.freq[0][0][0];
		r.freq2 = &chain.freq2[0][0][0];
	}

	cfg.ploid = data.ploid; cfg.popnum = K; cfg.locinum = data.locinum; cfg.totalsize = N;
	cfg.mode = mode; cfg.prior_flag = data.prior_flag; cfg.back_refl = data.back_refl; cfg.type_freq = data.type_freq;
	cfg.alpha_dpm = data.alpha_dpm; cfg.nstep_check_empty_cluster = data.nstep_check_empty_cluster;
	cfg.print_iter = data.print_iter; cfg.print_freq = data.print_freq; cfg.autopoly = data.autopoly;
	cfg.update = initial.update; cfg.burnin = initial.burnin; cfg.thinning = initial.thinning;
	cfg.ckrep = have_cvg ? cvg->ckrep : 0;
	cfg.shard_size = N; cfg.shard_count = 1;
	/* the chain's Philox key from the reference's own stream (-s seed1 seed2 seed3 seeds it, InStruct.c:430): the run
	 * stays a function of the three seeds, and every chain draws a different key */
	u_hi = ran1();                                          /* two statements: the order of the two draws is part of the seed */
	u_lo = ran1();
	cfg.seed = ((uint64_t)(u_hi * 4294967296.0) << 32) | (uint64_t)(u_lo * 4294967296.0);

	x = pack_store(data);
	st = ig_mcmc_updating(&cfg, x, data.allelenum, chn, (dip && (mode == 2 || mode == 4)) || data.ploid == 4 ? initial.initd[chn] : NULL, &r,
	                      have_cvg ? &cvg->convg_ld[chn * cvg->ckrep] : NULL);     /* mcmc.c:223-224 */
	free(x);
	if (st < 0) nrerror((char *)ig_last_error());          /* the reference's own error convention (nrutil.c:9-16): message, exit(1) */
	chain.steps = (long)r.steps; chain.step = (long)r.step;
	chain.totallkh = r.totallkh; chain.totallkh2 = r.totallkh2;
	chain.flag_empty_cluster = (st == IG_EMPTY_CLUSTER);  /* the caller discards the chain and retries, InStruct.c:185-190 */
	if (chain.flag_empty_cluster) fprintf(stdout, "Chain %d has an empty cluster, thus discarded!\n", chn + 1);
	if (z_share) {                                         /* CHAIN.z[i][k] = retained samples with individual i in cluster k (mcmc.c:1356-1362) */
		int i, k;
		for (i = 0; i < N; i++)
			for (k = 0; k < K; k++) chain.z[i][k] = lround(z_share[(size_t)i * K + k] * (double)r.step);
		free(z_share);
	}
	return chain;
}
