/*
 * inbreed.c -- the host program: InStruct's command line, input parsing and result file,
 * with the MCMC itself running on B200 GPUs through the C-ABI of include/instruct_b200.h.
 *
 * It is the from-scratch counterpart of the reference's driver (InStruct.c:152-202): same
 * flags (InStruct.c:251-432; README synopsis calls the binary `inbreed`), same reader
 * semantics (genostore.c), same `-o` file (writer.c).  Where the reference calls
 * mcmc_updating() (InStruct.c:184) this program calls ig_run_chain() on a context that keeps
 * the packed genotype store resident in HBM across chains.
 *
 * New flags (none collides with the reference's 35 spellings):
 *   --gpus N                 use N GPUs of this node (default 1)
 *   --shard chains|individuals
 *                            chains: chain c runs on GPU c mod N, no communication (default);
 *                            individuals: every chain is sharded over the N GPUs by individuals,
 *                            one NCCL int32 all-reduce + one all-gather per sweep
 *   --ref-compat-gr          print the Gelman-Rubin value exactly as the reference computes it
 *   --quiet-data             do not echo the recoded genotype matrix to stdout
 *   --save-store FILE        also write the packed genotype store (int16 matrix + name tables) to FILE
 *   --load-store FILE        read the packed store instead of parsing the text file (-d is then only echoed)
 *   --pack-only              stop after --save-store (text -> packed conversion; needs no GPU)
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "../../include/instruct_b200.h"
#include "genostore.h"
#include "writer.h"

/* ---- run parameters, defaults as InStruct.c:25-65 ------------------------------------ */
static double siglevel = 0.900, alpha_dpm = 10, max_mem = 1.0e9;
static int nloci = 100, popnum = 2, totalsize = 100, ploid = 2, thinning = 10, ckrep = 20, GR_flag = 1, chainnum = 2;
static long updatenum = 1000000, burnin = 500000;
static int gr_params = 0;       /* --gr-params 1: also append the per-parameter Gelman-Rubin lines to the -o file (not in the reference's format) */
static const char *missingdata = "-9", *datafilename, *outfilename, *initialfilename, *convgfilename;
static int label = 1, popdata = 1, prior_flag = 0, back_refl = 1, type_freq = 1, nstep_check_empty_cluster = 20;
static int n_extra_col = 0, markername_flag = 0, print_iter = 1, print_freq = 0, n_small = 1, n_large = 0, inf_K = 0;
static int distr_fmt = 1, autopoly = 1, data_fmt = 0, mode = 1;
static long seeds[3] = {13, 4, 1972};               /* random.c:10-12 */
static int n_gpus = 1, shard_individuals = 0, ref_compat_gr = 0, quiet_data = 0, pack_only = 0;
static int chains_per_gpu = 0;    /* --chains-per-gpu C: chains run concurrently on one GPU (0 = choose: several for small data sets, whose sweeps
                                   * are launch- and latency-bound and leave most of the SMs idle; 1 for HBM-sized ones) */
static int n_workers = 1;
static const char *save_store = NULL, *load_store = NULL;

static void die(const char *msg)                    /* nrerror's output convention, nrutil.c:9-16 */
{
	fprintf(stdout, "ERROR: \n%s\n...now exiting to system...\n", msg);
	exit(1);
}

static void parse_args(int argc, char **argv)
{
	static const char *synopsis =
	    "Synopsis:\n\tinbreed -d data_file -o output_file [-i initial_file] [-K population number] [-L loci number] "
	    "[-N total individual number] [-p ploid] [-u iteration number] [-b burn-in number] [-m missingdata] [-t thinning] "
	    "[-c chain number] [-s seed1 seed2 seed3] [-sl significance level] [-lb label] [-a popdata] [-g GR_flag] [-r ckrep] "
	    "[-f prior_flag] [-v mode] [-h alpha_dpm] [-e back_refl] [-y type_freq] [-j nstep_check_empty_cluster] "
	    "[-x extra_columns] [-w markername] [-cf convgfilename] [-pi print_iter] [-pf print_freq]  [-ik inf_K] "
	    "[-kv n_small n_large] [-df distr_fmt] [-ap autopoly] [-af data_fmt] [-mm max_mem] "
	    "[--gpus N] [--shard chains|individuals] [--chains-per-gpu C] [--ref-compat-gr] [--gr-params 0|1] [--quiet-data] "
	    "[--save-store file] [--load-store file] [--pack-only]\n";
	int i;
	if (argc == 2 && strcmp(argv[1], "-h") == 0) { fprintf(stdout, "%s", synopsis); exit(1); }
	if (argc < 5) die("Too few arguments in the command line!");
#define ARG(flag) (strcmp(argv[i], flag) == 0 && i + 1 < argc)
	for (i = 1; i < argc; i++) {
		if (ARG("-d")) datafilename = argv[i + 1];
		else if (ARG("-o")) outfilename = argv[i + 1];
		else if (ARG("-i")) initialfilename = argv[i + 1];
		else if (ARG("-cf")) convgfilename = argv[i + 1];
		else if (ARG("--gr-params")) gr_params = atoi(argv[i + 1]);
		else if (ARG("-L")) nloci = atoi(argv[i + 1]);
		else if (ARG("-N")) totalsize = atoi(argv[i + 1]);
		else if (ARG("-K")) popnum = atoi(argv[i + 1]);
		else if (ARG("-p")) ploid = atoi(argv[i + 1]);
		else if (ARG("-u")) updatenum = atol(argv[i + 1]);
		else if (ARG("-b")) { burnin = atol(argv[i + 1]); if (burnin == 0) die("Burn-in should not be zero!"); }
		else if (ARG("-t")) thinning = atoi(argv[i + 1]);
		else if (ARG("-c")) chainnum = atoi(argv[i + 1]);
		else if (ARG("-m")) missingdata = argv[i + 1];
		else if (ARG("-lb")) label = atoi(argv[i + 1]);
		else if (ARG("-a")) popdata = atoi(argv[i + 1]);
		else if (ARG("-g")) GR_flag = atoi(argv[i + 1]);
		else if (ARG("-f")) prior_flag = atoi(argv[i + 1]);
		else if (ARG("-v")) mode = atoi(argv[i + 1]);
		else if (ARG("-r")) ckrep = atoi(argv[i + 1]);
		else if (ARG("-e")) back_refl = atoi(argv[i + 1]);
		else if (ARG("-y")) type_freq = atoi(argv[i + 1]);
		else if (ARG("-x")) n_extra_col = atoi(argv[i + 1]);
		else if (ARG("-pi")) print_iter = atoi(argv[i + 1]);
		else if (ARG("-ap")) autopoly = atoi(argv[i + 1]);
		else if (ARG("-pf")) print_freq = atoi(argv[i + 1]);
		else if (ARG("-w")) markername_flag = atoi(argv[i + 1]);
		else if (ARG("-af")) data_fmt = atoi(argv[i + 1]);
		else if (ARG("-mm")) max_mem = atof(argv[i + 1]);
		else if (ARG("-ik")) inf_K = atoi(argv[i + 1]);
		else if (strcmp(argv[i], "-kv") == 0 && i + 2 < argc) { n_small = atoi(argv[i + 1]); n_large = atoi(argv[i + 2]); }
		else if (ARG("-df")) distr_fmt = atoi(argv[i + 1]);
		else if (ARG("-sl")) siglevel = atof(argv[i + 1]);
		else if (ARG("-h")) alpha_dpm = atof(argv[i + 1]);
		else if (ARG("-j")) nstep_check_empty_cluster = atoi(argv[i + 1]);
		else if (strcmp(argv[i], "-s") == 0 && i + 3 < argc) { seeds[0] = atol(argv[i + 1]); seeds[1] = atol(argv[i + 2]); seeds[2] = atol(argv[i + 3]); }
		else if (ARG("--gpus")) n_gpus = atoi(argv[i + 1]);
		else if (ARG("--chains-per-gpu")) chains_per_gpu = atoi(argv[i + 1]);
		else if (ARG("--shard")) shard_individuals = (strcmp(argv[i + 1], "individuals") == 0);
		else if (strcmp(argv[i], "--ref-compat-gr") == 0) ref_compat_gr = 1;
		else if (strcmp(argv[i], "--quiet-data") == 0) quiet_data = 1;
		else if (ARG("--save-store")) save_store = argv[i + 1];
		else if (ARG("--load-store")) load_store = argv[i + 1];
		else if (strcmp(argv[i], "--pack-only") == 0) pack_only = 1;
	}
#undef ARG
	if (!datafilename || !outfilename) die("Both -d data_file and -o output_file are required!");
	if (ckrep > (updatenum - burnin) / thinning)          /* InStruct.c:437-444 */
		die("The number of iterations for convergence assessment is greater than the total number of retained iterations from MCMC.");
	if (nstep_check_empty_cluster > (updatenum - burnin) / thinning)
		die("The number of iterations for checking the existence of empty cluster is greater than the total number of retained iterations from MCMC.");
}

/* ---- starting values and chain names: read_init(), initial.c:38-135.  Without -i the
 * library draws the starting selfing rates itself (pass NULL); names are Chain#n. --------- */
typedef struct { float *initd; char **name; int *have; } init_t;

static init_t read_init(const char *path, int nchain, int K)
{
	init_t in;
	int i, k;
	in.initd = (float *)calloc((size_t)nchain * K, sizeof(float));
	in.name = (char **)calloc((size_t)nchain, sizeof(char *));
	in.have = (int *)calloc((size_t)nchain, sizeof(int));
	for (i = 0; i < nchain; i++) { in.name[i] = (char *)malloc(100); snprintf(in.name[i], 100, "Chain#%d", i + 1); }
	if (path) {
		FILE *f = fopen(path, "r");
		char line[1000];
		int c = 0;
		if (!f) die("Cannot open inital file!");
		while (c < nchain && fgets(line, sizeof line, f)) {
			if (line[0] != '>') continue;
			line[strcspn(line, "\r\n")] = 0;
			snprintf(in.name[c], 100, "%s", line + 1);
			if (!fgets(line, sizeof line, f)) break;
			{
				char *p = line, *e;
				for (k = 0; k < K; k++) {
					double v = strtod(p, &e);
					if (e == p) die("The number of initial values for selfing rates is not equal the number of subpopulation assumed!\n");
					in.initd[(size_t)c * K + k] = (float)v;
					p = e;
				}
			}
			in.have[c] = 1;
			c++;
		}
		fclose(f);
	}
	return in;
}

/* ---- one worker per GPU -------------------------------------------------------------------- */
typedef struct {
	double totallkh, totallkh2, *indvlkh, *qq, *qq2, *self, *self2, *gen, *gen2, *freq, *freq2;
	int done;
} chain_out;

typedef struct {
	int gpu, rank, nrank;
	const gs_store *gs;
	const init_t *init;
	chain_out *out;         /* [chainnum] */
	double *convg;          /* [chainnum][ckrep] */
	double *convgS;         /* [chainnum][ckrep][K] population-rate traces (modes 2, 4, ploid 4), or NULL */
	unsigned char nccl_id[128];
	int status;
	char err[512];
} worker_t;

static ig_config base_config(const gs_store *gs)
{
	ig_config c;
	memset(&c, 0, sizeof c);
	c.ploid = ploid; c.popnum = popnum; c.locinum = gs->locinum; c.totalsize = gs->totalsize;
	c.mode = mode; c.prior_flag = prior_flag; c.back_refl = back_refl; c.type_freq = type_freq; c.alpha_dpm = alpha_dpm;
	c.nstep_check_empty_cluster = nstep_check_empty_cluster; c.print_iter = print_iter; c.print_freq = print_freq;
	c.autopoly = autopoly; c.update = updatenum; c.burnin = burnin; c.thinning = thinning; c.ckrep = GR_flag ? ckrep : 0;
	c.seed = (uint64_t)seeds[0] | ((uint64_t)seeds[1] << 21) | ((uint64_t)seeds[2] << 42);
	c.shard_count = 1;
	return c;
}

static void alloc_out(chain_out *o, int N, int K, int ns, size_t nfreq)
{
	o->indvlkh = (double *)calloc((size_t)N, 8); o->qq = (double *)calloc((size_t)N * K, 8); o->qq2 = (double *)calloc((size_t)N * K, 8);
	o->self = (double *)calloc((size_t)ns, 8); o->self2 = (double *)calloc((size_t)ns, 8);
	o->gen = (double *)calloc((size_t)N, 8); o->gen2 = (double *)calloc((size_t)N, 8);
	if (nfreq) { o->freq = (double *)calloc(nfreq, 8); o->freq2 = (double *)calloc(nfreq, 8); }
}

static void *worker(void *arg)
{
	worker_t *w = (worker_t *)arg;
	const gs_store *gs = w->gs;
	const int N = gs->totalsize, K = popnum, L = gs->locinum;
	ig_config cfg = base_config(gs);
	ig_ctx *ctx = NULL;
	int16_t *xs = NULL;
	const int16_t *x = gs->x;
	int chn;
	cfg.device = w->gpu;
	if (w->nrank > 1) {                                   /* individual-sharded: slice the store */
		const int cap = (N + w->nrank - 1) / w->nrank, b = w->rank * cap, n = (N - b < cap) ? N - b : cap;
		int l;
		cfg.shard_rank = w->rank; cfg.shard_count = w->nrank; cfg.shard_begin = b; cfg.shard_size = n;
		xs = (int16_t *)malloc((size_t)L * n * ploid * sizeof(int16_t));
		for (l = 0; l < L; l++) memcpy(xs + (size_t)l * n * ploid, gs->x + ((size_t)l * N + b) * ploid, (size_t)n * ploid * sizeof(int16_t));
		x = xs;
		if (w->rank != 0) cfg.print_iter = 0;
	}
	w->status = ig_create(&cfg, &ctx);
	if (w->status == IG_OK) w->status = ig_load_genotypes(ctx, x, gs->allelenum);
	if (w->status == IG_OK && w->nrank > 1) w->status = ig_comm_init(ctx, w->nccl_id);
	free(xs);
	for (chn = 0; w->status >= 0 && chn < chainnum; chn++) {
		chain_out *o = &w->out[chn];
		ig_chain_result r;
		int attempt = 0;
		if (w->nrank == 1 && chn % n_workers != w->rank) continue;  /* chains mode: chain c on worker c mod W, worker w on GPU w mod N */
		memset(&r, 0, sizeof r);
		r.indvlkh = o->indvlkh; r.qq = o->qq; r.qq2 = o->qq2; r.self_rates = o->self; r.self_rates2 = o->self2;
		r.gen = o->gen; r.gen2 = o->gen2; r.freq = o->freq; r.freq2 = o->freq2;
		for (;;) {                                        /* empty-cluster retry, InStruct.c:185-190 */
			if (w->rank == 0 || w->nrank == 1) fprintf(stdout, "\n\n%s Starts:\n", w->init->name[chn]);
			w->status = ig_run_chain(ctx, chn + 1000 * attempt, w->init->have[chn] ? w->init->initd + (size_t)chn * K : NULL,
			                         &r, w->convg ? w->convg + (size_t)chn * ckrep : NULL);
			if (w->status != IG_EMPTY_CLUSTER) break;
			fprintf(stdout, "Chain %d has an empty cluster, thus discarded!\n", chn + 1);
			attempt++;
		}
		if (w->status < 0) break;
		if (w->convgS && (w->rank == 0 || w->nrank == 1)) {
			int32_t rows = 0;
			if (ig_get_rate_trace(ctx, w->convgS + (size_t)chn * ckrep * K, (size_t)ckrep * K * sizeof(double), &rows) != IG_OK) w->convgS = NULL;
		}
		if (r.step != r.steps) { w->status = IG_ERR_STATE; snprintf(w->err, sizeof w->err, "The number of iterations attained is not the same as counted"); break; }
		o->totallkh = r.totallkh; o->totallkh2 = r.totallkh2; o->done = 1;
		if (w->rank == 0 || w->nrank == 1) fprintf(stdout, "\n\nChain %d is finished running.\n", chn + 1);
	}
	if (w->status < 0 && !w->err[0]) snprintf(w->err, sizeof w->err, "%s", ig_last_error());
	ig_destroy(ctx);
	return NULL;
}

/* all chains of one K: run (chains over GPUs, or one chain's individuals over GPUs), write the
 * chain tables in chain order, the Gelman-Rubin line; returns the smallest DIC over the chains */
static double run_for_K(const gs_store *gsp, int K, wr_run *runp, const wr_data *wdp)
{
	const gs_store gs = *gsp;
	wr_run run = *runp;
	const wr_data wd = *wdp;
	init_t init;
	chain_out *outs;
	double *convg = NULL, *convgS = NULL, dic_min = 0;
	worker_t *ws;
	pthread_t *th;
	int g, chn, N = gs.totalsize, ns;
	size_t nfreq;

	popnum = K;
	run.popnum = K;
	ns = ((mode == 3 || mode == 5) && ploid == 2) ? N : K;
	init = read_init(initialfilename, chainnum, K);
	nfreq = print_freq ? (size_t)K * gs.locinum * gs.allelenum_max : 0;
	outs = (chain_out *)calloc((size_t)chainnum, sizeof(chain_out));
	for (chn = 0; chn < chainnum; chn++) alloc_out(&outs[chn], N, K, ns, nfreq);
	if (GR_flag) convg = (double *)calloc((size_t)chainnum * ckrep, sizeof(double));
	if (GR_flag && ns == K && (ploid == 4 || mode == 2 || mode == 4)) convgS = (double *)calloc((size_t)chainnum * ckrep * K, sizeof(double));
	/* chains mode: n_workers = GPUs x chains per GPU host threads, each with its own context (its own copy of the store, its
	 * own stream), so that several chains' sweeps are in flight on one GPU; sharded mode: one worker per GPU */
	n_workers = n_gpus;
	if (!shard_individuals) {
		int cpg = chains_per_gpu;
		if (cpg <= 0) cpg = ((double)gs.totalsize * gs.locinum <= 2.0e6 && print_iter == 0) ? 8 : 1;   /* progress lines of concurrent chains would interleave */
		if (cpg > 16) cpg = 16;
		n_workers = n_gpus * cpg;
		if (n_workers > chainnum) n_workers = chainnum;
		if (n_workers < 1) n_workers = 1;
	}
	ws = (worker_t *)calloc((size_t)n_workers, sizeof(worker_t));
	th = (pthread_t *)calloc((size_t)n_workers, sizeof(pthread_t));
	for (g = 0; g < n_workers; g++) {
		ws[g].gpu = g % n_gpus; ws[g].rank = g; ws[g].nrank = shard_individuals ? n_gpus : 1;
		ws[g].gs = &gs; ws[g].init = &init; ws[g].convg = convg; ws[g].convgS = convgS;
		/* sharded ranks all compute the full moments; only rank 0's copy is kept */
		if (shard_individuals && g > 0) {
			ws[g].out = (chain_out *)calloc((size_t)chainnum, sizeof(chain_out));
			for (chn = 0; chn < chainnum; chn++) alloc_out(&ws[g].out[chn], N, K, ns, nfreq);
			ws[g].convg = GR_flag ? (double *)calloc((size_t)chainnum * ckrep, sizeof(double)) : NULL;
		} else ws[g].out = outs;
	}
	if (shard_individuals && n_gpus > 1) {
		if (ig_comm_unique_id(ws[0].nccl_id) != IG_OK) die(ig_last_error());
		for (g = 1; g < n_gpus; g++) memcpy(ws[g].nccl_id, ws[0].nccl_id, 128);
	}
	for (g = 0; g < n_workers; g++) pthread_create(&th[g], NULL, worker, &ws[g]);
	for (g = 0; g < n_workers; g++) pthread_join(th[g], NULL);
	for (g = 0; g < n_workers; g++) if (ws[g].status < 0) die(ws[g].err);

	for (chn = 0; chn < chainnum; chn++) {               /* chain_stat in chain order, InStruct.c:191 */
		wr_chain_t c;
		chain_out *o = &outs[chn];
		memset(&c, 0, sizeof c);
		c.chn_name = init.name[chn]; c.name_len = (int)strlen(init.name[chn]) + 1;
		c.totallkh = o->totallkh; c.totallkh2 = o->totallkh2; c.indvlkh = o->indvlkh; c.qq = o->qq; c.qq2 = o->qq2;
		c.self_rates = o->self; c.self_rates2 = o->self2; c.gen = o->gen; c.gen2 = o->gen2; c.freq = o->freq; c.freq2 = o->freq2;
		{
			double dic = 0;
			if (wr_chain(outfilename, &run, &wd, &c, &dic)) die("Cannot open output file!");
			if (chn == 0 || dic < dic_min) dic_min = dic;
		}
	}
	if (GR_flag == 1 && wr_convergence(outfilename, convg, chainnum, ckrep, convgfilename, ref_compat_gr) < 0)
		die("ERROR: Cannot open output file!\n");
	if (GR_flag == 1 && convgS && chainnum > 1) {
		/* per-parameter convergence: the reference's statistic (check_converg.c:100-153) on the
		 * trace of every population rate, clusters of chain c matched to chain 0's by the overlap of their posterior mean Q */
		const double **qq = (const double **)calloc((size_t)chainnum, sizeof(double *));
		for (chn = 0; chn < chainnum; chn++) qq[chn] = outs[chn].qq;
		wr_rate_convergence(gr_params ? outfilename : NULL, convgS, qq, chainnum, ckrep, K, N, (ploid == 2 && mode == 4) ? "inbreeding coefficient" : "selfing rate");
		free(qq);
	}
	free(ws); free(th); free(outs); free(convg); free(convgS);
	return dic_min;
}

int main(int argc, char **argv)
{
	gs_options go;
	gs_store gs;
	char err[512];
	wr_run run;
	wr_data wd;
	double memreq;
	int N, K, ns, ndev;

	parse_args(argc, argv);
	memset(&go, 0, sizeof go);
	go.ploid = ploid; go.totalsize = totalsize; go.locinum = nloci; go.missing = missingdata; go.label = label;
	go.popdata = popdata; go.n_extra_col = n_extra_col; go.markername_flag = markername_flag; go.datafmt = data_fmt;
	go.quiet = quiet_data;
	if (ploid != 2 && ploid != 4) die("ploid must be 2 or 4");
	if (ploid == 4 && autopoly != 1 && autopoly != 0) die("-ap must be 1 (autotetraploid) or 0 (allotetraploid)");
	/* --load-store: the packed store a previous run wrote with --save-store replaces the text reader */
	if (load_store) {
		if (gs_load(load_store, &gs, err, sizeof err)) die(err);
		if (gs.ploid != ploid) die("--load-store: the packed store was written for another ploidy (-p)");
		fprintf(stdout, "Packed genotype store %s: %d individuals, %d polymorphic loci of %d.\n", load_store, gs.totalsize, gs.locinum, gs.locinum_file);
	} else if (gs_read(datafilename, &go, &gs, err, sizeof err)) die(err);
	if (save_store && gs_save(save_store, &gs, err, sizeof err)) die(err);
	if (pack_only) {                                     /* text -> packed store, no chain (needs no GPU) */
		if (!save_store) die("--pack-only needs --save-store file");
		fprintf(stdout, "Packed genotype store written to %s\n", save_store);
		gs_free(&gs);
		return 0;
	}
	N = gs.totalsize; K = popnum; ns = ((mode == 3 || mode == 5) && ploid == 2) ? N : K;
	/* mem_cal, InStruct.c:204-225 (the estimate is the reference's; kept for its two log lines) */
	memreq = (print_freq ? 8.0 * K * gs.locinum * gs.allelenum_max : 0.0) + 8.0 + 8.0 * N + 8.0 * ns + 4.0 * N + 8.0 * N * K;
	memreq *= (double)((updatenum - burnin) / thinning);
	fprintf(stdout, "The memory required for this run is %f \n", memreq);
	fprintf(stdout, "The maximum memory allowed is %f \n", max_mem);

	memset(&run, 0, sizeof run);
	run.datafilename = datafilename; run.initialfilename = initialfilename; run.missingdata = missingdata;
	run.chainnum = chainnum; run.thinning = thinning; run.ploid = ploid; run.autopoly = autopoly; run.totalsize = N;
	run.locinum = gs.locinum; run.popnum = K; run.mode = mode; run.inf_K = inf_K; run.prior_flag = prior_flag;
	run.back_refl = back_refl; run.print_freq = print_freq; run.GR_flag = GR_flag; run.ckrep = ckrep; run.distr_fmt = distr_fmt;
	run.label = label; run.popdata = popdata; run.markername_flag = markername_flag; run.update = updatenum; run.burnin = burnin;
	run.siglevel = siglevel; run.alpha_dpm = alpha_dpm;
	if (wr_banner(outfilename, argc, argv, &run)) die("Cannot open output file!");

	ndev = ig_device_count();
	if (ndev < 1) die("no CUDA device: inbreed has no CPU path");
	if (n_gpus < 1) n_gpus = 1;
	if (n_gpus > ndev) { fprintf(stdout, "Only %d GPU(s) visible; using %d.\n", ndev, ndev); n_gpus = ndev; }
	if (!shard_individuals && n_gpus > chainnum) n_gpus = chainnum;

	memset(&wd, 0, sizeof wd);
	wd.indvname = gs.indvname; wd.poptype = gs.poptype; wd.marker_names = gs.marker_names; wd.alleletype = gs.alleletype;
	wd.popindx = gs.popindx; wd.missvec = gs.missvec; wd.allelenum = gs.allelenum; wd.pop_count = gs.pop_count;
	wd.allelenum_max = gs.allelenum_max;
	if (inf_K == 1) {
		/* inf_K_val, InStruct.c:536-601: every K of the range, all chains each, the K whose best
		 * chain has the smallest DIC wins.  (K, chain) pairs are independent: chains spread over GPUs. */
		int Kb, K_best = 0;
		double best = 0;
		FILE *f;
		if (n_large < 1 || n_small < 1 || n_small > n_large) {
			n_small = 1;
			n_large = (int)pow((double)N, 0.3) + 1;
			fprintf(stdout, "The range of value for K is not correct! Change to default value (%d - %d)!\n", n_small, n_large);
		}
		if (n_large > 16) die("K inference: the library supports K <= 16");
		for (Kb = n_small; Kb <= n_large; Kb++) {
			double d;
			if ((f = fopen(outfilename, "a+")) == NULL) die("Cannot open output file!");
			fprintf(f, "\n\nThe current K is %d\n", Kb);
			fclose(f);
			d = run_for_K(&gs, Kb, &run, &wd);
			if (Kb == n_small || d < best) { best = d; K_best = Kb; }
		}
		if ((f = fopen(outfilename, "a+")) == NULL) die("Cannot open output file!");
		fprintf(f, "\n\nThe range of value for K is (%d - %d)!\n", n_small, n_large);
		fprintf(f, "The optimal K is %d\n", K_best);
		fclose(f);
	} else run_for_K(&gs, popnum, &run, &wd);
	fprintf(stdout, "THE JOB IS SUCCESSFULLY FINISHED\n");
	gs_free(&gs);
	return 0;
}
