/*
 * instruct_b200.h -- C-ABI of the B200-native InStruct sampler.
 *
 * The drop-in boundary is the reference's single hot-path entry point
 *
 *     CHAIN mcmc_updating(SEQDATA data, INIT initial, int chn, CONVG *cvg);
 *                                        -- /root/reference mcmc.h:56, mcmc.c:63-87,
 *                                           sole caller InStruct.c:184 (and :565 under -ik 1)
 *
 * Everything below it (update_P, update_S_POP / update_S_IND / update_DP, update_G,
 * update_ZQ, update_alpha, cal_lkh, store_chn, check_empty_cluster -- mcmc.c:799-1974,
 * DPMM.c:124-398) runs as sm_100a CUDA kernels behind these entry points; the CLI, the
 * text reader and the output writer stay host C and call ig_mcmc_updating() where the
 * reference calls mcmc_updating().  INTEGRATION.md shows the binding.
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; every function returns an
 * ig_status (0 = ok) and never exits the process (the reference's nrerror() does,
 * nrutil.c:9-16); ig_last_error() returns the message of the last failure on the calling
 * thread.  All host arrays are row-major and caller-owned.  There is NO CPU fallback: if
 * no CUDA device is usable every call fails with IG_ERR_CUDA.
 */
#ifndef INSTRUCT_B200_H
#define INSTRUCT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ig_status {
	IG_OK = 0,
	IG_EMPTY_CLUSTER = 1,   /* chain discarded: flag_empty_cluster (mcmc.c:227-234); caller retries */
	IG_ERR_ARG = -1,
	IG_ERR_CUDA = -2,
	IG_ERR_NCCL = -3,
	IG_ERR_STATE = -4,
	IG_ERR_UNSUPPORTED = -5
} ig_status;

/* Flat replacement for the fields of SEQDATA (data_interface.h:10-56) and INIT
 * (initial.h:9-21) that mcmc_updating() reads. */
typedef struct ig_config {
	int32_t ploid;                      /* SEQDATA.ploid: 2, or 4 (mcmc_POP_tetra_selfing)      */
	int32_t popnum;                     /* SEQDATA.popnum, K                                    */
	int32_t locinum;                    /* SEQDATA.locinum, L (polymorphic loci)                */
	int32_t totalsize;                  /* SEQDATA.totalsize, N over ALL shards                 */
	int32_t mode;                       /* SEQDATA.mode: 2 population, 3 individual selfing     */
	int32_t prior_flag;                 /* SEQDATA.prior_flag: 0 uniform, 1 Dirichlet process   */
	int32_t back_refl;                  /* SEQDATA.back_refl: 1 reflective RW, 0 adaptive indep.*/
	int32_t type_freq;                  /* SEQDATA.type_freq: 1 conditional on Z, 0 expectation */
	double  alpha_dpm;                  /* SEQDATA.alpha_dpm                                    */
	int32_t nstep_check_empty_cluster;  /* SEQDATA.nstep_check_empty_cluster                    */
	int32_t print_iter;                 /* SEQDATA.print_iter                                   */
	int32_t print_freq;                 /* SEQDATA.print_freq: also accumulate P moments        */
	int32_t autopoly;                   /* SEQDATA.autopoly: 1 auto-, 0 allotetraploid (-ap)    */
	int64_t update;                     /* INIT.update: total sweeps                            */
	int64_t burnin;                     /* INIT.burnin                                          */
	int32_t thinning;                   /* INIT.thinning                                        */
	int32_t ckrep;                      /* CONVG.ckrep: retained log-lik values for GR          */
	uint64_t seed;                      /* replaces -s seed1 seed2 seed3 (random.c:50-58)       */
	int32_t device;                     /* CUDA device ordinal                                  */
	/* individual sharding of ONE chain over several GPUs: this context
	 * owns global individuals [shard_begin, shard_begin + shard_size).  A single-GPU chain has
	 * shard_begin = 0, shard_size = totalsize, shard_count = 1. */
	int32_t shard_begin;
	int32_t shard_size;
	int32_t shard_rank;
	int32_t shard_count;
	int32_t rng_rounds;                 /* Philox4x32 rounds of the bulk Z draw: 7 (default when 0; the Crush-resistant minimum of Salmon et al. SC11) or 10; every other draw uses 10 */
	int32_t use_graph;                  /* 0 (default): replay each sweep as a CUDA graph where every kernel argument is
	                                       sweep-invariant (no host DP step, -e 1, one GPU); 2: always launch directly */
	int32_t reserved[6];
} ig_config;

/* The running moments of CHAIN (mcmc.h:29-53), host-resident, caller-allocated. Sizes:
 * indvlkh, gen, gen2: N;  qq, qq2: N*K;  self_rates(2): K (mode 2, ploid 4) or N (mode 3);
 * freq, freq2: K*L*allelenum_max, only touched when print_freq = 1 (may be NULL). */
typedef struct ig_chain_result {
	int64_t steps;                      /* (update - burnin) / thinning, mcmc.c:485             */
	int64_t step;                       /* retained samples actually stored                     */
	int32_t flag_empty_cluster;
	int32_t pad;
	double totallkh, totallkh2;
	double *indvlkh;
	double *qq, *qq2;
	double *self_rates, *self_rates2;
	double *gen, *gen2;
	double *freq, *freq2;
} ig_chain_result;

typedef struct ig_ctx ig_ctx;

/* identifiers for ig_get_state / ig_set_state (test hooks for identical-state parity).
 * Shapes are the canonical host layouts, independent of the device tiling:          */
typedef enum ig_state_id {
	IG_STATE_X = 0,        /* int16  [L][Nloc][ploid]  genotype store, negative = missing (get only)  */
	IG_STATE_Z = 1,        /* int8   [L][Nloc][ploid]  UPMCMC.z                                       */
	IG_STATE_Q = 2,        /* double [N][K]            UPMCMC.qq (all shards)                         */
	IG_STATE_P = 3,        /* double [K][L][Amax]      UPMCMC.freq (device copy is fp32)              */
	IG_STATE_ALPHA = 4,    /* double [1]                                                             */
	IG_STATE_S = 5,        /* double [K] or [N]        UPMCMC.self_rates                             */
	IG_STATE_G = 6,        /* int32  [N]               UPMCMC.generation                             */
	IG_STATE_INDVLKH = 7,  /* double [N]               UPMCMC.indvlkh                                */
	IG_STATE_TOTALLKH = 8, /* double [1]                                                             */
	IG_STATE_TALLY = 9,    /* int32  [K][L][Amax]      n[k][l][a] held for the next update_P          */
	IG_STATE_CNT = 10,     /* int32  [Nloc][K]         per-individual ancestry counts (qqnum)         */
	IG_STATE_GPROP = 11,   /* int32  [N]               proposed generations of the current sweep     */
	IG_STATE_STATE = 12,   /* int32  [K]               UPMCMC.state (-e 0)                           */
	IG_STATE_MASK = 13,    /* uint8  [L][Nloc]         missindx (derived from X; get only)           */
	IG_STATE_GENO = 14,    /* int8   [L][Nloc][4]      UPMCMC.geno, tetraploid latent dosage          */
	IG_STATE_ITER = 15,    /* int64  [1]               sweep counter that keys the RNG               */
	/* autotetraploid tables (poly_geno.h:18-19), float [K][L][Gmax] natural logs, get only; setting
	 * IG_STATE_TABLES (any 4-byte payload) recomputes all three from the current P, S and S' */
	IG_STATE_TABLES = 16,      /* POLY.genofreq at the current selfing rates                        */
	IG_STATE_TABLES_PROP = 17, /* genofreq_tmp of update_S_POP at the proposed rates                */
	IG_STATE_EXFREQ = 18,      /* POLY.exfreq                                                       */
	IG_STATE_SPROP = 19,   /* double [K]               proposed selfing rates of the current sweep    */
	IG_STATE_DSTAT = 20,   /* double [K]               cal_lkd_props(k) - cal_lkd() (get only)        */
	IG_STATE_GMAX = 21,    /* int32  [1]               genotypes in the largest catalogue (get only)  */
	/* allotetraploid (-p 4 -ap 0) only: the second subgenome (copies 2,3), mcmc.h:16, poly_geno.c:441-518 */
	IG_STATE_P2 = 22,      /* double [K][L][Amax]      UPMCMC.freq2                                  */
	IG_STATE_TALLY2 = 23,  /* int32  [K][L][Amax]      tally of copies 2,3 (get only)                 */
	/* mode 3 with the Dirichlet-process prior (update_DP, DPMM.c:165-199): what gen_post_prob (DPMM.c:361-377) hands to
	 * disc_unif for individual j, with j taken out of its cluster and everybody else where they are */
	IG_STATE_DPWEIGHTS = 104,  /* double [N][N+1]  [j][0] = alpha/((G_j+1) G_j), [j][1..] = num_c dgeom(S_c, G_j) in value order (get only) */
	IG_STATE_DPCLUSTERS = 105  /* int64  [1]       number of clusters (get only); setting IG_STATE_S regroups equal values into clusters */
} ig_state_id;

/* sweep phases for ig_run_phase (test hook; ig_sweep runs them in the reference's order) */
typedef enum ig_phase {
	IG_PHASE_UPDATE_P = 1,     /* Dirichlet draw from the held tally (update_P, mcmc.c:846-857)     */
	IG_PHASE_UPDATE_S = 2,     /* update_S_POP / update_S_IND / update_DP, then propose G           */
	IG_PHASE_ZQ = 4,           /* fused update_G likelihoods + update_ZQ + tally + cal_lkh          */
	IG_PHASE_ALPHA = 8,        /* update_alpha + totallkh + empty-cluster sums                      */
	IG_PHASE_GENO = 16         /* ploid 4: update_geno + cal_lkd + tally (poly_geno.c:520,715)      */
} ig_phase;

const char *ig_version(void);
const char *ig_last_error(void);
int ig_device_count(void);

/* ---- context ---------------------------------------------------------------------- */
ig_status ig_create(const ig_config *cfg, ig_ctx **out);
void ig_destroy(ig_ctx *ctx);

/* Genotype store (a1): x is int16 [L][shard_size][ploid] in the packed locus-major /
 * individual-minor layout that the new data_interface packer emits (replaces
 * SEQDATA.seqdata + SEQDATA.missindx, data_interface.h:18,38); allelenum is int32 [L]
 * (SEQDATA.allelenum).  _device takes pointers already resident in HBM on cfg.device.
 * For ploid 4, x holds the sorted distinct-allele set per genotype padded with -1
 * (transform_data2, data_interface.c:617-640). */
ig_status ig_load_genotypes(ig_ctx *ctx, const int16_t *x_host, const int32_t *allelenum_host);
ig_status ig_load_genotypes_device(ig_ctx *ctx, const int16_t *x_dev, const int32_t *allelenum_dev);

/* ---- multi-GPU: one context per rank, NCCL communicator over the shards of one chain - */
ig_status ig_comm_unique_id(void *id128);                    /* 128-byte ncclUniqueId       */
ig_status ig_comm_init(ig_ctx *ctx, const void *id128);      /* uses cfg.shard_rank/count   */

/* ---- the drop-in call -------------------------------------------------------------- */
/* One chain on a prepared context: initialise (mcmc.c:193-206 / :315-332), run cfg.update
 * sweeps, accumulate the running moments after burn-in every cfg.thinning sweeps
 * (mcmc.c:218-226), write the first ckrep retained log-likelihoods to convg_ld
 * (= &cvg->convg_ld[chn*ckrep], mcmc.c:223-224; may be NULL), check for an empty cluster
 * (mcmc.c:227-234).  initd = INIT.initd[chn] (K floats; mode 2 / ploid 4), may be NULL
 * for mode 3.  Returns IG_OK, IG_EMPTY_CLUSTER or an error. */
ig_status ig_run_chain(ig_ctx *ctx, int32_t chain_id, const float *initd,
                       ig_chain_result *out, double *convg_ld);

/* The first ckrep retained draws of the POPULATION rates of the last ig_run_chain -- selfing rates (mode 2, ploid 4) or
 * inbreeding coefficients (mode 4) -- double [ckrep][K], same retained sweeps as convg_ld; *rows = how many are filled.
 * The reference checks convergence on the log-likelihood only (chain_converg / GelmanRubin, check_converg.c:44-153); with
 * this the caller can run the same statistic per parameter (`inbreed` does). */
ig_status ig_get_rate_trace(ig_ctx *ctx, double *out, size_t bytes, int32_t *rows);

/* mcmc_updating() in one call: create + load (host buffers, H2D inside) + run + destroy. */
ig_status ig_mcmc_updating(const ig_config *cfg, const int16_t *x_host, const int32_t *allelenum_host,
                           int32_t chain_id, const float *initd, ig_chain_result *out, double *convg_ld);

/* Device buffers freed by ig_destroy() are kept by the library and reused by the next context (one mcmc_updating() call
 * per chain would otherwise pay cudaMalloc / cudaFree of several GB every time); this returns them to the driver. */
ig_status ig_release_cache(void);

/* ---- finer-grained control (benchmarks, tests) --------------------------------------- */
ig_status ig_chain_init(ig_ctx *ctx, int32_t chain_id, const float *initd);
ig_status ig_sweep(ig_ctx *ctx, int32_t nsweeps);            /* asynchronous; see ig_sync */
ig_status ig_sync(ig_ctx *ctx);
/* nsweeps sweeps bracketed by CUDA events on the context's stream; returns elapsed ms */
ig_status ig_time_sweeps(ig_ctx *ctx, int32_t nsweeps, double *elapsed_ms);
ig_status ig_run_phase(ig_ctx *ctx, int32_t phase_mask);
ig_status ig_get_state(ig_ctx *ctx, int32_t id, void *host, size_t bytes);
ig_status ig_set_state(ig_ctx *ctx, int32_t id, const void *host, size_t bytes);

/* identical-state evaluators (parity level 2, BASELINE.json north_star):
 *  ig_loglik            log_ld_indv (mcmc.c:1726) of every local individual at gen[i]
 *  ig_proposal_loglik   proposal (mcmc.c:1630) at the population selfing rates S[K]
 *  ig_alpha_logratio    log of the update_alpha ratio (mcmc.c:1254-1260) at alpha'       */
ig_status ig_loglik(ig_ctx *ctx, const int32_t *gen, double *out);
ig_status ig_proposal_loglik(ig_ctx *ctx, const double *S, double *out);
ig_status ig_alpha_logratio(ig_ctx *ctx, double ralpha, double *out);

/* CUDA-event timing of the dominant kernel (zq_sweep) on the context's own stream:
 * enable, run sweeps, then read the number of timed launches and their total ms. */
ig_status ig_profile(ig_ctx *ctx, int32_t enable);
ig_status ig_profile_read(ig_ctx *ctx, int32_t *launches, double *zq_ms_total, int64_t *kernels_launched);
/* algorithmic bytes one zq_sweep launch moves (stated in DESIGN.md) */
ig_status ig_algorithmic_bytes(ig_ctx *ctx, double *bytes_per_sweep, double *copies_per_sweep);

#ifdef __cplusplus
}
#endif
#endif /* INSTRUCT_B200_H */
