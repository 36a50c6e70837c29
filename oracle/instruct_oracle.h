/*
 * oracle/instruct_oracle.h -- CPU restatement of the reference's diploid hot path.
 *
 * TEST INFRASTRUCTURE ONLY: may be used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs, always as the checker or the reported
 * baseline, never as the product path.
 *
 * Pinned (tests/test_oracle_vs_reference.py) against the unmodified reference compiled
 * in oracle/_ref/ (see oracle/Makefile): integer tallies and whole-chain running moments
 * bit-exact on identical Wichmann-Hill seeds, log-likelihoods to 1e-12.  The reference
 * ships no golden vectors of its own, so the committed fixtures in
 * tests/golden/ were generated from that compiled reference by tools/make_golden.py.
 *
 * Layout: the NEW packed layout, not the reference's pointer tensors --
 *   x  int16 [L][N][ploid]   allele index 0..A_l-1, negative = missing
 *   z  int8  [L][N][ploid]
 *   qq, qqnum double [N][K];  freq double [K][L][Amax]
 * Loop ORDER (and therefore RNG consumption order) follows the reference, so a chain run
 * here from the same seeds reproduces the reference's chain.
 */
#ifndef INSTRUCT_ORACLE_H
#define INSTRUCT_ORACLE_H
#include <stdint.h>

typedef struct { long s1, s2, s3; } orc_rng;     /* random.c:10-16 */

typedef struct orc_model orc_model;

orc_model *orc_new(int N, int L, int K, int ploid, int mode, int prior_flag, int back_refl,
                   int type_freq, double alpha_dpm, const int16_t *x, const int32_t *allelenum);
void orc_free(orc_model *m);

/* raw pointers into the model state, for ctypes/numpy views */
int8_t *orc_z(orc_model *m);
int *orc_zz(orc_model *m);                 /* mode 0: cluster of each individual */
double *orc_qq(orc_model *m);
double *orc_qqnum(orc_model *m);
double *orc_freq(orc_model *m);
double *orc_self(orc_model *m);
int *orc_state(orc_model *m);
int *orc_gen(orc_model *m);
double *orc_indvlkh(orc_model *m);
double *orc_alpha(orc_model *m);
double *orc_totallkh(orc_model *m);
int orc_amax(orc_model *m);
void orc_setseeds(orc_model *m, long a, long b, long c);
void orc_getseeds(orc_model *m, long *out3);

/* samplers (random.c) on the model's stream */
double orc_ran1(orc_model *m);
double orc_rgamma(orc_model *m, double a, double b);
double orc_rbeta(orc_model *m, double a, double b);
double orc_rnormal(orc_model *m, double mu, double sd);
int orc_rgeom(orc_model *m, double p);
int orc_disc_unif(orc_model *m, double *vec, int len);

/* pure functions on the current state (no RNG) */
void orc_missing_mask(const orc_model *m, uint8_t *mask /*[L][N]*/);
void orc_tally(const orc_model *m, int32_t *n /*[K][L][Amax]*/);
void orc_tally_range(const orc_model *m, int i0, int i1, int32_t *n);
void orc_count_z(const orc_model *m, double *cnt /*[N][K]*/);
double orc_genofreq(int a0, int a1, double f0, double f1, int gen);
double orc_log_ld_indv(const orc_model *m, int gen, int i);
double orc_log_ld_noselfing(const orc_model *m, int i);
double orc_proposal(const orc_model *m, const double *S);
double orc_dgeom(double s, int g);
double orc_genofreq_F(int a0, int a1, double f0, double f1, double F);      /* genofreq_inbreedcoff, mcmc.c:1707 */
double orc_log_ld_F(const orc_model *m, const double *inbreed, int by_pop, int i); /* log_ld_F_pop / _indv, mcmc.c:1776,1812 */
double orc_log_ld_F_total(const orc_model *m, const double *inbreed);       /* mcmc.c:1849 */
int orc_dt_stat(double s);
double orc_alpha_logratio(const orc_model *m, double ralpha);      /* log form of mcmc.c:1254-1260 */
double orc_alpha_ratio_product(const orc_model *m, double ralpha); /* the reference's product form */
int orc_check_empty_cluster(const orc_model *m);
void orc_z_conditional(const orc_model *m, int i, int l, int c, double *prob /*[K]*/);

/* conditional updates (consume the model's RNG exactly like the reference) */
void orc_update_P(orc_model *m);
void orc_update_S_POP(orc_model *m);
void orc_update_S_IND(orc_model *m);
void orc_update_F_POP(orc_model *m);   /* update_inbreedcoff_POP, mcmc.c:986 (mode 4; self_rates holds inbreed) */
void orc_update_F_IND(orc_model *m);   /* update_F_IND, mcmc.c:888 (mode 5, uniform prior) */
void orc_update_G(orc_model *m);
void orc_update_ZQ(orc_model *m, int init_flag);
void orc_update_Z(orc_model *m, int init_flag);      /* mode 0, mcmc.c:1094 */
double orc_log_ld_indv_K(const orc_model *m, int i, int k);   /* mcmc.c:1893 */
void orc_update_alpha(orc_model *m);
void orc_cal_lkh(orc_model *m);
void orc_init_DP(orc_model *m);
void orc_update_DP(orc_model *m);
int orc_dp_nclusters(const orc_model *m);
void orc_dp_from_values(orc_model *m);   /* test hook: clusters = groups of equal self_rates[] */
void orc_sweeps(orc_model *m, int n);

/* whole chain = the reference's mcmc_POP_selfing (mode 2) / mcmc_INDV_selfing (mode 3) */
typedef struct {
	long steps, step;
	int flag_empty_cluster;
	double totallkh, totallkh2;
	double *indvlkh;             /* [N]            */
	double *qq, *qq2;            /* [N][K]         */
	double *self_rates, *self_rates2; /* [K] or [N] */
	double *gen, *gen2;          /* [N]            */
	double *convg;               /* [ckrep]        */
} orc_chain;

orc_chain *orc_chain_new(const orc_model *m, int ckrep);
void orc_chain_free(orc_chain *c);
int orc_run_chain(orc_model *m, long update, long burnin, int thinning, int ckrep,
                  int nstep_check_empty, const float *initd, orc_chain *out);
void orc_store_chn(const orc_model *m, orc_chain *c);

double orc_gelman_rubin_ref(const double *vec, int numchains, int totrep);  /* check_converg.c:100, as written */
double orc_gelman_rubin(const double *vec /*[chains][n]*/, int numchains, int n); /* across-chain, corrected */

#endif
