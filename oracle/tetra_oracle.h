/*
 * oracle/tetra_oracle.h -- CPU restatement of the reference's AUTOTETRAPLOID sweep
 * (mcmc_POP_tetra_selfing, poly_geno.c:75-140, `-p 4 -ap 1`).
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's CPU legs as
 * the checker / reported baseline, never by the product path.
 *
 * Pinned (tests/test_tetra_oracle_vs_reference.py) against the unmodified reference compiled in
 * oracle/_ref/ (harness oracle/ref_harness_poly.c): genotype lists, the log genotype-frequency
 * tables (float) and whole-chain running moments agree BIT FOR BIT on identical Wichmann-Hill
 * seeds.  The reference ships no golden vectors for this path.
 *
 * Layout (this repo's, not the reference's pointer tensors):
 *   x     int16 [L][N][4]  the DISTINCT alleles observed, ascending, -1 padding (data_interface.c:571-669)
 *   nd    uint8 [L][N]     number of distinct alleles ("alleleid"); 0 = missing (data_interface.c:722-741)
 *   geno  int8  [L][N][4]  latent dosage resolution;  z int8 [L][N][4] ancestry of each copy
 *   freq  double [K][L][Amax];  tables float [K][L][Gmax]
 */
#ifndef TETRA_ORACLE_H
#define TETRA_ORACLE_H
#include <stdint.h>
#include "instruct_oracle.h"

typedef struct tet_model tet_model;

tet_model *tet_new(int N, int L, int K, int back_refl, const int16_t *x, const uint8_t *nd, const int32_t *allelenum);
/* autopoly = 0: the ALLOTETRAPLOID model (-ap 0): two subgenomes (copies 0,1 / copies 2,3) with their own
 * allele frequencies freq / freq2; update_P_allo :441, calc_exfreq_allo :1592, allo_genfreq :2122,
 * choose_two/tri/tetra_allo :962-1215.  Pinned bit for bit like the autotetraploid path. */
tet_model *tet_new2(int N, int L, int K, int back_refl, int autopoly, const int16_t *x, const uint8_t *nd, const int32_t *allelenum);
double *tet_freq2(tet_model *m);
void tet_tally_allo(const tet_model *m, int32_t *n1, int32_t *n2);
int tet_geno_conditional_allo(const tet_model *m, int i, int l, double *prob /*[12]*/);
void tet_free(tet_model *m);
void tet_setseeds(tet_model *m, long a, long b, long c);

int8_t *tet_z(tet_model *m);
int8_t *tet_geno(tet_model *m);
double *tet_qq(tet_model *m);
double *tet_qqnum(tet_model *m);
double *tet_freq(tet_model *m);
double *tet_self(tet_model *m);
int *tet_state(tet_model *m);
double *tet_alpha(tet_model *m);
double *tet_indvlkh(tet_model *m);
double *tet_totallkh(tet_model *m);
float *tet_exfreq(tet_model *m);
float *tet_genofreq(tet_model *m);
int tet_gmax(const tet_model *m);
int tet_amax(const tet_model *m);
int tet_genolist(const tet_model *m, int l, int *codes);
int tet_geno_index(const tet_model *m, int l, const int8_t *g4);     /* get_index_auto, poly_geno.c:1289 */

/* pure functions */
void tet_tally(const tet_model *m, int32_t *n /*[K][L][Amax]*/);
void tet_count_z(const tet_model *m, double *cnt /*[N][K]*/);
void tet_calc_exfreq(tet_model *m);                                   /* calc_exfreq_auto, :1515 */
void tet_calc_genofreq(tet_model *m, int k, double self, float *out /*[L][Gmax]*/);   /* calc_self_genofreq, :1219 */
double tet_site_loglik(const tet_model *m, int l, int i, const int8_t *z4);           /* calc_genofq, :1235 */
double tet_cal_lkd(tet_model *m);                                     /* :715 */
double tet_cal_lkd_props(const tet_model *m, int k, const float *tab /*[L][Gmax]*/);  /* :645 */
void tet_geno_conditional(const tet_model *m, int i, int l, double *prob /*[3]*/);     /* choose_two/tri_auto, :854,:907 */
void tet_z_conditional(const tet_model *m, int i, int l, int c, double *prob /*[K]*/);

/* conditional updates, consuming the stream exactly like the reference */
void tet_initial_geno(tet_model *m);       /* :316 */
void tet_update_P(tet_model *m);           /* update_P_auto :390 */
void tet_update_S(tet_model *m);           /* update_S_POP :584 */
void tet_update_ZQ(tet_model *m, int init_flag);   /* :750 */
void tet_update_geno(tet_model *m);        /* :520 */
void tet_sweeps(tet_model *m, int n);

/* whole chain = mcmc_POP_tetra_selfing; out uses orc_chain (gen/gen2 unused) */
orc_chain *tet_chain_new(const tet_model *m, int ckrep);
int tet_run_chain(tet_model *m, long update, long burnin, int thinning, int ckrep, int nstep_check_empty,
                  const float *initd, orc_chain *out);
#endif
