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
