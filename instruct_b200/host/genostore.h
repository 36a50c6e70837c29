/*
 * genostore.h -- host-side genotype store: text reader + recoder + packer.
 *
 * Replaces read_data() and its helpers (/root/reference data_interface.c:36-880) for the
 * diploid formats: same file formats (-af 0: one haploid copy per line; -af 1: one
 * individual per line), same column conventions (-lb label, -a popdata, -x extra columns,
 * -w marker-name line, -m missing token), same recoding rule (alleles numbered per locus
 * in order of first appearance scanning individuals then copies, data_interface.c:510-523),
 * same monomorphic-locus filter (:524-548) and same missing-genotype rule (:820-832).
 *
 * What changes (BASELINE.json north_star, "Genotype store"): the output is ONE packed
 * int16 array in locus-major / individual-minor order, x[l][i][c], negative = missing --
 * the layout ig_load_genotypes() takes -- instead of int*** tensors plus a separate
 * missindx matrix; and the reader streams the file (per-locus allele dictionaries, no
 * char[100] cell per allele copy), so a config-4 sized file needs 4 GB, not ~230 GB
 * (SURVEY.md App. B #9).
 */
#ifndef GENOSTORE_H
#define GENOSTORE_H
#include <stdint.h>

#define GS_ELMLEN 100        /* longest token, as in the reference (data_interface.c:18) */

typedef struct gs_options {
	int ploid;               /* 2, or 4 (autotetraploid: the store then holds distinct-allele sets) */
	int totalsize;           /* -N (corrected from the file like the reference does) */
	int locinum;             /* -L (corrected from the file) */
	const char *missing;     /* -m, default "-9" */
	int label;               /* -lb */
	int popdata;             /* -a  */
	int n_extra_col;         /* -x  */
	int markername_flag;     /* -w  */
	int datafmt;             /* -af */
	int quiet;               /* do not echo the transformed matrix to stdout */
} gs_options;

typedef struct gs_store {
	int ploid, totalsize, locinum, locinum_file, allelenum_max;
	int16_t *x;              /* [locinum][totalsize][ploid], -9 = missing; ploid 4: ascending distinct alleles, -1 padding */
	int32_t *allelenum;      /* [locinum]                                              */
	char ***alleletype;      /* [locinum][allelenum[l]] original allele strings        */
	int *locus_of;           /* [locinum] index of the locus in the file (0-based)     */
	char **marker_names;     /* [locinum_file] or NULL                                 */
	char **indvname;         /* [totalsize] or NULL                                    */
	int *popindx;            /* [totalsize] or NULL                                    */
	char **poptype;          /* [pop_count]                                            */
	int pop_count;
	char ***extra_col;       /* [totalsize][n_extra_col] or NULL                       */
	int n_extra_col;
	int *missvec;            /* [totalsize] number of loci with a missing genotype     */
} gs_store;

/* returns 0 on success; on failure returns non-zero and writes a message to err */
int gs_read(const char *path, const gs_options *opt, gs_store *out, char *err, int errlen);
void gs_free(gs_store *s);

/* Packed on-disk form of a store: the int16 matrix exactly as
 * ig_load_genotypes() takes it, plus every table the result writer needs (allele strings, marker and
 * individual names, population labels, extra columns), so that a second run on the same data skips the
 * text reader altogether -- at config-4 size the text file is ~6 GB of tokens, the packed store 4 GB read
 * with one fread.  Little-endian, versioned by the magic; gs_load() validates every length it reads. */
int gs_save(const char *path, const gs_store *s, char *err, int errlen);
int gs_load(const char *path, gs_store *out, char *err, int errlen);

#endif
