/*
 * oracle/ref_harness_data.c -- TEST INFRASTRUCTURE ONLY.
 * Second harness translation unit: includes the reference's data_interface.c in place
 * (resolved through -I/root/reference; nothing is copied) so that its static helpers --
 * get_missing() (data_interface.c:812) and the text reader behind read_data()
 * (data_interface.c:36) -- can pin the new packer in instruct_b200/ against the
 * reference's own parsing, recoding and missing-mask rules.
 * It is a separate TU because data_interface.c and initial.h both declare word_split().
 */
#include "data_interface.c"

void refd_get_missing(SEQDATA *data) { get_missing(data); }

/* Parse a reference-format text file with the reference's own reader.  Returns the
 * (possibly corrected) N and L and the dense recoding; buffers are caller-allocated with
 * the sizes given on the command line (the reader only ever shrinks L). */
int refd_read_data(const char *path, int ploid, int N, int K, int L, const char *missing,
                   int label, int popdata, int n_extra_col, int markername_flag, int datafmt,
                   int *outN, int *outL, int *outAmax,
                   int *x /*[N][L][ploid]*/, int *allelenum /*[L]*/, int *missindx /*[N][L]*/,
                   int *alleleid /*[N][L], ploid 4 only, may be NULL*/)
{
	int i, j, k;
	SEQDATA d = read_data((char *)path, ploid, N, K, L, (char *)missing, label, popdata, 0.9,
	                      1, 1, 20, 0, 2, n_extra_col, markername_flag, 10.0, 0, 0, 0, 1, 1,
	                      datafmt, 1e9);
	*outN = d.totalsize; *outL = d.locinum; *outAmax = d.allelenum_max;
	for (j = 0; j < d.locinum; j++) allelenum[j] = d.allelenum[j];
	for (i = 0; i < d.totalsize; i++)
		for (j = 0; j < d.locinum; j++) {
			missindx[(long)i * d.locinum + j] = d.missindx[i][j];
			if (alleleid && ploid == 4) alleleid[(long)i * d.locinum + j] = d.alleleid[i][j];
			for (k = 0; k < d.ploid; k++)
				x[((long)i * d.locinum + j) * d.ploid + k] = d.seqdata[i][j][k];
		}
	return 0;
}
