/*
 * genostore.c -- see genostore.h.  Written from scratch; follows the SEMANTICS of the
 * reference's reader (data_interface.c), not its structure: one streaming pass with a small
 * allele dictionary per locus instead of a char[100] cell per allele copy.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "genostore.h"

#define GS_MISSING (-9)      /* data_interface.c:494 */

typedef struct { char **name; int n, cap; } dict_t;

static int fail(char *err, int errlen, const char *msg)
{
	if (err && errlen > 0) snprintf(err, errlen, "%s", msg);
	return 1;
}

static char *dup_str(const char *s)
{
	size_t n = strlen(s) + 1;
	char *d = (char *)malloc(n);
	if (d) memcpy(d, s, n);
	return d;
}

/* split a line in place into whitespace-separated tokens (word_cnt/word_split,
 * data_interface.c:765-805); returns the token count */
static int tokenize(char *line, char ***tok, int *cap)
{
	int n = 0;
	char *p = line;
	for (;;) {
		while (*p && isspace((unsigned char)*p)) p++;
		if (!*p) break;
		if (n == *cap) {
			*cap = *cap ? *cap * 2 : 1024;
			*tok = (char **)realloc(*tok, (size_t)*cap * sizeof(char *));
		}
		(*tok)[n++] = p;
		while (*p && !isspace((unsigned char)*p)) p++;
		if (*p) *p++ = '\0';
	}
	return n;
}

static int dict_id(dict_t *d, const char *s)
{
	int i;
	for (i = 0; i < d->n; i++) if (strcmp(d->name[i], s) == 0) return i;
	if (d->n == d->cap) {
		d->cap = d->cap ? d->cap * 2 : 4;
		d->name = (char **)realloc(d->name, (size_t)d->cap * sizeof(char *));
	}
	d->name[d->n] = dup_str(s);
	return d->n++;
}

static int pop_id(gs_store *s, int *cap, const char *name)
{
	int i;
	for (i = 0; i < s->pop_count; i++) if (strcmp(s->poptype[i], name) == 0) return i;
	if (s->pop_count == *cap) {
		*cap = *cap ? *cap + 5 : 5;                    /* INCRE_POP, data_interface.c:19 */
		s->poptype = (char **)realloc(s->poptype, (size_t)*cap * sizeof(char *));
	}
	s->poptype[s->pop_count] = dup_str(name);
	return s->pop_count++;
}

int gs_read(const char *path, const gs_options *opt, gs_store *out, char *err, int errlen)
{
	FILE *fp;
	char *line = NULL, **tok = NULL, msg[256];
	size_t linecap = 0;
	ssize_t got;
	int tokcap = 0, ntok, lines = 0, first = 1, meta, cnt_token = 0, Lf = 0, N, ploid = opt->ploid;
	int i, l, c, popcap = 0, kept = 0;
	dict_t *dict = NULL;
	int16_t *raw = NULL;
	long row = 0;

	memset(out, 0, sizeof(*out));
	int datafmt = opt->datafmt;
	if (ploid != 2 && ploid != 4) return fail(err, errlen, "genostore: ploid must be 2 or 4");
	if (ploid == 4) datafmt = 1;                        /* read_data: ploid 4 is always the one-line format (data_interface.c:78-82) */
	if ((fp = fopen(path, "r")) == NULL) return fail(err, errlen, "Cannot open input file!");
	meta = opt->label + opt->popdata + opt->n_extra_col;

	/* ---- pass 1: number of loci from the first non-empty line (cnt_loci, :356-421) and the
	 *      number of data lines (cnt_lines, :427-487) */
	while ((got = getline(&line, &linecap, fp)) >= 0) {
		ntok = tokenize(line, &tok, &tokcap);
		if (ntok == 0) continue;
		if (first) {
			first = 0;
			if (opt->markername_flag) {
				Lf = ntok;
				out->marker_names = (char **)calloc((size_t)Lf, sizeof(char *));
				for (l = 0; l < Lf; l++) out->marker_names[l] = dup_str(tok[l]);
				fprintf(stdout, "The number of loci is %d now!\n", Lf);
				continue;
			}
			Lf = (datafmt == 0) ? ntok - meta : (ntok - meta) / ploid;
			if (Lf != opt->locinum)
				fprintf(stdout, "The Input Number of Loci is wrong!\nThe number of loci is %d now!\n", Lf);
		}
		lines++;
	}
	if (Lf < 1) { fclose(fp); free(line); free(tok); return fail(err, errlen, "no loci found in the input file"); }
	if (datafmt == 0) {
		if (lines % ploid != 0) { fclose(fp); free(line); free(tok); return fail(err, errlen, "Some individuals do not have two copies of haplotype!"); }
		N = lines / ploid;
	} else N = lines;
	if (N != opt->totalsize) fprintf(stdout, "The input population size is incorrect!\nThe population size is %d\n", N);
	cnt_token = (datafmt == 0) ? meta + Lf : meta + Lf * ploid;

	out->ploid = ploid; out->totalsize = N; out->locinum_file = Lf; out->n_extra_col = opt->n_extra_col;
	raw = (int16_t *)malloc((size_t)Lf * N * ploid * sizeof(int16_t));
	dict = (dict_t *)calloc((size_t)Lf, sizeof(dict_t));
	if (opt->label) out->indvname = (char **)calloc((size_t)N, sizeof(char *));
	if (opt->popdata) out->popindx = (int *)calloc((size_t)N, sizeof(int));
	if (opt->n_extra_col > 0) out->extra_col = (char ***)calloc((size_t)N, sizeof(char **));
	if (!raw || !dict) { fclose(fp); return fail(err, errlen, "out of memory reading the genotype store"); }

	/* ---- pass 2: parse, recode in file order (= individuals, then copies: the order
	 *      transform_data scans in, :510-523) */
	rewind(fp);
	first = 1;
	while ((got = getline(&line, &linecap, fp)) >= 0) {
		ntok = tokenize(line, &tok, &tokcap);
		if (ntok == 0) continue;
		if (first && opt->markername_flag) { first = 0; continue; }
		first = 0;
		if (ntok != cnt_token) {
			fclose(fp);
			snprintf(msg, sizeof msg, "The number of tokens in one line does not match the parameters input from the commandline (line %ld has %d, expected %d)", row + 1, ntok, cnt_token);
			return fail(err, errlen, msg);
		}
		i = (datafmt == 0) ? (int)(row / ploid) : (int)row;
		c = (datafmt == 0) ? (int)(row % ploid) : 0;
		if (c == 0) {
			if (opt->label) out->indvname[i] = dup_str(tok[opt->label - 1]);
			if (opt->popdata) out->popindx[i] = pop_id(out, &popcap, tok[opt->label + opt->popdata - 1]);
			if (opt->n_extra_col > 0) {
				out->extra_col[i] = (char **)calloc((size_t)opt->n_extra_col, sizeof(char *));
				for (l = 0; l < opt->n_extra_col; l++) out->extra_col[i][l] = dup_str(tok[opt->label + opt->popdata + l]);
			}
		} else if (opt->label && strcmp(out->indvname[i], tok[opt->label - 1]) != 0) {
			fclose(fp);
			return fail(err, errlen, "Some individuals have different number of haplotypes!");
		}
		if (datafmt == 0) {
			for (l = 0; l < Lf; l++) {
				const char *t = tok[meta + l];
				raw[((size_t)l * N + i) * ploid + c] = strcmp(t, opt->missing) == 0 ? GS_MISSING : (int16_t)dict_id(&dict[l], t);
			}
		} else {
			for (l = 0; l < Lf; l++)
				for (c = 0; c < ploid; c++) {
					const char *t = tok[meta + l * ploid + c];
					raw[((size_t)l * N + i) * ploid + c] = strcmp(t, opt->missing) == 0 ? GS_MISSING : (int16_t)dict_id(&dict[l], t);
				}
		}
		row++;
	}
	fclose(fp);
	free(line);
	free(tok);

	if (ploid == 4) {
		/* transform_data2 (data_interface.c:571-669): every locus is kept; a genotype becomes the
		 * ascending set of its distinct alleles padded with -1; "alleleid" is the set's size and
		 * 0 of them means missing (get_missing_tetra :722-741) */
		out->locinum = Lf;
		out->x = raw;
		out->allelenum = (int32_t *)malloc((size_t)Lf * sizeof(int32_t));
		out->alleletype = (char ***)calloc((size_t)Lf, sizeof(char **));
		out->locus_of = (int *)malloc((size_t)Lf * sizeof(int));
		out->missvec = (int *)calloc((size_t)N, sizeof(int));
		for (l = 0; l < Lf; l++) {
			if (dict[l].n < 1) { snprintf(msg, sizeof msg, "locus %d has no observed allele", l + 1); return fail(err, errlen, msg); }
			out->allelenum[l] = dict[l].n;
			out->alleletype[l] = dict[l].name;
			out->locus_of[l] = l;
			if (dict[l].n > out->allelenum_max) out->allelenum_max = dict[l].n;
			for (i = 0; i < N; i++) {
				int16_t *g4 = raw + ((size_t)l * N + i) * 4, set[4];
				int ns = 0, a, b;
				for (c = 0; c < 4; c++) {
					int seen = 0;
					if (g4[c] < 0) continue;
					for (a = 0; a < ns; a++) if (set[a] == g4[c]) seen = 1;
					if (!seen) set[ns++] = g4[c];
				}
				for (a = 0; a < ns; a++) for (b = a + 1; b < ns; b++) if (set[a] > set[b]) { int16_t t = set[a]; set[a] = set[b]; set[b] = t; }
				for (c = 0; c < 4; c++) g4[c] = c < ns ? set[c] : (int16_t)-1;
				if (ns == 0) out->missvec[i]++;
			}
		}
		free(dict);
		if (!opt->quiet) {                             /* the reference's echo, :652-676 */
			fprintf(stdout, "Print the number of alleles per individual per locus:\n");
			for (i = 0; i < N; i++) {
				for (l = 0; l < Lf; l++) { int ns = 0; for (c = 0; c < 4; c++) ns += out->x[((size_t)l * N + i) * 4 + c] >= 0; fprintf(stdout, "%d ", ns); }
				fprintf(stdout, "\n");
			}
			fprintf(stdout, "Print the transformed allele data:\n");
			for (i = 0; i < N; i++)
				for (c = 0; c < 4; c++) {
					for (l = 0; l < Lf; l++) {
						int v = out->x[((size_t)l * N + i) * 4 + c];
						if (c == 0 && v < 0) v = GS_MISSING;   /* seqdata[0] = -9 on a missing genotype (:648-649) */
						fprintf(stdout, "%d ", v);
					}
					fprintf(stdout, "\n");
				}
			fprintf(stdout, "End the printing of the transformed allele data.\n");
		}
		return 0;
	}

	/* ---- keep the polymorphic loci (:524-548) and compact */
	for (l = 0; l < Lf; l++) {
		if (dict[l].n >= 2) kept++;
		else fprintf(stdout, "The locus %d is not polymorphic.\n", l + 1);
	}
	fprintf(stdout, "The number of polymorphic loci is %d now.\n", kept);
	if (kept == 0) { free(raw); return fail(err, errlen, "no polymorphic locus in the input file"); }
	out->locinum = kept;
	out->x = (int16_t *)malloc((size_t)kept * N * ploid * sizeof(int16_t));
	out->allelenum = (int32_t *)malloc((size_t)kept * sizeof(int32_t));
	out->alleletype = (char ***)calloc((size_t)kept, sizeof(char **));
	out->locus_of = (int *)malloc((size_t)kept * sizeof(int));
	out->missvec = (int *)calloc((size_t)N, sizeof(int));
	kept = 0;
	for (l = 0; l < Lf; l++) {
		if (dict[l].n >= 2) {
			memcpy(out->x + (size_t)kept * N * ploid, raw + (size_t)l * N * ploid, (size_t)N * ploid * sizeof(int16_t));
			out->allelenum[kept] = dict[l].n;
			out->alleletype[kept] = dict[l].name;
			out->locus_of[kept] = l;
			if (dict[l].n > out->allelenum_max) out->allelenum_max = dict[l].n;
			kept++;
		} else {
			for (i = 0; i < dict[l].n; i++) free(dict[l].name[i]);
			free(dict[l].name);
		}
	}
	free(raw);
	free(dict);
	/* missvec: loci at which ANY copy is missing (get_missing, :812-835) */
	for (l = 0; l < out->locinum; l++)
		for (i = 0; i < N; i++) {
			int miss = 0;
			for (c = 0; c < ploid; c++) if (out->x[((size_t)l * N + i) * ploid + c] < 0) miss = 1;
			out->missvec[i] += miss;
		}
	if (!opt->quiet) {                                 /* the reference's echo, :554-566 */
		fprintf(stdout, "Print the transformed allele data:\n");
		for (i = 0; i < N; i++)
			for (c = 0; c < ploid; c++) {
				for (l = 0; l < out->locinum; l++) fprintf(stdout, "%d ", out->x[((size_t)l * N + i) * ploid + c]);
				fprintf(stdout, "\n");
			}
		fprintf(stdout, "End the printing of the transformed allele data.\n");
	}
	return 0;
}

/* ------------------------------------------------------------------ packed on-disk form */
#define GS_MAGIC "IGSTORE1"

static int wr_i32(FILE *f, int32_t v) { return fwrite(&v, 4, 1, f) == 1 ? 0 : 1; }
static int wr_str(FILE *f, const char *s)
{
	const int32_t n = s ? (int32_t)strlen(s) : -1;
	if (wr_i32(f, n)) return 1;
	return (n > 0 && fwrite(s, 1, (size_t)n, f) != (size_t)n) ? 1 : 0;
}
static int rd_i32(FILE *f, int32_t *v) { return fread(v, 4, 1, f) == 1 ? 0 : 1; }
static int rd_str(FILE *f, char **out)
{
	int32_t n;
	*out = NULL;
	if (rd_i32(f, &n)) return 1;
	if (n < 0) return 0;
	if (n > 1 << 20) return 1;
	*out = (char *)malloc((size_t)n + 1);
	if (!*out) return 1;
	if (n > 0 && fread(*out, 1, (size_t)n, f) != (size_t)n) return 1;
	(*out)[n] = '\0';
	return 0;
}

int gs_save(const char *path, const gs_store *s, char *err, int errlen)
{
	FILE *f = fopen(path, "wb");
	int bad = 0, i, l;
	const size_t nx = (size_t)s->locinum * s->totalsize * s->ploid;
	if (!f) return fail(err, errlen, "genostore: cannot create the packed store");
	bad |= fwrite(GS_MAGIC, 1, 8, f) != 8;
	bad |= wr_i32(f, s->ploid) | wr_i32(f, s->totalsize) | wr_i32(f, s->locinum) | wr_i32(f, s->locinum_file) | wr_i32(f, s->allelenum_max);
	bad |= wr_i32(f, s->pop_count) | wr_i32(f, s->n_extra_col);
	bad |= wr_i32(f, s->marker_names != NULL) | wr_i32(f, s->indvname != NULL) | wr_i32(f, s->popindx != NULL) | wr_i32(f, s->extra_col != NULL);
	bad |= fwrite(s->allelenum, 4, (size_t)s->locinum, f) != (size_t)s->locinum;
	bad |= fwrite(s->locus_of, 4, (size_t)s->locinum, f) != (size_t)s->locinum;
	bad |= fwrite(s->missvec, 4, (size_t)s->totalsize, f) != (size_t)s->totalsize;
	bad |= fwrite(s->x, 2, nx, f) != nx;
	for (l = 0; l < s->locinum && !bad; l++)
		for (i = 0; i < s->allelenum[l]; i++) bad |= wr_str(f, s->alleletype[l][i]);
	if (s->marker_names) for (l = 0; l < s->locinum_file && !bad; l++) bad |= wr_str(f, s->marker_names[l]);
	if (s->indvname) for (i = 0; i < s->totalsize && !bad; i++) bad |= wr_str(f, s->indvname[i]);
	if (s->popindx) bad |= fwrite(s->popindx, 4, (size_t)s->totalsize, f) != (size_t)s->totalsize;
	for (i = 0; i < s->pop_count && !bad; i++) bad |= wr_str(f, s->poptype[i]);
	if (s->extra_col)
		for (i = 0; i < s->totalsize && !bad; i++)
			for (l = 0; l < s->n_extra_col; l++) bad |= wr_str(f, s->extra_col[i] ? s->extra_col[i][l] : NULL);
	bad |= wr_i32(f, 0x45444e45);                         /* "ENDE": a truncated file is detected on load */
	if (fclose(f) != 0) bad = 1;
	return bad ? fail(err, errlen, "genostore: write error on the packed store") : 0;
}

int gs_load(const char *path, gs_store *out, char *err, int errlen)
{
	FILE *f = fopen(path, "rb");
	char magic[8];
	int32_t h[11], tail = 0;
	int bad = 0, i, l;
	size_t nx;
	memset(out, 0, sizeof(*out));
	if (!f) return fail(err, errlen, "genostore: cannot open the packed store");
	if (fread(magic, 1, 8, f) != 8 || memcmp(magic, GS_MAGIC, 8) != 0) { fclose(f); return fail(err, errlen, "genostore: not a packed store (bad magic)"); }
	for (i = 0; i < 11; i++) bad |= rd_i32(f, &h[i]);
	if (bad || (h[0] != 2 && h[0] != 4) || h[1] < 1 || h[2] < 0 || h[3] < h[2] || h[4] < 0 || h[4] > 32767 || h[5] < 0 || h[6] < 0) {
		fclose(f);
		return fail(err, errlen, "genostore: corrupt header in the packed store");
	}
	out->ploid = h[0]; out->totalsize = h[1]; out->locinum = h[2]; out->locinum_file = h[3]; out->allelenum_max = h[4];
	out->pop_count = h[5]; out->n_extra_col = h[6];
	nx = (size_t)out->locinum * out->totalsize * out->ploid;
	out->allelenum = (int32_t *)malloc(sizeof(int32_t) * (size_t)(out->locinum ? out->locinum : 1));
	out->locus_of = (int *)malloc(sizeof(int) * (size_t)(out->locinum ? out->locinum : 1));
	out->missvec = (int *)malloc(sizeof(int) * (size_t)out->totalsize);
	out->x = (int16_t *)malloc(sizeof(int16_t) * (nx ? nx : 1));
	if (!out->allelenum || !out->locus_of || !out->missvec || !out->x) bad = 1;
	if (!bad) {
		bad |= fread(out->allelenum, 4, (size_t)out->locinum, f) != (size_t)out->locinum;
		bad |= fread(out->locus_of, 4, (size_t)out->locinum, f) != (size_t)out->locinum;
		bad |= fread(out->missvec, 4, (size_t)out->totalsize, f) != (size_t)out->totalsize;
		bad |= fread(out->x, 2, nx, f) != nx;
	}
	for (l = 0; l < out->locinum && !bad; l++) if (out->allelenum[l] < 0 || out->allelenum[l] > out->allelenum_max) bad = 1;
	if (!bad) {
		out->alleletype = (char ***)calloc((size_t)(out->locinum ? out->locinum : 1), sizeof(char **));
		for (l = 0; l < out->locinum && !bad; l++) {
			out->alleletype[l] = (char **)calloc((size_t)(out->allelenum[l] ? out->allelenum[l] : 1), sizeof(char *));
			for (i = 0; i < out->allelenum[l] && !bad; i++) bad |= rd_str(f, &out->alleletype[l][i]);
		}
	}
	if (!bad && h[7]) {
		out->marker_names = (char **)calloc((size_t)out->locinum_file, sizeof(char *));
		for (l = 0; l < out->locinum_file && !bad; l++) bad |= rd_str(f, &out->marker_names[l]);
	}
	if (!bad && h[8]) {
		out->indvname = (char **)calloc((size_t)out->totalsize, sizeof(char *));
		for (i = 0; i < out->totalsize && !bad; i++) bad |= rd_str(f, &out->indvname[i]);
	}
	if (!bad && h[9]) {
		out->popindx = (int *)malloc(sizeof(int) * (size_t)out->totalsize);
		bad |= fread(out->popindx, 4, (size_t)out->totalsize, f) != (size_t)out->totalsize;
	}
	if (!bad) {
		out->poptype = (char **)calloc((size_t)(out->pop_count ? out->pop_count : 1), sizeof(char *));
		for (i = 0; i < out->pop_count && !bad; i++) bad |= rd_str(f, &out->poptype[i]);
	}
	if (!bad && h[10]) {
		out->extra_col = (char ***)calloc((size_t)out->totalsize, sizeof(char **));
		for (i = 0; i < out->totalsize && !bad; i++) {
			out->extra_col[i] = (char **)calloc((size_t)(out->n_extra_col ? out->n_extra_col : 1), sizeof(char *));
			for (l = 0; l < out->n_extra_col && !bad; l++) bad |= rd_str(f, &out->extra_col[i][l]);
		}
	}
	if (!bad) bad |= rd_i32(f, &tail) || tail != 0x45444e45;
	fclose(f);
	if (bad) { gs_free(out); return fail(err, errlen, "genostore: truncated or corrupt packed store"); }
	return 0;
}

void gs_free(gs_store *s)
{
	int i, l;
	if (!s) return;
	free(s->x);
	if (s->alleletype)
		for (l = 0; l < s->locinum; l++) {
			if (!s->alleletype[l]) continue;                 /* a load that stopped half way */
			for (i = 0; i < s->allelenum[l]; i++) free(s->alleletype[l][i]);
			free(s->alleletype[l]);
		}
	free(s->alleletype); free(s->allelenum); free(s->locus_of); free(s->missvec);
	if (s->marker_names) { for (l = 0; l < s->locinum_file; l++) free(s->marker_names[l]); free(s->marker_names); }   /* free(NULL) is fine */
	if (s->indvname) { for (i = 0; i < s->totalsize; i++) free(s->indvname[i]); free(s->indvname); }
	if (s->extra_col) {
		for (i = 0; i < s->totalsize; i++) if (s->extra_col[i]) { for (l = 0; l < s->n_extra_col; l++) free(s->extra_col[i][l]); free(s->extra_col[i]); }
		free(s->extra_col);
	}
	if (s->poptype) for (i = 0; i < s->pop_count; i++) free(s->poptype[i]);
	free(s->poptype); free(s->popindx);
	memset(s, 0, sizeof(*s));
}
