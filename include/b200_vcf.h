/*  b200_vcf.h -- text VCF header / record model without htslib, host C.
 *
 *  The `call` driver of the reference reads records with bcf_sr_next_line / bcf_unpack (vcfcall.c:471-499), fetches
 *  FORMAT/PL, INFO/QS, ... with bcf_get_format_int32 / bcf_get_info_float (mcall.c:1444-1510), edits the record with
 *  bcf_update_info / bcf_update_format / bcf_update_alleles / bcf_update_genotypes (mcall.c:1583-1681) and writes it
 *  with bcf_write1 (vcfcall.c:1147).  htslib is not in the reference tree, so this file restates the part of that
 *  behaviour the path needs for TEXT VCF, following the published VCFv4.2 rules and htslib's documented conventions:
 *
 *    - a FORMAT vector is as long as the longest sample's, shorter samples are padded with vector_end, "." is
 *      missing (what bcf_get_format_int32 returns);
 *    - an updated tag keeps its position, a new tag goes to the end, GT always comes first, n = 0 removes;
 *    - integers print as-is, floats with six significant digits and trailing zeros culled (kputd), a vector stops at
 *      its first vector_end and an empty one prints ".";
 *    - fields the caller never touches keep their input text.
 *
 *  tests/test_vcf_text.py pins it: every VCF of the reference's `call` tests, inputs and expected outputs, parses
 *  and re-formats to the same bytes.  Plain C, no exit(): functions return NULL / negative on malformed input.
 */
#ifndef B200_VCF_H
#define B200_VCF_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_I32_MISSING     INT32_MIN
#define B200_I32_VECTOR_END  (INT32_MIN+1)
#define B200_F32_MISSING_BITS     0x7F800001u
#define B200_F32_VECTOR_END_BITS  0x7F800002u

#define B200_VL_FIXED 0     /* Number=<n> */
#define B200_VL_VAR   1     /* Number=.   */
#define B200_VL_A     2
#define B200_VL_G     3
#define B200_VL_R     4

#define B200_HT_FLAG 0
#define B200_HT_INT  1
#define B200_HT_REAL 2
#define B200_HT_STR  3

typedef struct { char *s; size_t l, m; } b200_str_t;       /* growable string (kstring_t) */
int  b200_str_putsn(b200_str_t *s, const char *p, size_t n);
int  b200_str_puts(b200_str_t *s, const char *p);
int  b200_str_putc(b200_str_t *s, int c);
int  b200_str_putw(b200_str_t *s, long long v);
int  b200_str_putd(b200_str_t *s, double d);                /* htslib's kputd: what vcf_format prints for a float */

typedef struct
{
    char *id;
    int  is_fmt;        /* 0 INFO, 1 FORMAT */
    int  vl, number;    /* B200_VL_*, the count for B200_VL_FIXED */
    int  type;          /* B200_HT_* */
}
b200_vdef_t;

typedef struct
{
    char **lines; int nlines, mlines;       /* the "##" lines, in output order, without line ends */
    char **samples; int nsamples;           /* output samples */
    int  *smpl_map; int n_in_samples;       /* output sample i is input column smpl_map[i] (NULL: identity) */
    b200_vdef_t *defs; int ndefs, mdefs;
}
b200_vhdr_t;

/*  parses the "##" lines and the "#CHROM" line at the start of `text`; *consumed = offset of the first record  */
b200_vhdr_t *b200_vhdr_parse(const char *text, size_t len, size_t *consumed);
void b200_vhdr_destroy(b200_vhdr_t *h);
const b200_vdef_t *b200_vhdr_def(const b200_vhdr_t *h, int is_fmt, const char *id);
int  b200_vhdr_append(b200_vhdr_t *h, const char *line);                /* bcf_hdr_append: a second definition of the same ID is ignored */
void b200_vhdr_remove(b200_vhdr_t *h, int is_fmt, const char *id);      /* bcf_hdr_remove */
/*  bcf_hdr_subset + bcf_subset: keep the input samples map[0..n) in that order.  Returns -1 for an index out of range.  */
int  b200_vhdr_subset(b200_vhdr_t *h, int n, const int *map);
int  b200_vhdr_format(const b200_vhdr_t *h, b200_str_t *out);

typedef struct { char *key; char *val; } b200_vinfo_t;      /* val == NULL: a flag */

#define B200_FMT_TEXT 0     /* per-sample input text, untouched */
#define B200_FMT_INT  1     /* int32 vectors with missing / vector_end sentinels */
#define B200_FMT_REAL 2
#define B200_FMT_GT   3     /* htslib-encoded allele values ((allele+1)<<1 | phased), 0 = missing */
typedef struct
{
    char *key;
    int kind, n;            /* values per sample for INT / REAL / GT */
    char **txt;             /* [nsmpl] for TEXT */
    int32_t *iv;            /* [nsmpl*n] for INT / GT */
    float *fv;              /* [nsmpl*n] for REAL */
}
b200_vfmt_t;

typedef struct b200_vrec
{
    char *chrom; int64_t pos;           /* 0-based */
    char *id;
    int n_allele; char **allele;
    float qual;                         /* B200_F32_MISSING_BITS for "." */
    char *filter;
    int n_info, m_info; b200_vinfo_t *info;
    int n_fmt, m_fmt;  b200_vfmt_t *fmt;
    int nsmpl;
    /* storage: the split input line and everything allocated later */
    char *line; void **owned; int nowned, mowned;
}
b200_vrec_t;

b200_vrec_t *b200_vrec_parse(const b200_vhdr_t *h, const char *line, size_t len);
b200_vrec_t *b200_vrec_new(int nsmpl);                  /* an empty record (bcf_init) */
void b200_vrec_destroy(b200_vrec_t *r);
int  b200_vrec_format(const b200_vrec_t *r, b200_str_t *out);          /* one line with '\n' (vcf_format) */

const char *b200_vrec_info(const b200_vrec_t *r, const char *key, int *found);
/*  bcf_get_info_float / _int32: values into *dst (grown with realloc, capacity *mdst); returns their number, -1 tag absent  */
int  b200_vrec_info_floats(const b200_vrec_t *r, const char *key, float **dst, int *mdst);
int  b200_vrec_info_ints(const b200_vrec_t *r, const char *key, int32_t **dst, int *mdst);
/*  bcf_get_format_int32: nsmpl * (longest vector) values, vector_end padded; -1 tag absent, -2 not numeric  */
int  b200_vrec_fmt_ints(const b200_vrec_t *r, const char *key, int32_t **dst, int *mdst);
b200_vfmt_t *b200_vrec_fmt(const b200_vrec_t *r, const char *key);

int  b200_vrec_set_info_ints(b200_vrec_t *r, const char *key, const int32_t *v, int n);      /* n = 0 removes */
int  b200_vrec_set_info_floats(b200_vrec_t *r, const char *key, const float *v, int n);
int  b200_vrec_set_info_text(b200_vrec_t *r, const char *key, const char *val);              /* val copied; NULL = flag */
int  b200_vrec_set_fmt_ints(b200_vrec_t *r, const char *key, const int32_t *v, int nvals);   /* nvals = nsmpl * n; 0 removes */
int  b200_vrec_set_fmt_floats(b200_vrec_t *r, const char *key, const float *v, int nvals);
int  b200_vrec_set_genotypes(b200_vrec_t *r, const int32_t *gts, int nvals);                 /* bcf_update_genotypes */
int  b200_vrec_set_alleles(b200_vrec_t *r, const char *const *als, int n);                   /* bcf_update_alleles (strings copied) */

#ifdef __cplusplus
}
#endif
#endif
