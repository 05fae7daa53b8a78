/*  ORACLE-ONLY stub of the htslib <htslib/vcf.h> API surface that /root/reference/mcall.c
 *  (and call.h, prob1.h) needs to COMPILE UNMODIFIED.  It is test infrastructure: not ABI
 *  compatible with real htslib, never linked into the product library.  Declarations follow
 *  the public htslib API [htslib]; the in-memory behaviour is in ../shim.c.              */
#ifndef ORACLE_STUB_VCF_H
#define ORACLE_STUB_VCF_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <limits.h>
#include "hts.h"

#define BCF_HL_FLT  0
#define BCF_HL_INFO 1
#define BCF_HL_FMT  2
#define BCF_HL_CTG  3
#define BCF_HT_FLAG 0
#define BCF_HT_INT  1
#define BCF_HT_REAL 2
#define BCF_HT_STR  3
#define BCF_VL_FIXED 0
#define BCF_VL_VAR   1
#define BCF_VL_A     2
#define BCF_VL_G     3
#define BCF_VL_R     4
#define BCF_DT_ID     0
#define BCF_DT_CTG    1
#define BCF_DT_SAMPLE 2
#define BCF_UN_ALL 15
#define VCF_REF   0
#define VCF_SNP   1
#define VCF_MNP   2
#define VCF_INDEL 4
#define VCF_OTHER 8

typedef struct { const char *key; } bcf_idpair_t;
typedef struct {
    int32_t n[3];
    bcf_idpair_t *id[3];
    char **samples;
} bcf_hdr_t;

typedef struct { int type, n; } bcf_variant_t;
typedef struct { int id; int n, size, type; uint8_t *p; } bcf_fmt_t;
typedef struct { int key; int type; int len; } bcf_info_t;
typedef struct {
    char **allele;
    bcf_info_t *info;
    bcf_fmt_t *fmt;
    bcf_variant_t *var;
    int n_var, var_type;
} bcf_dec_t;
typedef struct {
    hts_pos_t pos;
    int32_t rid;
    float qual;
    uint32_t n_info:16, n_allele:16;
    uint32_t n_fmt:8, n_sample:24;
    bcf_dec_t d;
} bcf1_t;

#define bcf_int32_missing     INT32_MIN
#define bcf_int32_vector_end  (INT32_MIN+1)
#define bcf_float_missing     0x7F800001
#define bcf_float_vector_end  0x7F800002
static inline void bcf_float_set(float *ptr, uint32_t value) { union { uint32_t i; float f; } u; u.i = value; *ptr = u.f; }
#define bcf_float_set_vector_end(x) bcf_float_set(&(x),bcf_float_vector_end)
#define bcf_float_set_missing(x)    bcf_float_set(&(x),bcf_float_missing)

#define bcf_gt_phased(idx)      (((idx)+1)<<1|1)
#define bcf_gt_unphased(idx)    (((idx)+1)<<1)
#define bcf_gt_missing          0
#define bcf_gt_is_missing(val)  ((val)>>1 ? 0 : 1)
#define bcf_gt_is_phased(idx)   ((idx)&1)
#define bcf_gt_allele(val)      (((val)>>1)-1)
#define bcf_alleles2gt(a,b) ((a)>(b)?((a)*((a)+1)/2+(b)):((b)*((b)+1)/2+(a)))
static inline void bcf_gt2alleles(int igt, int *a, int *b)
{
    int k = 0, dk = 1;
    while ( k<igt ) { dk++; k += dk; }
    *b = dk - 1; *a = igt - k + *b;
}

#define bcf_hdr_nsamples(hdr) ((hdr)->n[BCF_DT_SAMPLE])

int bcf_hdr_append(bcf_hdr_t *h, const char *line);
int bcf_hdr_id2int(const bcf_hdr_t *hdr, int type, const char *id);
int bcf_hdr_idinfo_exists(const bcf_hdr_t *hdr, int type, int int_id);
int bcf_hdr_id2length(const bcf_hdr_t *hdr, int type, int int_id);
int bcf_hdr_id2type(const bcf_hdr_t *hdr, int type, int int_id);
const char *bcf_hdr_int2id(const bcf_hdr_t *hdr, int type, int int_id);
const char *bcf_seqname(const bcf_hdr_t *hdr, const bcf1_t *rec);
int bcf_get_variant_types(bcf1_t *rec);

int bcf_get_format_values(const bcf_hdr_t *hdr, bcf1_t *line, const char *tag, void **dst, int *ndst, int type);
int bcf_get_info_values(const bcf_hdr_t *hdr, bcf1_t *line, const char *tag, void **dst, int *ndst, int type);
int bcf_update_format(const bcf_hdr_t *hdr, bcf1_t *line, const char *key, const void *values, int n, int type);
int bcf_update_info(const bcf_hdr_t *hdr, bcf1_t *line, const char *key, const void *values, int n, int type);
int bcf_update_alleles(const bcf_hdr_t *hdr, bcf1_t *line, const char **alleles, int nals);

#define bcf_get_format_int32(hdr,line,tag,dst,ndst)  bcf_get_format_values(hdr,line,tag,(void**)(dst),ndst,BCF_HT_INT)
#define bcf_get_format_float(hdr,line,tag,dst,ndst)  bcf_get_format_values(hdr,line,tag,(void**)(dst),ndst,BCF_HT_REAL)
#define bcf_get_info_int32(hdr,line,tag,dst,ndst)    bcf_get_info_values(hdr,line,tag,(void**)(dst),ndst,BCF_HT_INT)
#define bcf_get_info_float(hdr,line,tag,dst,ndst)    bcf_get_info_values(hdr,line,tag,(void**)(dst),ndst,BCF_HT_REAL)
#define bcf_update_format_int32(hdr,line,key,values,n) bcf_update_format((hdr),(line),(key),(values),(n),BCF_HT_INT)
#define bcf_update_format_float(hdr,line,key,values,n) bcf_update_format((hdr),(line),(key),(values),(n),BCF_HT_REAL)
#define bcf_update_genotypes(hdr,line,gts,n)           bcf_update_format((hdr),(line),"GT",(gts),(n),BCF_HT_INT)
#define bcf_update_info_int32(hdr,line,key,values,n)   bcf_update_info((hdr),(line),(key),(values),(n),BCF_HT_INT)
#define bcf_update_info_float(hdr,line,key,values,n)   bcf_update_info((hdr),(line),(key),(values),(n),BCF_HT_REAL)
#endif
