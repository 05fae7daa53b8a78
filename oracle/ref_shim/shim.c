/*  shim.c -- ORACLE-ONLY in-memory stand-in for the handful of htslib functions that the
 *  reference's UNMODIFIED mcall.c calls, plus a flat-array driver (ref_mcall_batch) that fills
 *  call_t the way vcfcall.c does and runs mcall() site by site.
 *
 *  TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 *  `--impl reference` legs may load the library built from this file (oracle/_ref/libmcall_ref.so).
 *  The product library (bcftools_b200/csrc) never links or calls it.
 *
 *  The reference source is compiled where it lies (/root/reference/mcall.c, see ../Makefile);
 *  nothing of it is copied here.  Behaviour of each shim follows the public htslib contract
 *  [htslib] as far as mcall.c relies on it:
 *    - getters hand out a FRESH COPY (mcall.c mutates PLs in place: mcall.c:490,517-520,1174-1187)
 *      and grow (*dst,*ndst) like htslib does;
 *    - setters capture PL/GT/GQ/GP/AN into the flat result arrays; values==NULL removes a tag;
 *    - rec->n_info = rec->n_fmt = 0 so the Number=R loops (mcall.c:1206-1259) are no-ops.
 *  Driver set-up follows vcfcall.c:652-655 (ploidy all ploidy_max=2 before mcall_init, because the
 *  prior mcall.c:397-416 reads it), vcfcall.c:807-825 (per-record ploidy), vcfcall.c:1101-1111 (unseen).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <time.h>
#include <math.h>
#include "call.h"          /* the reference's own header, found through -I/root/reference */
#include "prob1.h"
#include "mcall_b200.h"    /* flat batch/result structs shared with the product C-ABI */

/* ---------------- per-site state the getters serve / setters fill ---------------------------- */
static struct
{
    int nsmpl;
    const int32_t *pl;  int npl;
    const float   *qs;  int nqs;
    const int32_t *ad;  int nad_total;
    int prior_an;  const int32_t *prior_ac;  int n_prior_ac;
    /* captures */
    int32_t *out_pl;  int pl_dropped, pl_written;
    float   *out_gp;
    int32_t *out_gt, *out_gq, *out_an;
    const char *allele_store[64];
}
S;

#define TAG_QS 1
#define TAG_AD 2
#define TAG_PAN 3
#define TAG_PAC 4
static const char *PRIOR_AN = "AN_PRIOR", *PRIOR_AC = "AC_PRIOR", *GRP_FILE = "@mem";

void error(const char *format, ...)
{
    va_list ap; va_start(ap, format); vfprintf(stderr, format, ap); va_end(ap);
    exit(-1);
}

static void *grow(void **dst, int *ndst, int n, size_t size)
{
    if ( *ndst < n ) { *dst = realloc(*dst, (size_t)n*size); *ndst = n; }
    return *dst;
}

int bcf_get_format_values(const bcf_hdr_t *hdr, bcf1_t *line, const char *tag, void **dst, int *ndst, int type)
{
    if ( !strcmp(tag,"PL") && S.pl )
    {
        memcpy(grow(dst,ndst,S.npl,4), S.pl, (size_t)S.npl*4);
        return S.npl;
    }
    if ( !strcmp(tag,"AD") && S.ad )
    {
        memcpy(grow(dst,ndst,S.nad_total,4), S.ad, (size_t)S.nad_total*4);
        return S.nad_total;
    }
    return -1;
}
int bcf_get_info_values(const bcf_hdr_t *hdr, bcf1_t *line, const char *tag, void **dst, int *ndst, int type)
{
    if ( !strcmp(tag,"QS") )
    {
        if ( S.nqs<=0 ) return -1;
        memcpy(grow(dst,ndst,S.nqs,4), S.qs, (size_t)S.nqs*4);
        return S.nqs;
    }
    if ( !strcmp(tag,PRIOR_AN) )
    {
        if ( S.prior_an==bcf_int32_missing ) return -1;
        ((int32_t*)grow(dst,ndst,1,4))[0] = S.prior_an;
        return 1;
    }
    if ( !strcmp(tag,PRIOR_AC) )
    {
        if ( S.n_prior_ac<=0 ) return -1;
        memcpy(grow(dst,ndst,S.n_prior_ac,4), S.prior_ac, (size_t)S.n_prior_ac*4);
        return S.n_prior_ac;
    }
    return -1;      /* I16 and everything else: absent */
}
int bcf_update_format(const bcf_hdr_t *hdr, bcf1_t *line, const char *key, const void *values, int n, int type)
{
    if ( !strcmp(key,"PL") )
    {
        if ( !values ) { S.pl_dropped = 1; return 0; }
        S.pl_written = n;
        if ( S.out_pl ) memcpy(S.out_pl, values, (size_t)n*4);
    }
    else if ( !strcmp(key,"GT") ) { if ( S.out_gt ) memcpy(S.out_gt, values, (size_t)n*4); }
    else if ( !strcmp(key,"GQ") ) { if ( S.out_gq ) memcpy(S.out_gq, values, (size_t)n*4); }
    else if ( !strcmp(key,"GP") ) { if ( S.out_gp ) memcpy(S.out_gp, values, (size_t)n*4); }
    return 0;
}
int bcf_update_info(const bcf_hdr_t *hdr, bcf1_t *line, const char *key, const void *values, int n, int type)
{
    if ( values && !strcmp(key,"AN") && S.out_an ) *S.out_an = *(const int32_t*)values;
    return 0;
}
int bcf_update_alleles(const bcf_hdr_t *hdr, bcf1_t *line, const char **alleles, int nals)
{
    int i;
    for (i=0; i<nals && i<64; i++) S.allele_store[i] = alleles[i];
    line->d.allele = (char**) S.allele_store;
    line->n_allele = nals;
    return 0;
}
int bcf_hdr_append(bcf_hdr_t *h, const char *line) { return 0; }
int bcf_hdr_id2int(const bcf_hdr_t *hdr, int type, const char *id)
{
    if ( type==BCF_DT_SAMPLE )
    {
        if ( id[0]!='s' ) return -1;
        int i = atoi(id+1);
        return i>=0 && i<bcf_hdr_nsamples(hdr) ? i : -1;
    }
    if ( !strcmp(id,"QS") ) return TAG_QS;
    if ( !strcmp(id,"AD") ) return TAG_AD;
    if ( !strcmp(id,PRIOR_AN) ) return TAG_PAN;
    if ( !strcmp(id,PRIOR_AC) ) return TAG_PAC;
    return -1;
}
int bcf_hdr_idinfo_exists(const bcf_hdr_t *hdr, int type, int int_id)
{
    if ( type==BCF_HL_FMT ) return int_id==TAG_AD;
    return int_id==TAG_QS || int_id==TAG_PAN || int_id==TAG_PAC;
}
int bcf_hdr_id2length(const bcf_hdr_t *hdr, int type, int int_id) { return BCF_VL_FIXED; }
int bcf_hdr_id2type(const bcf_hdr_t *hdr, int type, int int_id) { return BCF_HT_INT; }
const char *bcf_hdr_int2id(const bcf_hdr_t *hdr, int type, int int_id) { return "?"; }
const char *bcf_seqname(const bcf_hdr_t *hdr, const bcf1_t *rec) { return "chr"; }
int bcf_get_variant_types(bcf1_t *rec) { return 0; }

/* unreachable without -a PV4 / -C alleles */
int test16(float *anno16, anno16_t *a) { return -1; }
vcmp_t *vcmp_init(void) { return NULL; }
void vcmp_destroy(vcmp_t *vcmp) { }
int vcmp_set_ref(vcmp_t *vcmp, char *ref1, char *ref2) { return -1; }
int vcmp_find_allele(vcmp_t *vcmp, char **als1, int nals1, char *al2) { return -1; }

/* ---------------- `-G <file>` emulation: the group file is synthesised from mcb_params -------- */
static const mcb_params *G_params;
char **hts_readlist(const char *fn, int is_file, int *_n)
{
    const mcb_params *p = G_params;
    if ( !p || strcmp(fn,GRP_FILE) ) return NULL;
    char **lines = (char**) malloc(sizeof(char*)*p->nsmpl);
    int g, n = 0;
    uint32_t i;
    for (g=0; g<p->ngroups; g++)
        for (i=p->grp_off[g]; i<p->grp_off[g+1]; i++)
        {
            char buf[64];
            snprintf(buf,sizeof buf,"s%u\tg%d", p->grp_smpl[i], g);
            lines[n++] = strdup(buf);
        }
    *_n = n;
    return lines;
}
typedef struct { char **keys; int *vals; int n, m; } smap_t;
void *khash_str2int_init(void) { return calloc(1,sizeof(smap_t)); }
void khash_str2int_destroy(void *hash)
{
    smap_t *h = (smap_t*) hash; int i;
    for (i=0; i<h->n; i++) free(h->keys[i]);
    free(h->keys); free(h->vals); free(h);
}
void khash_str2int_destroy_free(void *hash) { khash_str2int_destroy(hash); }
static int smap_find(smap_t *h, const char *s) { int i; for (i=0; i<h->n; i++) if ( !strcmp(h->keys[i],s) ) return i; return -1; }
int khash_str2int_has_key(void *hash, const char *str) { return smap_find((smap_t*)hash,str)>=0; }
int khash_str2int_get(void *hash, const char *str, int *value)
{
    smap_t *h = (smap_t*) hash; int i = smap_find(h,str);
    if ( i<0 ) return -1;
    *value = h->vals[i]; return 0;
}
int khash_str2int_set(void *hash, const char *str, int value)
{
    smap_t *h = (smap_t*) hash;
    if ( h->n==h->m ) { h->m = h->m ? 2*h->m : 16; h->keys = (char**)realloc(h->keys,sizeof(char*)*h->m); h->vals = (int*)realloc(h->vals,sizeof(int)*h->m); }
    h->keys[h->n] = strdup(str); h->vals[h->n] = value; h->n++;
    return 0;
}

/* ---------------- flat driver ---------------------------------------------------------------- */
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC,&t); return t.tv_sec + 1e-9*t.tv_nsec; }

/*  Runs the reference mcall() over a batch.  ploidy_tab: [nploidy][nsmpl] vectors addressed by
 *  batch->ploidy_id (NULL = all diploid).  *secs (optional) receives the time spent inside mcall()
 *  only.  Returns 0, or -1 on bad input.  site range [site_beg,site_end) lets several processes
 *  shard one batch.                                                                             */
int ref_mcall_batch(const mcb_params *p, const uint8_t *ploidy_tab, int nploidy,
                    const mcb_batch *b, const mcb_result *r, int site_beg, int site_end, double *secs)
{
    int i, j, nsmpl = p->nsmpl;
    static char *dummy_als[MCB_MAX_NALS+1];
    static char dummy_names[MCB_MAX_NALS+1][8];
    for (i=0; i<=MCB_MAX_NALS; i++) { snprintf(dummy_names[i],8,"A%d",i); dummy_als[i] = dummy_names[i]; }

    bcf_hdr_t hdr; memset(&hdr,0,sizeof hdr);
    hdr.n[BCF_DT_SAMPLE] = nsmpl;
    hdr.samples = (char**) malloc(sizeof(char*)*nsmpl);
    for (i=0; i<nsmpl; i++) { char buf[32]; snprintf(buf,sizeof buf,"s%d",i); hdr.samples[i] = strdup(buf); }
    bcf_idpair_t ctg = { "chr" };
    hdr.id[BCF_DT_CTG] = &ctg; hdr.n[BCF_DT_CTG] = 1;

    call_t aux; memset(&aux,0,sizeof aux);
    aux.hdr   = &hdr;
    aux.theta = p->theta;
    aux.flag  = p->flag;
    aux.output_tags = p->output_tags;
    aux.ploidy = (uint8_t*) malloc(nsmpl);
    for (i=0; i<nsmpl; i++) aux.ploidy[i] = p->init_ploidy ? p->init_ploidy[i] : 2;   /* vcfcall.c:652-655 */
    if ( p->ngroups>1 ) { G_params = p; aux.sample_groups = (char*)GRP_FILE; aux.sample_groups_tag = (char*)"AD"; }
    if ( p->use_prior ) { aux.prior_AN = (char*)PRIOR_AN; aux.prior_AC = (char*)PRIOR_AC; }
    mcall_init(&aux);

    bcf1_t rec; memset(&rec,0,sizeof rec);
    double t_call = 0;
    int last_pid = -1;
    for (i=site_beg; i<site_end; i++)
    {
        int nals = b->nals[i], ngt = nals*(nals+1)/2;
        int pid = b->ploidy_id ? b->ploidy_id[i] : 0;
        if ( pid!=last_pid )
        {
            if ( ploidy_tab && pid<nploidy ) memcpy(aux.ploidy, ploidy_tab + (size_t)pid*nsmpl, nsmpl);
            else memset(aux.ploidy, 2, nsmpl);
            last_pid = pid;
        }
        aux.unseen = b->unseen ? b->unseen[i] : 0;
        rec.pos = i; rec.rid = 0; rec.qual = 0; rec.n_info = 0; rec.n_fmt = 0; rec.n_sample = nsmpl;
        rec.n_allele = nals; rec.d.allele = dummy_als;

        S.nsmpl = nsmpl;
        S.pl = b->pl + b->pl_off[i]; S.npl = nsmpl*ngt;
        S.nqs = b->nqs ? b->nqs[i] : nals; S.qs = b->qs ? b->qs + (size_t)i*p->max_nals : NULL;
        if ( !S.qs ) S.nqs = 0;
        S.ad = NULL; S.nad_total = 0;
        if ( p->ngroups>1 && b->ad ) { S.ad = b->ad + b->ad_off[i]; S.nad_total = nsmpl*b->nad[i]; }
        S.prior_an = (p->use_prior && b->prior_an) ? b->prior_an[i] : bcf_int32_missing;
        S.prior_ac = (p->use_prior && b->prior_ac) ? b->prior_ac + (size_t)i*p->max_nals : NULL;
        S.n_prior_ac = 0;
        if ( S.prior_ac ) { S.n_prior_ac = nals-1; }
        S.out_pl = r->pl ? r->pl + b->pl_off[i] : NULL;
        S.out_gp = r->gp ? r->gp + b->pl_off[i] : NULL;
        S.out_gt = r->gt ? r->gt + (size_t)i*nsmpl*2 : NULL;
        S.out_gq = r->gq ? r->gq + (size_t)i*nsmpl : NULL;
        int32_t an = 0; S.out_an = &an;
        S.pl_dropped = 0; S.pl_written = 0;

        uint32_t flags = 0;
        int ret;
        if ( p->ngroups<=1 && S.nqs<=0 ) { ret = 0; flags |= MCB_SITE_NO_QS; }      /* reference would error() out: mcall.c:1457 */
        else
        {
            double t0 = now_s();
            ret = mcall(&aux, &rec);
            t_call += now_s() - t0;
        }
        if ( nals > 32 ) flags |= MCB_SITE_TOO_MANY_ALS;
        r->ret[i] = ret;
        if ( ret<=0 )
        {
            if ( r->site_flags ) r->site_flags[i] = flags;
            continue;
        }
        if ( S.pl_dropped ) flags |= MCB_SITE_PL_DROPPED;
        if ( aux.unseen && (aux.als_new & (1u<<aux.unseen)) ) flags |= MCB_SITE_UNSEEN_SEL;
        if ( r->als_new ) r->als_new[i] = aux.als_new;
        if ( r->als_map ) for (j=0; j<p->max_nals; j++) r->als_map[(size_t)i*p->max_nals+j] = j<nals ? aux.als_map[j] : -1;
        if ( r->qual ) r->qual[i] = rec.qual;
        if ( r->ac ) for (j=0; j<p->max_nals; j++) r->ac[(size_t)i*p->max_nals+j] = j<aux.nals_new ? aux.ac[j] : 0;
        if ( r->an ) r->an[i] = an;
        if ( r->site_flags ) r->site_flags[i] = flags;
    }
    if ( secs ) *secs = t_call;

    free(aux.ploidy); aux.ploidy = NULL;
    mcall_destroy(&aux);
    for (i=0; i<nsmpl; i++) free(hdr.samples[i]);
    free(hdr.samples);
    G_params = NULL;
    return 0;
}
