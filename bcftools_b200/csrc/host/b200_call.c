/*  b200_call.c -- the host batcher (plain C): mirrors mcall_init / mcall / mcall_destroy (call.h:131-147) over
 *  pinned structure-of-arrays slabs and the C-ABI of mcall_b200.h.  See include/b200_call.h.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include "b200_call.h"

struct b200_batcher
{
    mcb_ctx *ctx;
    int nsmpl, max_nals, cap, n, nready;
    /* pinned input slabs (structure of arrays, site x sample x genotype) */
    int32_t *pl;   int64_t pl_cap, pl_used;   int64_t *pl_off;
    int32_t *ad;   int64_t ad_cap, ad_used;   int64_t *ad_off;   uint8_t *nad;
    uint8_t *nals, *unseen, *nqs;   uint16_t *ploidy_id;
    float *qs;   int32_t *prior_an, *prior_ac;
    void **user;
    /* pinned result slabs */
    mcb_result res;
    /* ploidy vectors registered with the context: a new id whenever the driver changed call->ploidy */
    uint8_t *last_ploidy;   int n_ploidy;
    int grouped, use_prior, typed;
};

static void default_handler(const char *msg) { fputs(msg, stderr); exit(-1); }      /* version.c:40-47 */
static void (*g_handler)(const char *msg) = default_handler;
void b200_set_error_handler(void (*handler)(const char *msg)) { g_handler = handler ? handler : default_handler; }

static void b200_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_handler(buf);
}

static void *pinned(size_t bytes)
{
    void *p = mcb_host_alloc(bytes ? bytes : 1);
    if ( !p ) b200_error("b200: could not allocate %zu bytes of pinned memory\n", bytes);
    return p;
}

void b200_mcall_init(b200_call_t *call)
{
    b200_batcher_t *b = (b200_batcher_t*) calloc(1, sizeof *b);
    b->nsmpl = call->nsmpl;
    b->max_nals = call->max_nals>0 ? call->max_nals : 5;
    b->cap = call->max_records>0 ? call->max_records : 4096;
    b->grouped = call->nsmpl_grp > 1;
    b->use_prior = call->use_prior;
    b->typed = call->bcf_typed && !b->grouped;

    mcb_params p; memset(&p, 0, sizeof p);
    p.nsmpl = call->nsmpl;  p.max_nals = b->max_nals;
    p.theta = call->theta;                      /* raw: mcb_init applies the Watterson factor and log (mcall.c:397-416) */
    p.init_ploidy = call->ploidy;               /* as it is at init time: all ploidy_max, vcfcall.c:652-655 */
    p.flag = call->flag & (CALL_KEEPALT|CALL_VARONLY);
    p.output_tags = call->output_tags & (CALL_FMT_GQ|CALL_FMT_GP);
    p.ngroups = b->grouped ? call->nsmpl_grp : 1;
    p.grp_off = call->grp_off;  p.grp_smpl = call->grp_smpl;
    p.use_prior = call->use_prior;  p.device = call->device;
    int rc = mcb_init(&b->ctx, &p);
    if ( rc ) b200_error("b200_mcall_init: %s (%s)\n", mcb_strerror(rc), b->ctx ? mcb_last_cuda_error(b->ctx) : "");

    const int S = b->nsmpl, M = b->max_nals, R = b->cap;
    const int64_t gmax = (int64_t)M*(M+1)/2;
    b->pl_cap = (int64_t)R*(((int64_t)S*gmax + 7) & ~7ll);      /* sites start on 16-byte boundaries (8 int16 / 4 int32 elements) */
    b->pl = (int32_t*) pinned((size_t)b->pl_cap*(b->typed ? 2 : 4));      /* typed: an int16 slab (mcb_batch.pl_type = 2) */
    b->pl_off = (int64_t*) pinned(sizeof(int64_t)*R);
    if ( b->grouped )
    {
        b->ad_cap = (int64_t)R*(((int64_t)S*M + 3) & ~3ll);
        b->ad = (int32_t*) pinned((size_t)b->ad_cap*4);
        b->ad_off = (int64_t*) pinned(sizeof(int64_t)*R);
        b->nad = (uint8_t*) pinned(R);
    }
    b->nals = (uint8_t*) pinned(R);  b->unseen = (uint8_t*) pinned(R);  b->nqs = (uint8_t*) pinned(R);
    b->ploidy_id = (uint16_t*) pinned(2*(size_t)R);
    b->qs = (float*) pinned(sizeof(float)*(size_t)R*M);
    b->prior_an = (int32_t*) pinned(4*(size_t)R);  b->prior_ac = (int32_t*) pinned(4*(size_t)R*M);
    b->user = (void**) calloc(R, sizeof(void*));
    b->res.ret = (int32_t*) pinned(4*(size_t)R);          b->res.als_new = (uint32_t*) pinned(4*(size_t)R);
    b->res.als_map = (int8_t*) pinned((size_t)R*M);       b->res.qual = (float*) pinned(4*(size_t)R);
    b->res.ac = (int32_t*) pinned(4*(size_t)R*M);         b->res.an = (int32_t*) pinned(4*(size_t)R);
    b->res.site_flags = (uint32_t*) pinned(4*(size_t)R);  b->res.diag = NULL;
    if ( b->typed )
    {
        b->res.gt8 = (int8_t*) pinned(2*(size_t)R*S);
        b->res.gq8 = (p.output_tags & (CALL_FMT_GQ|CALL_FMT_GP)) ? (int8_t*) pinned((size_t)R*S) : NULL;
        b->res.pl16 = (int16_t*) pinned((size_t)b->pl_cap*2);
    }
    else
    {
    b->res.gt = (int32_t*) pinned(8*(size_t)R*S);
    b->res.gq = (p.output_tags & (CALL_FMT_GQ|CALL_FMT_GP)) ? (int32_t*) pinned(4*(size_t)R*S) : NULL;
    b->res.pl = (int32_t*) pinned((size_t)b->pl_cap*4);
    }
    b->res.gp = (p.output_tags & CALL_FMT_GP) ? (float*) pinned((size_t)b->pl_cap*4) : NULL;
    b->res.pl_off_out = (int64_t*) pinned(sizeof(int64_t)*R);
    b->last_ploidy = (uint8_t*) malloc(S);
    memset(b->last_ploidy, 2, S);               /* id 0 of the context = all diploid */
    b->n_ploidy = 0;
    call->batcher = b;
}

int b200_mcall_flush(b200_call_t *call)
{
    b200_batcher_t *b = call->batcher;
    b->nready = 0;
    if ( !b->n ) return 0;
    mcb_batch in; memset(&in, 0, sizeof in);
    in.pl_type = b->typed ? 2 : 4;
    in.nsites = b->n;  in.pl = b->pl;  in.pl_off = b->pl_off;  in.nals = b->nals;  in.unseen = b->unseen;
    in.ploidy_id = b->ploidy_id;
    if ( b->grouped ) { in.ad = b->ad; in.ad_off = b->ad_off; in.nad = b->nad; }
    else { in.qs = b->qs; in.nqs = b->nqs; }
    if ( b->use_prior ) { in.prior_an = b->prior_an; in.prior_ac = b->prior_ac; }
    int rc = mcb_call_host(b->ctx, &in, &b->res);
    if ( rc ) b200_error("b200_mcall: %s (%s)\n", mcb_strerror(rc), mcb_last_cuda_error(b->ctx));
    b->nready = b->n;
    b->n = 0;  b->pl_used = 0;  b->ad_used = 0;
    return b->nready;
}

int b200_mcall(b200_call_t *call, const b200_rec_t *rec)
{
    b200_batcher_t *b = call->batcher;
    const int S = b->nsmpl, M = b->max_nals, nals = rec->n_allele;
    if ( b->nready ) b->nready = 0;             /* results of the previous flush are gone once a new record is queued */
    if ( nals<1 || nals>M ) b200_error("b200_mcall: %d alleles, the batcher was initialised for at most %d\n", nals, M);
    const int ngt = nals*(nals+1)/2;
    if ( rec->nPLs != S*ngt )                   /* mcall.c:1445-1446 */
        b200_error("Wrong number of PL fields? nals=%d npl=%d\n", nals, rec->nPLs);
    int i = b->n;
    b->pl_off[i] = b->pl_used;
    if ( b->typed )             /* the record's own typed vector goes into the int16 slab; int8 is widened here */
    {
        int16_t *dst = (int16_t*) b->pl + b->pl_used;
        if ( !rec->PL_typed || (rec->PL_bt!=1 && rec->PL_bt!=2) ) b200_error("b200_mcall: bcf_typed needs FORMAT/PL as an int8 or int16 typed vector\n");
        if ( rec->PL_bt==2 ) memcpy(dst, rec->PL_typed, sizeof(int16_t)*(size_t)rec->nPLs);
        else
        {
            const int8_t *src = (const int8_t*) rec->PL_typed;
            for (int k=0; k<rec->nPLs; k++) dst[k] = src[k]==INT8_MIN ? INT16_MIN : (src[k]==INT8_MIN+1 ? INT16_MIN+1 : src[k]);
        }
        b->pl_used += ((int64_t)rec->nPLs + 7) & ~7ll;         /* 16-byte aligned sites */
    }
    else
    {
    memcpy(b->pl + b->pl_used, rec->PLs, sizeof(int32_t)*(size_t)rec->nPLs);
    b->pl_used += ((int64_t)rec->nPLs + 3) & ~3ll;
    }
    b->nals[i] = (uint8_t)nals;  b->unseen[i] = call->unseen;
    if ( b->grouped )
    {
        if ( rec->nADs < 1 || rec->nADs % S )   /* mcall.c:1476 */
            b200_error("Error: FORMAT/AD is required with the -G option, mpileup must be run with \"-a AD\" or \"-a QS\"\n");
        b->ad_off[i] = b->ad_used;  b->nad[i] = (uint8_t)(rec->nADs/S);
        memcpy(b->ad + b->ad_used, rec->ADs, sizeof(int32_t)*(size_t)rec->nADs);
        b->ad_used += ((int64_t)rec->nADs + 3) & ~3ll;
    }
    else
    {
        if ( rec->nQS<=0 ) b200_error("The QS annotation not present at record %d of the batch\n", i);   /* mcall.c:1457 */
        int nq = rec->nQS < M ? rec->nQS : M;
        memset(b->qs + (size_t)i*M, 0, sizeof(float)*M);
        memcpy(b->qs + (size_t)i*M, rec->QS, sizeof(float)*nq);
        b->nqs[i] = (uint8_t)nq;
    }
    if ( b->use_prior )
    {
        b->prior_an[i] = rec->prior_an;
        for (int j=0; j<M; j++) b->prior_ac[(size_t)i*M+j] = j<rec->n_prior_ac ? rec->prior_ac[j] : MCB_INT32_VECTOR_END;
    }
    /* set_ploidy() rewrites call->ploidy between records (vcfcall.c:807-825): register each distinct vector once */
    if ( call->ploidy && memcmp(b->last_ploidy, call->ploidy, S) )
    {
        memcpy(b->last_ploidy, call->ploidy, S);
        b->n_ploidy++;
        int rc = mcb_set_ploidy(b->ctx, b->n_ploidy, call->ploidy);
        if ( rc ) b200_error("b200_mcall: mcb_set_ploidy: %s\n", mcb_strerror(rc));
    }
    b->ploidy_id[i] = (uint16_t)b->n_ploidy;
    b->user[i] = rec->user;
    if ( ++b->n == b->cap ) return b200_mcall_flush(call);
    return 0;
}

int b200_mcall_result(b200_call_t *call, int i, b200_out_t *out)
{
    b200_batcher_t *b = call->batcher;
    if ( i<0 || i>=b->nready ) return -1;
    const int S = b->nsmpl, M = b->max_nals;
    memset(out, 0, sizeof *out);
    out->ret = b->res.ret[i];
    out->user = b->user[i];
    out->site_flags = b->res.site_flags[i];
    if ( out->ret<=0 ) return 0;
    out->als_new = b->res.als_new[i];
    out->als_map = b->res.als_map + (size_t)i*M;
    out->qual = b->res.qual[i];
    out->ac = b->res.ac + (size_t)i*M;  out->an = b->res.an[i];
    const int ref_gt = (out->site_flags & MCB_SITE_REF_GT) != 0;
    if ( b->typed )
    {
        out->gts8 = b->res.gt8 + (size_t)i*S*2;
        out->GQs8 = (b->res.gq8 && !ref_gt) ? b->res.gq8 + (size_t)i*S : NULL;
    }
    else
    {
    out->gts = b->res.gt + (size_t)i*S*2;
    out->GQs = (b->res.gq && !ref_gt) ? b->res.gq + (size_t)i*S : NULL;
    }
    if ( !(out->site_flags & MCB_SITE_PL_DROPPED) && b->res.pl_off_out[i]>=0 )
    {
        if ( b->typed ) out->PLs16 = b->res.pl16 + b->res.pl_off_out[i];
        else
        out->PLs = b->res.pl + b->res.pl_off_out[i];
        out->nPLs = S*out->ret*(out->ret+1)/2;
        if ( b->res.gp && !(out->site_flags & MCB_SITE_REF_GT) ) out->GPs = b->res.gp + b->res.pl_off_out[i];
    }
    return 0;
}

void b200_mcall_destroy(b200_call_t *call)
{
    b200_batcher_t *b = call->batcher;
    if ( !b ) return;
    mcb_destroy(b->ctx);
    mcb_host_free(b->pl); mcb_host_free(b->pl_off); mcb_host_free(b->ad); mcb_host_free(b->ad_off); mcb_host_free(b->nad);
    mcb_host_free(b->nals); mcb_host_free(b->unseen); mcb_host_free(b->nqs); mcb_host_free(b->ploidy_id);
    mcb_host_free(b->qs); mcb_host_free(b->prior_an); mcb_host_free(b->prior_ac);
    mcb_host_free(b->res.ret); mcb_host_free(b->res.als_new); mcb_host_free(b->res.als_map); mcb_host_free(b->res.qual);
    mcb_host_free(b->res.ac); mcb_host_free(b->res.an); mcb_host_free(b->res.site_flags); mcb_host_free(b->res.gt);
    mcb_host_free(b->res.gq); mcb_host_free(b->res.gp); mcb_host_free(b->res.pl); mcb_host_free(b->res.pl_off_out);
    mcb_host_free(b->res.gt8); mcb_host_free(b->res.gq8); mcb_host_free(b->res.pl16);
    free(b->user); free(b->last_ploidy); free(b);
    call->batcher = NULL;
}
