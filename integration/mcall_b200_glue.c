/*  mcall_b200_glue.c -- the adapter a bcftools maintainer adds next to vcfcall.c to run `call -m` on the B200 library.
 *
 *  Written against the reference tree (call.h, mcall.c of bcftools 1.11+dev) and htslib; it is NOT part of
 *  libmcall_b200.so (htslib is not in this image).  tests/test_integration_glue.py type-checks it with
 *  `gcc -fsyntax-only` against the reference's own call.h and the htslib declarations of oracle/ref_shim/htslib/.
 *
 *  vcfcall.c changes (three call sites):
 *      vcfcall.c:697-698    mcall_init(&args->aux)             ->  b200_glue_init(&glue, &args->aux, args->out_fh)
 *      vcfcall.c:1136-1147  ret = mcall(&args.aux, bcf_rec); ... bcf_write1(...)
 *                                                               ->  b200_glue_mcall(&glue, &args.aux, bcf_rec)
 *                           (records are written from b200_glue_drain in input order; -g / -i output hooks move there too)
 *      vcfcall.c:1149, 722  before destroy_data                 ->  b200_glue_finish(&glue, &args.aux); b200_glue_destroy(&glue)
 *  The batching, pinned slabs, ploidy-vector registration, near-tie adjudication and the CUDA calls are in
 *  bcftools_b200/csrc/host/b200_call.c behind include/b200_call.h; this file only moves data between htslib records and
 *  that interface, and applies the results with the same htslib calls mcall() uses (mcall.c:1576-1684).
 */
#include <string.h>
#include <htslib/vcf.h>
#include "call.h"
#include "b200_call.h"

void mcall_trim_and_update_numberR(call_t *call, bcf1_t *rec, int nals_ori, int nals_new);      /* mcall.c:1196 */

typedef struct
{
    b200_call_t b;
    htsFile *out_fh;
    int32_t *PLs, *ADs, *itmp;  int mPLs, mADs, mitmp;
    float *QS;  int mQS;
    uint32_t *grp_off, *grp_smpl;
}
b200_glue_t;

/* replaces mcall_init(): the header lines of mcall.c:382-394 stay with the caller (call mcall_init() as well, or copy them) */
void b200_glue_init(b200_glue_t *g, call_t *call, htsFile *out_fh)
{
    memset(g, 0, sizeof *g);
    g->out_fh = out_fh;
    g->b.nsmpl       = bcf_hdr_nsamples(call->hdr);
    g->b.flag        = call->flag & (CALL_KEEPALT|CALL_VARONLY);
    g->b.output_tags = call->output_tags & (CALL_FMT_GQ|CALL_FMT_GP);
    g->b.theta       = call->theta;         /* raw -P value: the library applies mcall.c:397-416 with the init-time ploidy */
    g->b.ploidy      = call->ploidy;        /* shared storage: set_ploidy() rewrites it in place (vcfcall.c:807-825) */
    g->b.use_prior   = call->prior_AN ? 1 : 0;
    g->b.max_records = 4096;
    g->b.async_flush = 1;
    if ( call->nsmpl_grp > 1 )              /* -G: smpl_grp_t.smpl lists, group after group (call.h:53-61) */
    {
        int i, j, k = 0;
        g->grp_off  = (uint32_t*) malloc(sizeof(uint32_t)*(call->nsmpl_grp+1));
        g->grp_smpl = (uint32_t*) malloc(sizeof(uint32_t)*g->b.nsmpl);
        for (i=0; i<call->nsmpl_grp; i++)
        {
            g->grp_off[i] = k;
            for (j=0; j<call->smpl_grp[i].nsmpl; j++) g->grp_smpl[k++] = call->smpl_grp[i].smpl[j];
        }
        g->grp_off[call->nsmpl_grp] = k;
        g->b.nsmpl_grp = call->nsmpl_grp; g->b.grp_off = g->grp_off; g->b.grp_smpl = g->grp_smpl;
    }
    b200_mcall_init(&g->b);
}

/* the tail of mcall() (mcall.c:1576-1684) for one record, from the device's result */
static void b200_glue_apply(b200_glue_t *g, call_t *call, const b200_out_t *o)
{
    bcf1_t *rec = (bcf1_t*) o->user;
    int i, nsmpl = g->b.nsmpl, nals_ori = rec->n_allele, nals_new = o->ret;
    if ( o->ret<=0 ) { bcf_destroy(rec); return; }                          /* not a variant under -v: vcfcall.c:1144 */
    call->als_new = o->als_new;  call->nals_new = nals_new;
    hts_expand(int, nals_ori, call->nals_map, call->als_map);
    for (i=0; i<nals_ori; i++) call->als_map[i] = o->als_map[i];
    if ( o->als_new==1 ) bcf_update_format_int32(call->hdr, rec, "PL", NULL, 0);                    /* mcall.c:1583 */
    else
    {
        if ( (call->output_tags & CALL_FMT_GP) && o->GPs ) bcf_update_format_float(call->hdr, rec, "GP", o->GPs, o->nPLs);     /* mcall.c:1621 */
        if ( (call->output_tags & CALL_FMT_GQ) && o->GQs ) bcf_update_format_int32(call->hdr, rec, "GQ", o->GQs, nsmpl);       /* mcall.c:1623 */
        bcf_update_format_int32(call->hdr, rec, "PL", o->PLs, o->nPLs);                                                      /* mcall.c:1193 */
    }
    if ( nals_ori!=nals_new ) mcall_trim_and_update_numberR(call, rec, nals_ori, nals_new);         /* mcall.c:1627-1628 */
    rec->qual = o->qual;                                                                            /* mcall.c:1631-1645 */
    if ( nals_new>1 ) bcf_update_info_int32(call->hdr, rec, "AC", o->ac+1, nals_new-1);             /* mcall.c:1648 */
    { int32_t an = o->an; bcf_update_info_int32(call->hdr, rec, "AN", &an, 1); }                    /* mcall.c:1650 */
    hts_expand(char*, nals_new, call->nals, call->als);                                             /* mcall.c:1653-1656 */
    for (i=0; i<nals_ori; i++) if ( call->als_map[i]>=0 ) call->als[call->als_map[i]] = rec->d.allele[i];
    bcf_update_alleles(call->hdr, rec, (const char**)call->als, nals_new);
    bcf_update_genotypes(call->hdr, rec, o->gts, nsmpl*2);                                          /* mcall.c:1657 */
    if ( bcf_get_info_float(call->hdr, rec, "I16", &call->anno16, &call->n16)==16 )                 /* mcall.c:1660-1666 */
    {
        int32_t dp[4], mq;
        for (i=0; i<4; i++) dp[i] = call->anno16[i];
        bcf_update_info_int32(call->hdr, rec, "DP4", dp, 4);
        mq = (call->anno16[8]+call->anno16[10])/(call->anno16[0]+call->anno16[1]+call->anno16[2]+call->anno16[3]);
        bcf_update_info_int32(call->hdr, rec, "MQ", &mq, 1);
    }
    bcf_update_info_int32(call->hdr, rec, "I16", NULL, 0);                                          /* mcall.c:1681 */
    if ( bcf_write1(g->out_fh, call->hdr, rec)!=0 ) error("[%s] Error: failed to write the record\n", __func__);     /* vcfcall.c:1147 */
    bcf_destroy(rec);
}
static void b200_glue_drain(b200_glue_t *g, call_t *call, int n)
{
    int i;
    for (i=0; i<n; i++) { b200_out_t o; b200_mcall_result(&g->b, i, &o); b200_glue_apply(g, call, &o); }
}

/* replaces `ret = mcall(&args.aux, bcf_rec)` and the write behind it: the head of mcall() (mcall.c:1437-1537), then the queue */
void b200_glue_mcall(b200_glue_t *g, call_t *call, bcf1_t *rec)
{
    b200_rec_t r;
    int nsmpl = g->b.nsmpl, nals = rec->n_allele;
    memset(&r, 0, sizeof r);
    r.n_allele = nals;
    r.nPLs = bcf_get_format_int32(call->hdr, rec, "PL", &g->PLs, &g->mPLs);                         /* mcall.c:1444 */
    if ( r.nPLs!=nsmpl*nals*(nals+1)/2 ) error("Wrong number of PL fields? nals=%d npl=%d\n", nals, r.nPLs);
    r.PLs = g->PLs;
    if ( call->nsmpl_grp==1 )
    {
        r.nQS = bcf_get_info_float(call->hdr, rec, "QS", &g->QS, &g->mQS);                          /* mcall.c:1456 */
        if ( r.nQS<=0 ) error("The QS annotation not present at %s:%d\n", bcf_seqname(call->hdr,rec), (int)rec->pos+1);
        r.QS = g->QS;
    }
    else
    {
        r.nADs = bcf_get_format_int32(call->hdr, rec, call->sample_groups_tag, &g->ADs, &g->mADs);  /* mcall.c:1475 */
        if ( r.nADs<1 ) error("Error: FORMAT/%s is required with the -G option\n", call->sample_groups_tag);
        r.ADs = g->ADs;
    }
    r.prior_an = bcf_int32_missing;
    if ( call->prior_AN && bcf_get_info_int32(call->hdr, rec, call->prior_AN, &g->itmp, &g->mitmp)==1 && g->itmp[0]>0 )   /* mcall.c:1507-1510 */
    {
        int an = g->itmp[0];
        r.n_prior_ac = bcf_get_info_int32(call->hdr, rec, call->prior_AC, &g->itmp, &g->mitmp);
        if ( r.n_prior_ac==nals-1 ) { r.prior_an = an; r.prior_ac = g->itmp; } else r.n_prior_ac = 0;
    }
    bcf_update_info_int32(call->hdr, rec, "QS", NULL, 0);                                           /* mcall.c:1537 */
    r.user = bcf_dup(rec);          /* the synced reader reuses its buffer (vcfcall.c:478): keep a copy until the batch returns */
    g->b.unseen = call->unseen;     /* vcfcall.c:1101-1111; call->ploidy is shared storage */
    b200_glue_drain(g, call, b200_mcall(&g->b, &r));
}

void b200_glue_finish(b200_glue_t *g, call_t *call)
{
    int n;
    while ( (n = b200_mcall_flush(&g->b)) > 0 ) b200_glue_drain(g, call, n);
}
void b200_glue_destroy(b200_glue_t *g)
{
    b200_mcall_destroy(&g->b);
    free(g->PLs); free(g->ADs); free(g->itmp); free(g->QS); free(g->grp_off); free(g->grp_smpl);
}
