/*  mcall_oracle.c -- CPU restatement of the `bcftools call -m` arithmetic on flat arrays.
 *
 *  TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 *  baseline legs may build, load or call this file.  The product library (bcftools_b200/csrc)
 *  has no CPU fallback and never links it.
 *
 *  Parity status: PINNED.  This restatement is checked (tests/test_oracle_*.py) against
 *    (1) the reference's own golden files test/<name>.out for every `call -m` case of test/test.pl:276-308
 *        that does not use -C alleles / -g (fixtures under tests/golden/, made by tests/golden/make_golden.py),
 *    (2) the reference's unmodified mcall.c compiled in place (oracle/_ref, see oracle/Makefile) on
 *        randomized inputs.
 *
 *  It is a restatement, not a copy: one record is a (PL block, QS|AD, nals, unseen, ploidy vector)
 *  tuple, there is no bcf1_t/htslib, and the steps are organised as pure functions.  The ORDER OF
 *  FLOATING-POINT OPERATIONS follows the reference exactly, because that is what decides bit-exact
 *  GT/ALT/PL and the 6-digit QUAL (SURVEY.md §8a "numeric semantics"):
 *    - qsum is float32, fa=q[a]/(q[a]+q[b]) is evaluated in float32 then widened      (mcall.c:629-633, 671-677)
 *    - sample sums are sequential in group order, one log() per sample per allele set    (mcall.c:607-611, 635-645, 680-690)
 *    - no FMA contraction (compile with plain -O2, see oracle/Makefile)
 *    - GPs round-trip through float32 before GQ's max/sum                                  (mcall.c:802, 826, 858-877)
 *    - -4.343 for QUAL, -4.34294 for GQ                                                     (mcall.c:1554, 1640, 877)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <time.h>
#include "mcall_b200.h"

#define MISSING   MCB_INT32_MISSING
#define VEC_END   MCB_INT32_VECTOR_END

typedef struct
{
    double pl2p[256];       /* mcall.c:56-61 */
    double theta;           /* log space, after the Watterson factor, mcall.c:397-416 */
}
model_t;

typedef struct
{
    float  qsum[MCB_MAX_NALS];
    double ref_lk, max_lk, lk_sum;
    uint32_t als;
    int nals;
}
group_t;

static inline int hom_index(int a) { return (a+1)*(a+2)/2 - 1; }            /* index of genotype a/a, mcall.c:605 */
static inline int gt_index(int a, int b) { return a>b ? a*(a+1)/2+b : b*(b+1)/2+a; }   /* bcf_alleles2gt [htslib] */

/*  mcall.c:397-416.  n = sum of ploidies as they are at init time.  */
static double init_theta(double theta, const uint8_t *init_ploidy, int nsmpl)
{
    if ( !(theta>0) ) return theta;
    int i, n = 0;
    if ( !init_ploidy ) n = 2*nsmpl;
    else for (i=0; i<nsmpl; i++) n += init_ploidy[i];
    double aM = 1;
    for (i=2; i<n; i++) aM += 1./i;
    theta *= aM;
    if ( theta >= 1 ) theta = 0.99;
    return log(theta);
}

static inline double pl_to_p(const model_t *m, int32_t pl)
{
    return pl < 256 ? m->pl2p[pl] : pow(10., -pl/10.);      /* mcall.c:472 */
}

/*  One sample of set_pdg(), mcall.c:460-543.  pl[] is read-write: missing values are replaced
 *  by the unseen-allele likelihoods (mcall.c:495-527) and the replacement is what later gets
 *  trimmed and written out.  Returns 1 when the sample carries data, 0 when pdg was zeroed.   */
static int sample_pdg(const model_t *m, int32_t *pl, double *pdg, int ngt, int nals, int unseen)
{
    int j;
    double sum = 0;
    for (j=0; j<ngt; j++)
    {
        if ( pl[j]==VEC_END ) { j = 0; break; }     /* not diploid-shaped: treat as missing, mcall.c:465-470 */
        if ( pl[j]==MISSING ) break;
        pdg[j] = pl_to_p(m, pl[j]);
        sum += pdg[j];
    }
    if ( j==0 ) { j = ngt; sum = ngt; }             /* first value missing => all missing, mcall.c:476-481 */
    /* `unseen<0` (mcall.c:482) can never hold: call->unseen is uint8_t with 0 = none */
    if ( j<ngt )
    {
        int ia, ib, k;
        j = 0; sum = 0;
        for (ia=0; ia<nals; ia++)
            for (ib=0; ib<=ia; ib++)
            {
                if ( pl[j]==MISSING )
                {
                    k = gt_index(ia,unseen);
                    if ( pl[k]==MISSING ) k = gt_index(ib,unseen);
                    if ( pl[k]==MISSING ) k = gt_index(unseen,unseen);
                    pl[j] = pl[k]==MISSING ? 255 : pl[k];
                }
                /* the reference indexes pl2p[] unguarded here (mcall.c:522); values >255 are
                   out of its defined domain, we use the guarded form */
                pdg[j] = pl_to_p(m, pl[j]);
                sum += pdg[j];
                j++;
            }
    }
    if ( sum==ngt )     /* PL=0,0,..,0 or all missing: no data, mcall.c:529-537 */
    {
        for (j=0; j<ngt; j++) pdg[j] = 0;
        return 0;
    }
    for (j=0; j<ngt; j++) pdg[j] /= sum;
    return 1;
}

static inline double logsumexp2(double a, double b)        /* mcall.c:573-579 */
{
    if ( a>b ) return log(1 + exp(b-a)) + a;
    return log(1 + exp(a-b)) + b;
}

/*  Allele-set search for one group, mcall.c:591-710.  smpl[0..ns) are the group's samples.  */
static void best_allele_set(const model_t *m, const double *pdg, const uint8_t *ploidy,
                            const uint32_t *smpl, int ns, int nals, group_t *g)
{
    int ngt = nals*(nals+1)/2;
    int a, b, c, s;
    uint32_t max_als = 0;
    double ref_lk = -HUGE_VAL, max_lk = -HUGE_VAL, lk_sum = -HUGE_VAL;
    const float *q = g->qsum;

#define CONSIDER(mask,in_sum) do { \
        if ( max_lk<lk && set ) { max_lk = lk; max_als = (mask); } \
        if ( in_sum ) lk_sum = logsumexp2(lk, lk_sum); \
    } while (0)

    for (a=0; a<nals; a++)                                  /* one allele, mcall.c:601-615 */
    {
        double lk = 0; int set = 0, aa = hom_index(a);
        for (s=0; s<ns; s++)
        {
            double p = pdg[(size_t)smpl[s]*ngt + aa];
            if ( p ) { lk += log(p); set = 1; }
        }
        if ( a==0 ) ref_lk = lk; else lk += m->theta;
        CONSIDER(1u<<a, a>0 && set);
    }
    for (a=0; a<nals; a++)                                  /* two alleles, mcall.c:618-651 */
    {
        if ( q[a]==0 ) continue;
        int aa = hom_index(a);
        for (b=0; b<a; b++)
        {
            if ( q[b]==0 ) continue;
            double lk = 0; int set = 0;
            double fa = q[a]/(q[a]+q[b]);                   /* float32 expression, then widened */
            double fb = q[b]/(q[a]+q[b]);
            double fa2 = fa*fa, fb2 = fb*fb, fab = 2*fa*fb;
            int bb = hom_index(b), ab = aa - a + b;
            for (s=0; s<ns; s++)
            {
                const double *p = pdg + (size_t)smpl[s]*ngt;
                int pld = ploidy ? ploidy[smpl[s]] : 2;
                double val = 0;
                if ( pld==2 ) val = fa2*p[aa] + fb2*p[bb] + fab*p[ab];
                else if ( pld==1 ) val = fa*p[aa] + fb*p[bb];
                if ( val ) { lk += log(val); set = 1; }
            }
            if ( a!=0 ) lk += m->theta;
            if ( b!=0 ) lk += m->theta;
            CONSIDER(1u<<a|1u<<b, set);
        }
    }
    for (a=0; a<nals; a++)                                  /* three alleles, mcall.c:654-698 */
    {
        if ( q[a]==0 ) continue;
        int aa = hom_index(a);
        for (b=0; b<a; b++)
        {
            if ( q[b]==0 ) continue;
            int bb = hom_index(b), ab = aa - a + b;
            for (c=0; c<b; c++)
            {
                if ( q[c]==0 ) continue;
                double lk = 0; int set = 0;
                double fa = q[a]/(q[a]+q[b]+q[c]);
                double fb = q[b]/(q[a]+q[b]+q[c]);
                double fc = q[c]/(q[a]+q[b]+q[c]);
                double fa2 = fa*fa, fb2 = fb*fb, fc2 = fc*fc;
                double fab = 2*fa*fb, fac = 2*fa*fc, fbc = 2*fb*fc;
                int cc = hom_index(c), ac = aa - a + c, bc = bb - b + c;
                for (s=0; s<ns; s++)
                {
                    const double *p = pdg + (size_t)smpl[s]*ngt;
                    int pld = ploidy ? ploidy[smpl[s]] : 2;
                    double val = 0;
                    if ( pld==2 ) val = fa2*p[aa] + fb2*p[bb] + fc2*p[cc] + fab*p[ab] + fac*p[ac] + fbc*p[bc];
                    else if ( pld==1 ) val = fa*p[aa] + fb*p[bb] + fc*p[cc];
                    if ( val ) { lk += log(val); set = 1; }
                }
                if ( a!=0 ) lk += m->theta;
                if ( b!=0 ) lk += m->theta;
                if ( c!=0 ) lk += m->theta;
                CONSIDER(1u<<a|1u<<b|1u<<c, set);
            }
        }
    }
#undef CONSIDER
    g->max_lk = max_lk; g->ref_lk = ref_lk; g->lk_sum = lk_sum; g->als = max_als;
    g->nals = 0;
    for (a=0; a<nals; a++) if ( max_als & 1u<<a ) g->nals++;
}

static inline float f32_bits(uint32_t b) { union { uint32_t i; float f; } u; u.i = b; return u.f; }

/*  Genotypes of one group, mcall.c:745-886.  gps is the site's [nsmpl][ngt_new] float scratch.  */
static void call_group_genotypes(const double *pdg, const uint8_t *ploidy, const uint32_t *smpl, int ns,
                                 int nals, int nals_new, const int *als_map, const group_t *g,
                                 int32_t *gts, int32_t *ac, float *gps, int32_t *gqs, uint32_t output_tags)
{
    int ngt = nals*(nals+1)/2, ngt_new = nals_new*(nals_new+1)/2;
    int s, a, b, i;
    for (s=0; s<ns; s++)
    {
        int is = smpl[s];
        const double *p = pdg + (size_t)is*ngt;
        float *gp = gps + (size_t)is*ngt_new;
        int32_t *gt = gts + (size_t)is*2;
        int pld = ploidy ? ploidy[is] : 2;
        if ( !pld ) { gt[0] = MCB_GT_MISSING; gt[1] = VEC_END; gp[0] = -1; continue; }
        for (i=0; i<ngt; i++) if ( p[i]!=0.0 ) break;
        if ( i==ngt )       /* zero depth, mcall.c:776-784 */
        {
            gt[0] = MCB_GT_MISSING; gt[1] = pld==2 ? MCB_GT_MISSING : VEC_END; gp[0] = -1;
            continue;
        }
        gt[0] = MCB_GT_UNPHASED(0);
        gt[1] = pld==2 ? MCB_GT_UNPHASED(0) : VEC_END;
        double best = 0;
        for (a=0; a<nals; a++)          /* homozygous / haploid, mcall.c:793-808 */
        {
            if ( !(g->als & 1u<<a) ) continue;
            double lk = pld==2 ? p[hom_index(a)]*g->qsum[a]*g->qsum[a] : p[hom_index(a)]*g->qsum[a];
            int igt = pld==2 ? gt_index(als_map[a],als_map[a]) : als_map[a];
            gp[igt] = lk;
            if ( best < lk ) { best = lk; gt[0] = MCB_GT_UNPHASED(als_map[a]); }
        }
        if ( pld==2 )
        {
            gt[1] = gt[0];
            for (a=0; a<nals; a++)      /* heterozygous, mcall.c:812-834 */
            {
                if ( !(g->als & 1u<<a) ) continue;
                for (b=0; b<a; b++)
                {
                    if ( !(g->als & 1u<<b) ) continue;
                    double lk = 2*p[hom_index(a)-a+b]*g->qsum[a]*g->qsum[b];
                    gp[gt_index(als_map[a],als_map[b])] = lk;
                    if ( best < lk ) { best = lk; gt[0] = MCB_GT_UNPHASED(als_map[b]); gt[1] = MCB_GT_UNPHASED(als_map[a]); }
                }
            }
        }
        else gt[1] = VEC_END;
        ac[(gt[0]>>1)-1]++;
        if ( gt[1]!=VEC_END ) ac[(gt[1]>>1)-1]++;
    }
    if ( !(output_tags & (MCB_CALL_FMT_GQ|MCB_CALL_FMT_GP)) ) return;
    for (s=0; s<ns; s++)                /* mcall.c:843-885 */
    {
        int is = smpl[s];
        float *gp = gps + (size_t)is*ngt_new;
        int pld = ploidy ? ploidy[is] : 2;
        int nmax = pld==2 ? ngt_new : (pld==1 ? g->nals : 0);
        double max = gp[0], sum;
        if ( max<0 || nmax==0 )
        {
            if ( output_tags & MCB_CALL_FMT_GP )
            {
                for (i=0; i<nmax; i++) gp[i] = 0;
                if ( nmax==0 ) { gp[i] = f32_bits(MCB_FLOAT_MISSING_BITS); nmax++; }
                if ( nmax < ngt_new ) gp[nmax] = f32_bits(MCB_FLOAT_VECTOR_END_BITS);
            }
            gqs[is] = 0;
            continue;
        }
        sum = gp[0];
        for (i=1; i<nmax; i++) { if ( max < gp[i] ) max = gp[i]; sum += gp[i]; }
        max = -4.34294*log(1 - max/sum);
        gqs[is] = max<=INT8_MAX ? max : INT8_MAX;
        if ( output_tags & MCB_CALL_FMT_GP )
        {
            for (i=0; i<nmax; i++) gp[i] = gp[i]/sum;
            for (; i<ngt_new; i++) gp[i] = f32_bits(MCB_FLOAT_VECTOR_END_BITS);
        }
    }
}

/*  mcall.c:713-743  */
static void ref_genotypes(const double *pdg, const uint8_t *ploidy, int nsmpl, int nals, int32_t *gts, int32_t *ac)
{
    int ngt = nals*(nals+1)/2, s, i;
    for (i=0; i<nals; i++) ac[i] = 0;
    for (s=0; s<nsmpl; s++)
    {
        int pld = ploidy ? ploidy[s] : 2;
        const double *p = pdg + (size_t)s*ngt;
        for (i=0; i<ngt; i++) if ( p[i]!=0.0 ) break;
        if ( i==ngt || !pld )
        {
            gts[2*s] = MCB_GT_MISSING;
            gts[2*s+1] = pld==2 ? MCB_GT_MISSING : VEC_END;
        }
        else
        {
            gts[2*s] = MCB_GT_UNPHASED(0);
            gts[2*s+1] = pld==2 ? MCB_GT_UNPHASED(0) : VEC_END;
            ac[0] += pld;
        }
    }
}

/*  mcall.c:1158-1194, written out-of-place (src is the filled PL block).  */
static void trim_pls(const int32_t *src, int32_t *dst, const uint8_t *ploidy, int nsmpl, int nals, int nals_new, const int *pl_map)
{
    int nsrc = nals*(nals+1)/2, ndst = nals_new*(nals_new+1)/2, s, k;
    for (s=0; s<nsmpl; s++)
    {
        int pld = ploidy ? ploidy[s] : 2;
        if ( pld==2 )
            for (k=0; k<ndst; k++) dst[k] = src[pl_map[k]];
        else if ( pld==1 )
        {
            for (k=0; k<nals_new; k++) dst[k] = src[pl_map[hom_index(k)]];
            /* the reference writes one VEC_END and leaves stale values behind it (in-place rewrite);
               everything after the first VEC_END is don't-care, we pad with VEC_END */
            for (; k<ndst; k++) dst[k] = VEC_END;
        }
        else { dst[0] = MISSING; for (k=1; k<ndst; k++) dst[k] = VEC_END; }
        src += nsrc; dst += ndst;
    }
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC,&t); return t.tv_sec + 1e-9*t.tv_nsec; }

/*  Same signature as ref_mcall_batch() in oracle/ref_shim/shim.c.  Restates mcall(), mcall.c:1430-1684,
 *  minus everything that only edits the htslib record (Number=R tags, I16->DP4/MQ, PV4).        */
int oracle_mcall_batch(const mcb_params *p, const uint8_t *ploidy_tab, int nploidy,
                       const mcb_batch *b, const mcb_result *r, int site_beg, int site_end, double *secs)
{
    int nsmpl = p->nsmpl, i, j, k, ig;
    model_t m;
    for (i=0; i<256; i++) m.pl2p[i] = pow(10., -i/10.);
    m.theta = init_theta(p->theta, p->init_ploidy, nsmpl);

    int ngrp = p->ngroups>1 ? p->ngroups : 1;
    group_t *grp = (group_t*) calloc(ngrp, sizeof(group_t));
    uint32_t *all = (uint32_t*) malloc(sizeof(uint32_t)*nsmpl);
    for (i=0; i<nsmpl; i++) all[i] = i;
    uint32_t off1[2] = { 0, (uint32_t)nsmpl };
    const uint32_t *goff = ngrp>1 ? p->grp_off : off1, *gsmpl = ngrp>1 ? p->grp_smpl : all;

    int maxgt = MCB_MAX_NALS*(MCB_MAX_NALS+1)/2;
    int32_t *pl  = (int32_t*) malloc(sizeof(int32_t)*(size_t)nsmpl*maxgt);
    double  *pdg = (double*) malloc(sizeof(double)*(size_t)nsmpl*maxgt);
    float   *gps = (float*) malloc(sizeof(float)*(size_t)nsmpl*maxgt);
    int32_t *gts = (int32_t*) malloc(sizeof(int32_t)*(size_t)nsmpl*2);
    int32_t *gqs = (int32_t*) malloc(sizeof(int32_t)*(size_t)nsmpl);
    uint8_t *dip = (uint8_t*) malloc(nsmpl);
    memset(dip, 2, nsmpl);
    int als_map[MCB_MAX_NALS], pl_map[MCB_MAX_NALS*(MCB_MAX_NALS+1)/2], ac[MCB_MAX_NALS+1];

    double t0 = now_s();
    for (i=site_beg; i<site_end; i++)
    {
        int nals = b->nals[i], unseen = b->unseen ? b->unseen[i] : 0;
        int ngt = nals*(nals+1)/2;
        int pid = b->ploidy_id ? b->ploidy_id[i] : 0;
        const uint8_t *ploidy = (ploidy_tab && pid<nploidy) ? ploidy_tab + (size_t)pid*nsmpl : dip;
        uint32_t flags = 0;
        r->ret[i] = 0;
        if ( r->site_flags ) r->site_flags[i] = 0;

        if ( nals > 32 ) { if ( r->site_flags ) r->site_flags[i] = MCB_SITE_TOO_MANY_ALS; continue; }

        /* PL -> P(D|G), mcall.c:1444-1451 */
        memcpy(pl, b->pl + b->pl_off[i], sizeof(int32_t)*(size_t)nsmpl*ngt);
        for (j=0; j<nsmpl; j++) sample_pdg(&m, pl + (size_t)j*ngt, pdg + (size_t)j*ngt, ngt, nals, unseen);

        /* quality sums, mcall.c:1454-1504 */
        if ( ngrp==1 )
        {
            int nqs = b->qs ? (b->nqs ? b->nqs[i] : nals) : 0;
            if ( nqs<=0 ) { if ( r->site_flags ) r->site_flags[i] = MCB_SITE_NO_QS; continue; }
            for (j=0; j<nals; j++) grp[0].qsum[j] = j<nqs ? b->qs[(size_t)i*p->max_nals+j] : 0;
        }
        else
        {
            int nad = b->nad[i];
            const int32_t *ad = b->ad + b->ad_off[i];
            for (ig=0; ig<ngrp; ig++)
            {
                float *q = grp[ig].qsum;
                for (j=0; j<nals; j++) q[j] = 0;
                for (k=goff[ig]; k<(int)goff[ig+1]; k++)
                {
                    const int32_t *ptr = ad + (size_t)gsmpl[k]*nad;
                    float sum = 0;
                    for (j=0; j<nad; j++)
                    {
                        if ( ptr[j]==VEC_END ) break;
                        if ( ptr[j]!=MISSING ) sum += ptr[j];
                    }
                    if ( !sum ) continue;
                    for (j=0; j<nad; j++)
                    {
                        if ( ptr[j]==VEC_END ) break;
                        if ( ptr[j]!=MISSING ) q[j] += ptr[j]/sum;
                    }
                }
            }
        }
        /* -F reference-panel prior, mcall.c:1507-1527 */
        if ( p->use_prior && b->prior_an && b->prior_ac && b->prior_an[i]!=MISSING && b->prior_an[i]>0 )
        {
            int an = b->prior_an[i], ac0 = an;
            const int32_t *pac = b->prior_ac + (size_t)i*p->max_nals;
            for (j=0; j<nals-1; j++)
            {
                if ( pac[j]==VEC_END ) break;
                if ( pac[j]==MISSING ) continue;
                ac0 -= pac[j];
                for (ig=0; ig<ngrp; ig++)
                {
                    uint32_t ns = goff[ig+1]-goff[ig];
                    grp[ig].qsum[j+1] = (grp[ig].qsum[j+1] + 0.5*pac[j]) / (ns + 0.5*an);
                }
            }
            if ( ac0<0 ) { free(grp); free(all); free(pl); free(pdg); free(gps); free(gts); free(gqs); free(dip); return MCB_EPRIOR; }
            for (ig=0; ig<ngrp; ig++)
            {
                uint32_t ns = goff[ig+1]-goff[ig];
                grp[ig].qsum[0] = (grp[ig].qsum[0] + 0.5*ac0) / (ns + 0.5*an);
            }
        }
        /* normalise, mcall.c:1530-1535 */
        for (ig=0; ig<ngrp; ig++)
        {
            float sum = 0;
            for (j=0; j<nals; j++) sum += grp[ig].qsum[j];
            if ( sum ) for (j=0; j<nals; j++) grp[ig].qsum[j] /= sum;
        }

        /* per-group allele sets and the site QUAL candidates, mcall.c:1546-1561 */
        uint32_t als_new = 0;
        double ref_lk = -HUGE_VAL, lk_sum = -HUGE_VAL, max_qual = -HUGE_VAL;
        for (ig=0; ig<ngrp; ig++)
        {
            best_allele_set(&m, pdg, ploidy, gsmpl+goff[ig], goff[ig+1]-goff[ig], nals, &grp[ig]);
            als_new |= grp[ig].als;
            if ( grp[ig].max_lk==-HUGE_VAL ) continue;
            double qual = -4.343*(grp[ig].ref_lk - logsumexp2(grp[ig].lk_sum, grp[ig].ref_lk));
            if ( max_qual < qual ) { max_qual = qual; lk_sum = grp[ig].lk_sum; ref_lk = grp[ig].ref_lk; }
        }
        als_new |= 1;
        int is_variant = als_new!=1;
        if ( (p->flag & MCB_CALL_VARONLY) && !is_variant ) continue;

        int nals_new = 0;
        for (j=0; j<nals; j++)          /* mcall.c:1569-1575 */
        {
            if ( j>0 && j==unseen ) continue;
            if ( p->flag & MCB_CALL_KEEPALT ) als_new |= 1u<<j;
            if ( als_new & (1u<<j) ) nals_new++;
        }
        {                               /* trimming maps, mcall.c:547-570 */
            int nout = 0, l = 0, a, c;
            for (a=0; a<nals; a++) als_map[a] = (als_new & (1u<<a)) ? nout++ : -1;
            k = 0;
            for (a=0; a<nals; a++)
                for (c=0; c<=a; c++) { if ( (als_new & (1u<<a)) && (als_new & (1u<<c)) ) pl_map[k++] = l; l++; }
        }
        if ( unseen && (als_new & (1u<<unseen)) ) flags |= MCB_SITE_UNSEEN_SEL;

        int nAC = 0;
        for (j=0; j<=MCB_MAX_NALS; j++) ac[j] = 0;
        int32_t *out_pl = r->pl ? r->pl + b->pl_off[i] : NULL;
        if ( als_new==1 )               /* REF only, mcall.c:1580-1584 */
        {
            ref_genotypes(pdg, ploidy, nsmpl, nals, gts, ac);
            flags |= MCB_SITE_PL_DROPPED | MCB_SITE_REF_GT;
        }
        else if ( !is_variant )         /* -A kept ALTs at a non-variant site, mcall.c:1585-1589 */
        {
            ref_genotypes(pdg, ploidy, nsmpl, nals, gts, ac);
            flags |= MCB_SITE_REF_GT;
            if ( out_pl ) trim_pls(pl, out_pl, ploidy, nsmpl, nals, nals_new, pl_map);
        }
        else                            /* mcall.c:1590-1626 */
        {
            int ngt_new = nals_new*(nals_new+1)/2;
            if ( p->output_tags & (MCB_CALL_FMT_GQ|MCB_CALL_FMT_GP) )
            {
                memset(gps, 0, sizeof(float)*(size_t)nsmpl*ngt_new);
                memset(gqs, 0, sizeof(int32_t)*(size_t)nsmpl);
            }
            for (ig=0; ig<ngrp; ig++)
                call_group_genotypes(pdg, ploidy, gsmpl+goff[ig], goff[ig+1]-goff[ig], nals, nals_new, als_map,
                                     &grp[ig], gts, ac, gps, gqs, p->output_tags);
            for (j=1; j<nals_new; j++) nAC += ac[j];
            if ( !nAC && (p->flag & MCB_CALL_VARONLY) ) continue;
            if ( (p->output_tags & MCB_CALL_FMT_GP) && r->gp ) memcpy(r->gp + b->pl_off[i], gps, sizeof(float)*(size_t)nsmpl*ngt_new);
            if ( (p->output_tags & MCB_CALL_FMT_GQ) && r->gq ) memcpy(r->gq + (size_t)i*nsmpl, gqs, sizeof(int32_t)*(size_t)nsmpl);
            if ( out_pl ) trim_pls(pl, out_pl, ploidy, nsmpl, nals, nals_new, pl_map);
        }

        float qual;                     /* mcall.c:1631-1645 */
        if ( nAC ) qual = max_qual;
        else if ( lk_sum!=-HUGE_VAL ) qual = -4.343*(lk_sum - logsumexp2(lk_sum,ref_lk));
        else if ( ac[0] ) qual = m.theta ? -4.343*m.theta : 0;
        else qual = f32_bits(MCB_FLOAT_MISSING_BITS);

        r->ret[i] = nals_new;
        if ( r->als_new ) r->als_new[i] = als_new;
        if ( r->als_map ) for (j=0; j<p->max_nals; j++) r->als_map[(size_t)i*p->max_nals+j] = j<nals ? als_map[j] : -1;
        if ( r->qual ) r->qual[i] = qual;
        if ( r->ac ) for (j=0; j<p->max_nals; j++) r->ac[(size_t)i*p->max_nals+j] = j<nals_new ? ac[j] : 0;
        if ( r->an ) r->an[i] = nAC + ac[0];
        if ( r->site_flags ) r->site_flags[i] = flags;
        if ( r->diag ) { double *d = r->diag + (size_t)i*4; d[0] = max_qual; d[1] = lk_sum; d[2] = ref_lk; d[3] = 0; }
        if ( r->gt ) memcpy(r->gt + (size_t)i*nsmpl*2, gts, sizeof(int32_t)*(size_t)nsmpl*2);
    }
    if ( secs ) *secs = now_s() - t0;
    free(grp); free(all); free(pl); free(pdg); free(gps); free(gts); free(gqs); free(dip);
    return 0;
}

double oracle_theta(double theta, const uint8_t *init_ploidy, int nsmpl) { return init_theta(theta, init_ploidy, nsmpl); }
