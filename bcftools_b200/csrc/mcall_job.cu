/*  mcall_job.cu -- mcall_job.h: one job over N devices, contiguous site ranges, ordered results.  Host code only.  */
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>
#include "mcall_job.h"

struct mcb_job
{
    mcb_params p;
    std::vector<mcb_ctx*> ctx;
    std::vector<int> device;
    std::string err;
    int64_t opt_pack = 0;       /* 1: close the gaps between the ranges' compacted PL / GP blocks (a host memmove of up to the whole output) */
};

extern "C" int mcb_job_init(mcb_job **out, const mcb_params *params, const int *devices, int ndevices)
{
    if ( !out || !params || !devices || ndevices < 1 ) return MCB_EINVAL;
    mcb_job *job = new mcb_job();
    job->p = *params;
    *out = job;
    for (int k=0; k<ndevices; k++)
    {
        mcb_params p = *params;
        p.device = devices[k];
        mcb_ctx *c = nullptr;
        const int rc = mcb_init(&c, &p);
        if ( c ) { job->ctx.push_back(c); job->device.push_back(devices[k]); }
        if ( rc ) { job->err = c ? mcb_last_cuda_error(c) : ""; return rc; }
    }
    return MCB_OK;
}

extern "C" void mcb_job_destroy(mcb_job *job)
{
    if ( !job ) return;
    for (mcb_ctx *c : job->ctx) mcb_destroy(c);
    delete job;
}

extern "C" int mcb_job_ndevices(const mcb_job *job) { return job ? (int)job->ctx.size() : 0; }
extern "C" const char *mcb_job_last_error(const mcb_job *job) { return job ? job->err.c_str() : ""; }

extern "C" int mcb_job_set_ploidy(mcb_job *job, int id, const uint8_t *ploidy)
{
    if ( !job ) return MCB_EINVAL;
    for (mcb_ctx *c : job->ctx) { const int rc = mcb_set_ploidy(c, id, ploidy); if ( rc ) { job->err = mcb_last_cuda_error(c); return rc; } }
    return MCB_OK;
}
extern "C" int mcb_job_set_option(mcb_job *job, const char *key, int64_t value)
{
    if ( !job || !key ) return MCB_EINVAL;
    if ( !strcmp(key, "pack") ) { job->opt_pack = value; return MCB_OK; }
    for (mcb_ctx *c : job->ctx) { const int rc = mcb_set_option(c, key, value); if ( rc ) return rc; }
    return MCB_OK;
}

/*  range boundaries: equal shares of the PL volume (a site's cost grows with its genotype count)  */
extern "C" int mcb_job_partition(int32_t nsmpl, const uint8_t *nals, int32_t nsites, int32_t nparts, int32_t *first_site)
{
    if ( nsmpl<1 || nsites<0 || nparts<1 || !first_site || (nsites && !nals) ) return MCB_EINVAL;
    std::vector<int64_t> ext((size_t)nsites + 1, 0);
    for (int i=0; i<nsites; i++) { const int64_t n = nals[i]; ext[i+1] = ext[i] + (int64_t)nsmpl*n*(n+1)/2; }
    first_site[0] = 0;
    for (int k=1, i=0; k<nparts; k++)
    {
        const int64_t want = ext[nsites]/nparts*k + ext[nsites]%nparts*k/nparts;
        while ( i<nsites && ext[i] < want ) i++;
        first_site[k] = i;
    }
    first_site[nparts] = nsites;
    return MCB_OK;
}

extern "C" int mcb_job_call_host(mcb_job *job, const mcb_batch *b, const mcb_result *r, int32_t *first_site)
{
    if ( !job || !b || !r || !r->ret || b->nsites<0 || !b->pl || !b->pl_off || !b->nals ) return MCB_EINVAL;
    const int N = (int)job->ctx.size(), R = b->nsites, S = job->p.nsmpl, M = job->p.max_nals;
    std::vector<int32_t> beg(N + 1, R);
    { const int prc = mcb_job_partition(S, b->nals, R, N, beg.data()); if ( prc ) return prc; }
    if ( first_site ) for (int k=0; k<=N; k++) first_site[k] = beg[k];
    if ( R==0 ) return MCB_OK;

    const bool compact = r->pl_off_out != nullptr;
    std::vector<int> rc(N, MCB_OK);
    std::vector<int64_t> used(N, 0), base(N, 0);
    auto run = [&](int k)
    {
        const int b0 = beg[k], n = beg[k+1] - b0;
        if ( n<=0 ) return;
        mcb_batch sb = *b;
        sb.nsites = n;
        sb.pl_off = b->pl_off + b0;  sb.nals = b->nals + b0;
        if ( b->unseen ) sb.unseen = b->unseen + b0;
        if ( b->ploidy_id ) sb.ploidy_id = b->ploidy_id + b0;
        if ( b->qs ) sb.qs = b->qs + (size_t)b0*M;
        if ( b->nqs ) sb.nqs = b->nqs + b0;
        if ( b->ad_off ) sb.ad_off = b->ad_off + b0;
        if ( b->nad ) sb.nad = b->nad + b0;
        if ( b->prior_an ) sb.prior_an = b->prior_an + b0;
        if ( b->prior_ac ) sb.prior_ac = b->prior_ac + (size_t)b0*M;
        mcb_result sr = *r;
        sr.ret = r->ret + b0;
        if ( r->als_new ) sr.als_new = r->als_new + b0;
        if ( r->als_map ) sr.als_map = r->als_map + (size_t)b0*M;
        if ( r->qual ) sr.qual = r->qual + b0;
        if ( r->ac ) sr.ac = r->ac + (size_t)b0*M;
        if ( r->an ) sr.an = r->an + b0;
        if ( r->site_flags ) sr.site_flags = r->site_flags + b0;
        if ( r->diag ) sr.diag = r->diag + (size_t)b0*4;
        if ( r->gt ) sr.gt = r->gt + (size_t)b0*S*2;
        if ( r->gq ) sr.gq = r->gq + (size_t)b0*S;
        if ( r->gt8 ) sr.gt8 = r->gt8 + (size_t)b0*S*2;
        if ( r->gq8 ) sr.gq8 = r->gq8 + (size_t)b0*S;
        if ( compact )
        {
            /* the range compacts into its own stretch of the output buffers: [pl_off[b0], pl_off[next range]) is nobody else's */
            base[k] = b->pl_off[b0] - b->pl_off[0];
            sr.pl_off_out = r->pl_off_out + b0;
            if ( r->pl ) sr.pl = r->pl + base[k];
            if ( r->pl16 ) sr.pl16 = r->pl16 + base[k];
            if ( r->gp ) sr.gp = r->gp + base[k];
        }
        rc[k] = mcb_call_host(job->ctx[k], &sb, &sr);
        if ( rc[k]==MCB_OK && compact )
        {
            int64_t st[4];
            mcb_get_stats(job->ctx[k], st);
            used[k] = st[2];
        }
    };
    std::vector<std::thread> th;
    for (int k=1; k<N; k++) th.emplace_back(run, k);
    run(0);
    for (auto &t : th) t.join();
    for (int k=0; k<N; k++)
        if ( rc[k] ) { job->err = mcb_last_cuda_error(job->ctx[k]); return rc[k]; }

    if ( compact )
    {
        /* ordered concatenation of the ranges' compacted blocks: every range compacted into its own stretch, so the blocks
           already stand in site order at ascending offsets and pl_off_out only needs the stretch's base.  Closing the gaps
           between the stretches (option pack=1) moves up to the whole output once more through the host's memory: measured
           35 ms of a 70 ms call for 768 sites x 100,000 samples on two devices (profiles/r02_job_strong_scaling.log).  */
        int64_t cum = 0;
        for (int k=0; k<N; k++)
        {
            const int64_t at = job->opt_pack ? cum : base[k];
            if ( job->opt_pack && used[k] && cum != base[k] )
            {
                if ( r->pl ) memmove(r->pl + cum, r->pl + base[k], (size_t)used[k]*4);
                if ( r->pl16 ) memmove(r->pl16 + cum, r->pl16 + base[k], (size_t)used[k]*2);
                if ( r->gp ) memmove(r->gp + cum, r->gp + base[k], (size_t)used[k]*4);
            }
            for (int i=beg[k]; i<beg[k+1]; i++) if ( r->pl_off_out[i] >= 0 ) r->pl_off_out[i] += at;
            cum += used[k];
        }
    }
    return MCB_OK;
}
