/*  mcall_kernels.cu -- the fused site kernel (phase 1 site reduction + phase 2 per-sample genotype).
 *  See mcall_kernels.cuh for the design summary and the reference line map.
 */
#include "mcall_kernels.cuh"
#include <math_constants.h>

namespace mcb {

#define I32_MISSING   INT32_MIN
#define I32_VEC_END   (INT32_MIN+1)
#define MAX_STAGE     16
#define BLOCK         256
#define NWARP         (BLOCK/32)

/* ------------------------------------------------------------------------------------------------
 *  bulk-copy engine + mbarrier (PTX; SASS: UBLKCP / SYNCS)
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init()   { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

/* ------------------------------------------------------------------------------------------------
 *  small helpers
 * ---------------------------------------------------------------------------------------------- */
__host__ __device__ constexpr int hom_idx(int a) { return (a+1)*(a+2)/2 - 1; }           /* a/a, mcall.c:605 */
__host__ __device__ constexpr int gt_idx(int a, int b) { return a>b ? a*(a+1)/2+b : b*(b+1)/2+a; }   /* bcf_alleles2gt */
__host__ __device__ constexpr int pair_idx(int a, int b) { return a*(a-1)/2 + b; }       /* a>b, enumeration order of mcall.c:620-624 */
__host__ __device__ constexpr int tri_idx(int a, int b, int c) { return a*(a-1)*(a-2)/6 + b*(b-1)/2 + c; }  /* a>b>c, mcall.c:656-665 */

/*  running product with the exponent tracked separately: log(prod) = log(M) + (E - 1023*n)*ln2  */
__device__ __forceinline__ void acc_mul(double &M, int &E, double v)
{
    int hi = __double2hiint(v), lo = __double2loint(v);
    E += hi >> 20;
    M = __dmul_rn(M, __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo));
}
__device__ __forceinline__ void acc_renorm(double &M, int &E)
{
    int hi = __double2hiint(M), lo = __double2loint(M);
    E += (hi >> 20) - 1023;
    M = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
}
__device__ __forceinline__ double logsumexp2_dev(double a, double b)       /* mcall.c:573-579 */
{
    if ( a>b ) return log(1 + exp(b-a)) + a;
    return log(1 + exp(a-b)) + b;
}

template<int NALS> struct Shape
{
    static constexpr int G     = NALS*(NALS+1)/2;
    static constexpr int NPAIR = NALS*(NALS-1)/2;
    static constexpr int NTRI  = NALS*(NALS-1)*(NALS-2)/6;
    static constexpr int NSUB  = NALS + NPAIR + NTRI;
    static constexpr int NACC  = NPAIR + NTRI + 2;          /* products: pairs, triples, N_all, N_called */
};

template<int NALS> struct Shared
{
    using S = Shape<NALS>;
    double   pl2p[256];
    double   gq_thr[128];
    uint64_t bars[MAX_STAGE];
    /* per-site coefficients of the allele sets (mcall.c:629-633, 671-677) */
    double   cf_pair[(S::NPAIR ? S::NPAIR : 1)*5];   /* fa2 fb2 fab | fa fb */
    double   cf_tri[(S::NTRI ? S::NTRI : 1)*9];      /* fa2 fb2 fc2 fab fac fbc | fa fb fc */
    uint32_t live;                                   /* bit k: pair k evaluated; bit NPAIR+k: triple k evaluated */
    /* cross-warp reduction scratch */
    double   red_M[NWARP][S::NACC];
    int      red_E[NWARP][S::NACC];
    long long red_pls[NWARP][NALS];
    int      red_cnt[NWARP][2];
    /* site decision record */
    float    qf[NALS];
    double   q[NALS];
    double   max_qual, lk_sum, ref_lk, gap;
    uint32_t grp_als, als_new, flags;
    int      grp_nals, nals_new, is_variant, ret_early, pl_dropped, ref_gt;
    int      als_map[NALS];
    int      pl_map[S::G];
    int      ac[8];
    unsigned long long ac_packed[2];
};

/* ------------------------------------------------------------------------------------------------
 *  set_pdg for one sample (mcall.c:460-543).  `row` is the sample's PL vector in shared memory;
 *  the missing-value fill of mcall.c:495-527 is written back to it because the filled values are
 *  what gets trimmed and output later.  Returns true when the sample carries data.
 * ---------------------------------------------------------------------------------------------- */
template<int NALS>
__device__ __noinline__ int fix_missing(int32_t *row, int unseen)
{
    constexpr int G = Shape<NALS>::G;
    /* first scan: a vector_end anywhere before the first missing, or a missing first value => all missing */
    int j;
    for (j=0; j<G; j++)
    {
        if ( row[j]==I32_VEC_END ) return 0;
        if ( row[j]==I32_MISSING ) break;
    }
    if ( j==0 ) return 0;
    if ( j==G ) return 1;       /* nothing missing after all (negative garbage): leave as is */
    j = 0;
    for (int ia=0; ia<NALS; ia++)
        for (int ib=0; ib<=ia; ib++)
        {
            if ( row[j]==I32_MISSING )
            {
                int k = gt_idx(ia,unseen);
                if ( row[k]==I32_MISSING ) k = gt_idx(ib,unseen);
                if ( row[k]==I32_MISSING ) k = gt_idx(unseen,unseen);
                row[j] = row[k]==I32_MISSING ? 255 : row[k];
            }
            else if ( row[j] < 0 ) return 0;    /* vector_end behind a missing value: undefined in the reference */
            j++;
        }
    return 1;
}

template<int NALS>
__device__ __forceinline__ bool load_sample(int32_t *row, int unseen, const double *s_pl2p, const DevTables *tab,
                                            int (&pl)[Shape<NALS>::G], double (&p)[Shape<NALS>::G], double &sum, uint32_t &flags)
{
    constexpr int G = Shape<NALS>::G;
    int orv = 0;
    #pragma unroll
    for (int j=0; j<G; j++) { pl[j] = row[j]; orv |= pl[j]; }
    if ( orv < 0 )
    {
        if ( !fix_missing<NALS>(row, unseen) ) return false;
        orv = 0;
        #pragma unroll
        for (int j=0; j<G; j++) { pl[j] = row[j]; orv |= pl[j]; }
    }
    if ( orv==0 ) return false;         /* PL=0,..,0: sum==n_gt, no data (mcall.c:529-537) */
    if ( orv & ~255 )
    {
        /* PL >= 256 (mcall.c:472): host-built table in global memory, 0 beyond the double range */
        #pragma unroll
        for (int j=0; j<G; j++)
        {
            int v = pl[j];
            p[j] = (unsigned)v < 256u ? s_pl2p[v] : ((unsigned)v < (unsigned)MCB_PL2P_BIG ? tab->pl2p_big[v] : 0.0);
            if ( v > 2500 ) flags |= MCB_SITE_PL_RANGE;
        }
    }
    else
    {
        #pragma unroll
        for (int j=0; j<G; j++) p[j] = s_pl2p[pl[j]];
    }
    sum = p[0];
    #pragma unroll
    for (int j=1; j<G; j++) sum = __dadd_rn(sum, p[j]);
    return true;
}

/* ------------------------------------------------------------------------------------------------
 *  the fused kernel
 * ---------------------------------------------------------------------------------------------- */
template<int NALS, bool PLOIDY>
__global__ void __launch_bounds__(BLOCK) mcall_site_kernel(const KArgs a)
{
    using S = Shape<NALS>;
    constexpr int G = S::G, NPAIR = S::NPAIR, NTRI = S::NTRI, NSUB = S::NSUB, NACC = S::NACC;
    constexpr double LN2 = 0.693147180559945309417232121458, LN10_10 = 0.2302585092994045684017991454684;

    __shared__ Shared<NALS> sh;
    extern __shared__ __align__(128) unsigned char ring_raw[];
    int32_t *ring = reinterpret_cast<int32_t*>(ring_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nsmpl = a.nsmpl, TS = a.tile_smpl, nstage = a.nstage;
    const int tile_ints = TS*G;
    const int ntiles = (nsmpl + TS - 1)/TS;
    const bool resident = ntiles <= nstage;
    const int total_visits = resident ? ntiles : 2*ntiles;

    for (int i=tid; i<256; i+=BLOCK) sh.pl2p[i] = a.tab->pl2p[i];
    for (int i=tid; i<128; i+=BLOCK) sh.gq_thr[i] = a.tab->gq_thr[i];
    if ( tid==0 )
    {
        for (int i=0; i<nstage; i++) mbar_init(&sh.bars[i], 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t phase_bits = 0;
    const int nsites = *a.site_count;

    for (int isite = blockIdx.x; isite < nsites; isite += gridDim.x)
    {
        const int site = a.site_list[isite];
        const int32_t *site_pl = a.pl + a.pl_off[site];
        const int unseen = a.unseen ? a.unseen[site] : 0;
        const uint8_t *ploidy = nullptr;
        if ( PLOIDY )
        {
            int pid = a.ploidy_id ? a.ploidy_id[site] : 0;
            if ( pid >= a.nploidy ) pid = 0;
            ploidy = a.ploidy_tab + (size_t)pid*nsmpl;
        }

        auto issue = [&](int v)
        {
            int t = v % ntiles, stage = v % nstage;
            int n = min(TS, nsmpl - t*TS);
            uint32_t bytes = ((uint32_t)(n*G*4) + 15u) & ~15u;
            mbar_expect_tx(&sh.bars[stage], bytes);
            bulk_g2s(ring + (size_t)stage*tile_ints, site_pl + (size_t)t*tile_ints, bytes, &sh.bars[stage]);
        };
        if ( tid==0 )
            for (int v=0; v<nstage && v<total_visits; v++) issue(v);

        /* ---- site set-up by thread 0: qsum (mcall.c:1454-1464), -F prior (1507-1527), normalisation (1530-1535) */
        if ( tid==0 )
        {
            int nqs = a.nqs ? a.nqs[site] : NALS;
            float q[NALS];
            #pragma unroll
            for (int j=0; j<NALS; j++) q[j] = (a.qs && j<nqs) ? a.qs[(size_t)site*a.max_nals + j] : 0.f;
            uint32_t flags = (a.qs && nqs>0) ? 0 : MCB_SITE_NO_QS;
            if ( a.use_prior && a.prior_an && a.prior_ac )
            {
                int an = a.prior_an[site];
                if ( an!=I32_MISSING && an>0 )
                {
                    const int32_t *pac = a.prior_ac + (size_t)site*a.max_nals;
                    int ac0 = an;
                    for (int j=0; j<NALS-1; j++)
                    {
                        if ( pac[j]==I32_VEC_END ) break;
                        if ( pac[j]==I32_MISSING ) continue;
                        ac0 -= pac[j];
                        q[j+1] = (float)( __ddiv_rn(__dadd_rn((double)q[j+1], __dmul_rn(0.5,(double)pac[j])),
                                                    __dadd_rn((double)(uint32_t)nsmpl, __dmul_rn(0.5,(double)an))) );
                    }
                    if ( ac0<0 ) flags |= MCB_SITE_BAD_PRIOR;
                    q[0] = (float)( __ddiv_rn(__dadd_rn((double)q[0], __dmul_rn(0.5,(double)ac0)),
                                              __dadd_rn((double)(uint32_t)nsmpl, __dmul_rn(0.5,(double)an))) );
                }
            }
            float qsum = 0;
            #pragma unroll
            for (int j=0; j<NALS; j++) qsum = __fadd_rn(qsum, q[j]);
            if ( qsum != 0 )
            {
                #pragma unroll
                for (int j=0; j<NALS; j++) q[j] = __fdiv_rn(q[j], qsum);
            }
            #pragma unroll
            for (int j=0; j<NALS; j++) { sh.qf[j] = q[j]; sh.q[j] = (double)q[j]; }
            sh.flags = flags;
            sh.ac_packed[0] = sh.ac_packed[1] = 0;
            for (int j=0; j<8; j++) sh.ac[j] = 0;
        }
        __syncthreads();
        /* ---- allele-set coefficients, one lane per set.  float32 expression then widened (mcall.c:629-630, 671-673) */
        if ( warp==0 )
        {
            uint32_t live = 0;
            if ( lane < NPAIR )
            {
                int aa = 1; while ( aa*(aa+1)/2 <= lane ) aa++;     /* pair_idx(aa,bb)==lane */
                int bb = lane - aa*(aa-1)/2;
                float qa = sh.qf[aa], qb = sh.qf[bb];
                if ( qa!=0 && qb!=0 )
                {
                    float den = __fadd_rn(qa,qb);
                    double fa = (double)__fdiv_rn(qa,den), fb = (double)__fdiv_rn(qb,den);
                    double *cf = sh.cf_pair + lane*5;
                    cf[0] = __dmul_rn(fa,fa); cf[1] = __dmul_rn(fb,fb); cf[2] = __dmul_rn(__dmul_rn(2.0,fa),fb);
                    cf[3] = fa; cf[4] = fb;
                    live = 1u<<lane;
                }
            }
            else if ( lane-NPAIR < NTRI )
            {
                int k = lane-NPAIR;
                int aa = 2; while ( (aa+1)*aa*(aa-1)/6 <= k ) aa++;
                int r = k - aa*(aa-1)*(aa-2)/6;
                int bb = 1; while ( bb*(bb+1)/2 <= r ) bb++;
                int cc = r - bb*(bb-1)/2;
                float qa = sh.qf[aa], qb = sh.qf[bb], qc = sh.qf[cc];
                if ( qa!=0 && qb!=0 && qc!=0 )
                {
                    float den = __fadd_rn(__fadd_rn(qa,qb),qc);
                    double fa = (double)__fdiv_rn(qa,den), fb = (double)__fdiv_rn(qb,den), fc = (double)__fdiv_rn(qc,den);
                    double *cf = sh.cf_tri + k*9;
                    cf[0] = __dmul_rn(fa,fa); cf[1] = __dmul_rn(fb,fb); cf[2] = __dmul_rn(fc,fc);
                    cf[3] = __dmul_rn(__dmul_rn(2.0,fa),fb); cf[4] = __dmul_rn(__dmul_rn(2.0,fa),fc); cf[5] = __dmul_rn(__dmul_rn(2.0,fb),fc);
                    cf[6] = fa; cf[7] = fb; cf[8] = fc;
                    live = 1u<<lane;
                }
            }
            #pragma unroll
            for (int off=16; off; off>>=1) live |= __shfl_xor_sync(0xffffffffu, live, off);
            if ( lane==0 ) sh.live = live;
        }
        __syncthreads();
        const uint32_t live = sh.live;

        /* =========================== phase 1: site reduction ==================================== */
        double accM[NACC]; int accE[NACC];
        long long plsum[NALS];
        int cnt_all = 0, cnt_called = 0, since_renorm = 0;
        uint32_t tflags = 0;
        #pragma unroll
        for (int k=0; k<NACC; k++) { accM[k] = 1.0; accE[k] = 0; }
        #pragma unroll
        for (int k=0; k<NALS; k++) plsum[k] = 0;

        for (int t=0; t<ntiles; t++)
        {
            const int stage = t % nstage;
            mbar_wait(&sh.bars[stage], (phase_bits>>stage)&1u);
            phase_bits ^= 1u<<stage;
            int32_t *tile = ring + (size_t)stage*tile_ints;
            const int s0 = t*TS, n = min(TS, nsmpl - s0);
            for (int s=tid; s<n; s+=BLOCK)
            {
                int pl[G]; double p[G]; double sum;
                if ( !load_sample<NALS>(tile + s*G, unseen, sh.pl2p, a.tab, pl, p, sum, tflags) ) continue;
                /* single-allele sets: log(pdg[aa]) = -PL*ln10/10 - log(sum), every sample incl. ploidy 0 (mcall.c:607-611) */
                #pragma unroll
                for (int k=0; k<NALS; k++) plsum[k] += pl[hom_idx(k)];
                cnt_all++;
                acc_mul(accM[NACC-2], accE[NACC-2], sum);
                int pld = 2;
                if ( PLOIDY ) { pld = ploidy[s0+s]; if ( pld==0 ) continue; }      /* ploidy 0: val stays 0 (mcall.c:639-644) */
                cnt_called++;
                if ( PLOIDY ) acc_mul(accM[NACC-1], accE[NACC-1], sum);
                if ( !PLOIDY || pld==2 )
                {
                    #pragma unroll
                    for (int x=1; x<NALS; x++)
                        #pragma unroll
                        for (int y=0; y<x; y++)
                        {
                            const int k = pair_idx(x,y);
                            if ( live & (1u<<k) )
                            {
                                const double *cf = sh.cf_pair + k*5;
                                double val = fma(cf[2], p[gt_idx(x,y)], fma(cf[1], p[hom_idx(y)], cf[0]*p[hom_idx(x)]));
                                acc_mul(accM[k], accE[k], val);
                            }
                        }
                    #pragma unroll
                    for (int x=2; x<NALS; x++)
                        #pragma unroll
                        for (int y=1; y<x; y++)
                            #pragma unroll
                            for (int z=0; z<y; z++)
                            {
                                const int k = tri_idx(x,y,z);
                                if ( live & (1u<<(NPAIR+k)) )
                                {
                                    const double *cf = sh.cf_tri + k*9;
                                    double val = fma(cf[5], p[gt_idx(y,z)], fma(cf[4], p[gt_idx(x,z)], fma(cf[3], p[gt_idx(x,y)],
                                                 fma(cf[2], p[hom_idx(z)], fma(cf[1], p[hom_idx(y)], cf[0]*p[hom_idx(x)])))));
                                    acc_mul(accM[NPAIR+k], accE[NPAIR+k], val);
                                }
                            }
                }
                else    /* haploid (mcall.c:642-643, 687-688) */
                {
                    #pragma unroll
                    for (int x=1; x<NALS; x++)
                        #pragma unroll
                        for (int y=0; y<x; y++)
                        {
                            const int k = pair_idx(x,y);
                            if ( live & (1u<<k) )
                            {
                                const double *cf = sh.cf_pair + k*5;
                                double val = fma(cf[4], p[hom_idx(y)], cf[3]*p[hom_idx(x)]);
                                acc_mul(accM[k], accE[k], val);
                            }
                        }
                    #pragma unroll
                    for (int x=2; x<NALS; x++)
                        #pragma unroll
                        for (int y=1; y<x; y++)
                            #pragma unroll
                            for (int z=0; z<y; z++)
                            {
                                const int k = tri_idx(x,y,z);
                                if ( live & (1u<<(NPAIR+k)) )
                                {
                                    const double *cf = sh.cf_tri + k*9;
                                    double val = fma(cf[8], p[hom_idx(z)], fma(cf[7], p[hom_idx(y)], cf[6]*p[hom_idx(x)]));
                                    acc_mul(accM[NPAIR+k], accE[NPAIR+k], val);
                                }
                            }
                }
                if ( ++since_renorm >= 256 )
                {
                    #pragma unroll
                    for (int k=0; k<NACC; k++) acc_renorm(accM[k], accE[k]);
                    since_renorm = 0;
                }
            }
            if ( !resident )
            {
                fence_proxy_async();
                __syncthreads();
                if ( tid==0 && t+nstage < total_visits ) issue(t+nstage);
            }
        }

        /* ---- block reduction of the products (mantissa multiply, exponent add) and the integer sums */
        #pragma unroll
        for (int k=0; k<NACC; k++)
        {
            acc_renorm(accM[k], accE[k]);
            #pragma unroll
            for (int off=16; off; off>>=1)
            {
                accM[k] = __dmul_rn(accM[k], __shfl_xor_sync(0xffffffffu, accM[k], off));
                accE[k] += __shfl_xor_sync(0xffffffffu, accE[k], off);
            }
            acc_renorm(accM[k], accE[k]);
        }
        #pragma unroll
        for (int k=0; k<NALS; k++)
            #pragma unroll
            for (int off=16; off; off>>=1) plsum[k] += __shfl_xor_sync(0xffffffffu, plsum[k], off);
        #pragma unroll
        for (int off=16; off; off>>=1)
        {
            cnt_all    += __shfl_xor_sync(0xffffffffu, cnt_all, off);
            cnt_called += __shfl_xor_sync(0xffffffffu, cnt_called, off);
            tflags     |= __shfl_xor_sync(0xffffffffu, tflags, off);
        }
        if ( lane==0 )
        {
            #pragma unroll
            for (int k=0; k<NACC; k++) { sh.red_M[warp][k] = accM[k]; sh.red_E[warp][k] = accE[k]; }
            #pragma unroll
            for (int k=0; k<NALS; k++) sh.red_pls[warp][k] = plsum[k];
            sh.red_cnt[warp][0] = cnt_all; sh.red_cnt[warp][1] = cnt_called;
            if ( tflags ) atomicOr(&sh.flags, tflags);
        }
        __syncthreads();

        /* ---- epilogue, warp 0: lane k <-> allele set k in the reference's enumeration order ------------- */
        if ( warp==0 )
        {
            int n_all = 0, n_called = 0;
            #pragma unroll
            for (int w=0; w<NWARP; w++) { n_all += sh.red_cnt[w][0]; n_called += sh.red_cnt[w][1]; }
            if ( !PLOIDY ) n_called = n_all;
            auto total_log = [&](int k, int n) -> double
            {
                double M = 1.0; int E = 0;
                #pragma unroll
                for (int w=0; w<NWARP; w++) { M = __dmul_rn(M, sh.red_M[w][k]); E += sh.red_E[w][k]; }
                /* |E| < 2^31: at most 2^20 samples x 2047 */
                return log(M) + (double)(E - 1023*n)*LN2;
            };
            const double lnN_all    = n_all ? total_log(NACC-2, n_all) : 0.0;
            const double lnN_called = PLOIDY ? (n_called ? total_log(NACC-1, n_called) : 0.0) : lnN_all;

            double lk = 0; bool cand = false, in_sum = false; uint32_t mask = 0;
            if ( lane < NALS )
            {
                long long ps = 0;
                #pragma unroll
                for (int w=0; w<NWARP; w++) ps += sh.red_pls[w][lane];
                bool set = n_all > 0;
                lk = set ? -LN10_10*(double)ps - lnN_all : 0.0;
                if ( lane>0 ) lk += a.theta;
                cand = set; in_sum = set && lane>0; mask = 1u<<lane;
            }
            else if ( lane < NSUB )
            {
                int k = lane - NALS;                    /* accumulator index: pairs then triples */
                bool lv = (live >> k) & 1u;
                bool set = lv && n_called > 0;
                int nonref = 0;
                if ( k < NPAIR )
                {
                    int aa = 1; while ( aa*(aa+1)/2 <= k ) aa++;
                    int bb = k - aa*(aa-1)/2;
                    mask = 1u<<aa | 1u<<bb; nonref = (aa!=0) + (bb!=0);
                }
                else
                {
                    int kk = k - NPAIR;
                    int aa = 2; while ( (aa+1)*aa*(aa-1)/6 <= kk ) aa++;
                    int r = kk - aa*(aa-1)*(aa-2)/6;
                    int bb = 1; while ( bb*(bb+1)/2 <= r ) bb++;
                    int cc = r - bb*(bb-1)/2;
                    mask = 1u<<aa | 1u<<bb | 1u<<cc; nonref = (aa!=0) + (bb!=0) + (cc!=0);
                }
                lk = set ? total_log(k, n_called) - lnN_called : 0.0;
                for (int j=0; j<nonref; j++) lk += a.theta;
                cand = set; in_sum = set;
            }
            /* first strict maximum in enumeration order (UPDATE_MAX_LKs, mcall.c:582-585) */
            double best = cand ? lk : -CUDART_INF; int best_lane = cand ? lane : 64;
            #pragma unroll
            for (int off=16; off; off>>=1)
            {
                double ob = __shfl_xor_sync(0xffffffffu, best, off);
                int    ol = __shfl_xor_sync(0xffffffffu, best_lane, off);
                if ( ob > best || (ob==best && ol < best_lane) ) { best = ob; best_lane = ol; }
            }
            double second = (cand && lane!=best_lane) ? lk : -CUDART_INF;
            #pragma unroll
            for (int off=16; off; off>>=1) second = fmax(second, __shfl_xor_sync(0xffffffffu, second, off));
            /* lk_sum = log sum exp over every evaluated set except {REF} (mcall.c:584, 614) */
            double mx = in_sum ? lk : -CUDART_INF;
            #pragma unroll
            for (int off=16; off; off>>=1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            double term = in_sum ? exp(lk - mx) : 0.0;
            #pragma unroll
            for (int off=16; off; off>>=1) term += __shfl_xor_sync(0xffffffffu, term, off);
            const double grp_lk_sum = mx > -CUDART_INF ? mx + log(term) : -CUDART_INF;
            const double grp_ref_lk = __shfl_sync(0xffffffffu, lk, 0);
            const uint32_t grp_als = __shfl_sync(0xffffffffu, mask, best_lane & 31);

            if ( lane==0 )
            {
                const bool any = best_lane < 64;
                uint32_t gals = any ? grp_als : 0;
                uint32_t flags = sh.flags;
                double max_qual = -CUDART_INF, lk_sum = -CUDART_INF, ref_lk = -CUDART_INF;
                if ( any )          /* mcall.c:1553-1560 */
                {
                    max_qual = -4.343*(grp_ref_lk - logsumexp2_dev(grp_lk_sum, grp_ref_lk));
                    lk_sum = grp_lk_sum; ref_lk = grp_ref_lk;
                }
                double gap = any ? best - second : CUDART_INF;
                if ( any && gap < a.tie_eps ) flags |= MCB_SITE_NEAR_TIE;
                uint32_t als_new = gals | 1u;               /* mcall.c:1552, 1564 */
                int is_variant = als_new!=1;
                int ret_early = ((a.flag & MCB_CALL_VARONLY) && !is_variant) || (flags & MCB_SITE_NO_QS);
                int nals_new = 0;
                #pragma unroll
                for (int j=0; j<NALS; j++)                  /* mcall.c:1569-1575 */
                {
                    if ( j>0 && j==unseen ) continue;
                    if ( a.flag & MCB_CALL_KEEPALT ) als_new |= 1u<<j;
                    if ( als_new & (1u<<j) ) nals_new++;
                }
                int nout = 0, kk = 0, l = 0;                /* mcall.c:547-570 */
                #pragma unroll
                for (int x=0; x<NALS; x++) sh.als_map[x] = (als_new & (1u<<x)) ? nout++ : -1;
                #pragma unroll
                for (int x=0; x<NALS; x++)
                    #pragma unroll
                    for (int y=0; y<=x; y++) { if ( (als_new & (1u<<x)) && (als_new & (1u<<y)) ) sh.pl_map[kk++] = l; l++; }
                for (; kk<G; kk++) sh.pl_map[kk] = 0;
                if ( unseen && (als_new & (1u<<unseen)) ) flags |= MCB_SITE_UNSEEN_SEL;
                sh.pl_dropped = als_new==1;
                sh.ref_gt = (als_new==1) || !is_variant;
                if ( sh.pl_dropped ) flags |= MCB_SITE_PL_DROPPED;
                if ( sh.ref_gt ) flags |= MCB_SITE_REF_GT;
                int gn = 0;
                #pragma unroll
                for (int j=0; j<NALS; j++) gn += (gals>>j)&1u;
                sh.grp_als = gals; sh.grp_nals = gn; sh.als_new = als_new; sh.nals_new = nals_new;
                sh.is_variant = is_variant; sh.ret_early = ret_early; sh.flags = flags;
                sh.max_qual = max_qual; sh.lk_sum = lk_sum; sh.ref_lk = ref_lk; sh.gap = gap;
            }
        }
        __syncthreads();

        /* =========================== phase 2: per-sample genotypes ============================== */
        if ( sh.ret_early )
        {
            if ( tid==0 )
            {
                a.ret[site] = 0;
                if ( a.site_flags ) a.site_flags[site] = sh.flags;
            }
            /* drain the phase-2 tiles already in flight for this site (streaming mode): they hit L2 */
            if ( !resident )
                for (int v=ntiles; v<ntiles+nstage && v<total_visits; v++)
                {
                    const int stage = v % nstage;
                    mbar_wait(&sh.bars[stage], (phase_bits>>stage)&1u);
                    phase_bits ^= 1u<<stage;
                }
            __syncthreads();
            continue;
        }
        {
            const uint32_t gals = sh.grp_als;
            const int nals_new = sh.nals_new, ngt_new = nals_new*(nals_new+1)/2, grp_nals = sh.grp_nals;
            const bool ref_gt = sh.ref_gt, pl_dropped = sh.pl_dropped;
            const bool want_gq = a.gq && (a.output_tags & (MCB_CALL_FMT_GQ|MCB_CALL_FMT_GP));
            int32_t *out_pl = (a.out_pl && !pl_dropped) ? a.out_pl + a.pl_off[site] : nullptr;
            int2 *out_gt = a.gt ? reinterpret_cast<int2*>(a.gt) + (size_t)site*nsmpl : nullptr;
            int32_t *out_gq = want_gq ? a.gq + (size_t)site*nsmpl : nullptr;
            unsigned long long ac_lo = 0, ac_hi = 0;
            uint32_t tflags2 = 0;

            for (int t=0; t<ntiles; t++)
            {
                const int v = resident ? t : ntiles + t, stage = v % nstage;
                if ( !resident )
                {
                    mbar_wait(&sh.bars[stage], (phase_bits>>stage)&1u);
                    phase_bits ^= 1u<<stage;
                }
                int32_t *tile = ring + (size_t)stage*tile_ints;
                const int s0 = t*TS, n = min(TS, nsmpl - s0);
                for (int s=tid; s<n; s+=BLOCK)
                {
                    int pl[G]; double p[G]; double sum = 1;
                    int32_t *row = tile + s*G;
                    const bool has = load_sample<NALS>(row, unseen, sh.pl2p, a.tab, pl, p, sum, tflags2);
                    const int pld = PLOIDY ? ploidy[s0+s] : 2;
                    int gt0, gt1, gq = 0;
                    if ( !pld ) { gt0 = MCB_GT_MISSING; gt1 = I32_VEC_END; }
                    else if ( !has ) { gt0 = MCB_GT_MISSING; gt1 = pld==2 ? MCB_GT_MISSING : I32_VEC_END; }
                    else if ( ref_gt )          /* mcall.c:713-743 */
                    {
                        gt0 = MCB_GT_UNPHASED(0); gt1 = pld==2 ? MCB_GT_UNPHASED(0) : I32_VEC_END;
                        ac_lo += (unsigned long long)pld;
                    }
                    else                        /* mcall.c:787-840, literal */
                    {
                        double best = 0; int g0 = 0, g1 = 0;
                        double ghom[NALS], ghet[NPAIR ? NPAIR : 1];
                        #pragma unroll
                        for (int x=0; x<NALS; x++)
                        {
                            ghom[x] = 0;
                            if ( gals & (1u<<x) )
                            {
                                const double pdg = __ddiv_rn(p[hom_idx(x)], sum);
                                const double lk = pld==2 ? __dmul_rn(__dmul_rn(pdg, sh.q[x]), sh.q[x]) : __dmul_rn(pdg, sh.q[x]);
                                ghom[x] = (double)__double2float_rn(lk);
                                if ( best < lk ) { best = lk; g0 = sh.als_map[x]; }
                            }
                        }
                        if ( pld==2 )
                        {
                            g1 = g0;
                            #pragma unroll
                            for (int x=1; x<NALS; x++)
                                #pragma unroll
                                for (int y=0; y<x; y++)
                                {
                                    ghet[pair_idx(x,y)] = 0;
                                    if ( (gals & (1u<<x)) && (gals & (1u<<y)) )
                                    {
                                        const double pdg = __ddiv_rn(p[gt_idx(x,y)], sum);
                                        const double lk = __dmul_rn(__dmul_rn(__dmul_rn(2.0,pdg), sh.q[x]), sh.q[y]);
                                        ghet[pair_idx(x,y)] = (double)__double2float_rn(lk);
                                        if ( best < lk ) { best = lk; g0 = sh.als_map[y]; g1 = sh.als_map[x]; }
                                    }
                                }
                            gt0 = MCB_GT_UNPHASED(g0); gt1 = MCB_GT_UNPHASED(g1);
                            if ( g0<4 ) ac_lo += 1ull<<(16*g0); else ac_hi += 1ull<<(16*(g0-4));
                            if ( g1<4 ) ac_lo += 1ull<<(16*g1); else ac_hi += 1ull<<(16*(g1-4));
                        }
                        else
                        {
                            gt0 = MCB_GT_UNPHASED(g0); gt1 = I32_VEC_END;
                            if ( g0<4 ) ac_lo += 1ull<<(16*g0); else ac_hi += 1ull<<(16*(g0-4));
                        }
                        if ( want_gq )          /* mcall.c:843-878: max and sum over gps[0..nmax) in index order */
                        {
                            double gmax = 0, gsum = 0;
                            if ( pld==2 )
                            {
                                #pragma unroll
                                for (int x=0; x<NALS; x++)
                                {
                                    if ( !(gals & (1u<<x)) ) continue;
                                    #pragma unroll
                                    for (int y=0; y<x; y++)
                                    {
                                        if ( !(gals & (1u<<y)) ) continue;
                                        if ( gt_idx(sh.als_map[x], sh.als_map[y]) < ngt_new )
                                        {
                                            const double g = ghet[pair_idx(x,y)];
                                            if ( gmax < g ) gmax = g;
                                            gsum = __dadd_rn(gsum, g);
                                        }
                                    }
                                    if ( hom_idx(sh.als_map[x]) < ngt_new )
                                    {
                                        const double g = ghom[x];
                                        if ( gmax < g ) gmax = g;
                                        gsum = __dadd_rn(gsum, g);
                                    }
                                }
                            }
                            else
                            {
                                #pragma unroll
                                for (int x=0; x<NALS; x++)
                                    if ( (gals & (1u<<x)) && sh.als_map[x] < grp_nals )
                                    {
                                        const double g = ghom[x];
                                        if ( gmax < g ) gmax = g;
                                        gsum = __dadd_rn(gsum, g);
                                    }
                            }
                            const double xx = __dadd_rn(1.0, -__ddiv_rn(gmax, gsum));
                            if ( !(xx==xx) ) gq = 127;      /* NaN (0/0): `max<=INT8_MAX` is false => INT8_MAX */
                            else
                            {
                                /* (int)(-4.34294*log(x)) from host-libm thresholds; float estimate, then exact fix-up */
                                int k = xx > 0 ? (int)(-3.0102999f*__log2f((float)xx)) : 127;
                                k = max(0, min(127, k));
                                while ( k<127 && xx <= sh.gq_thr[k+1] ) k++;
                                while ( k>0 && xx > sh.gq_thr[k] ) k--;
                                gq = k;
                            }
                        }
                    }
                    if ( out_gt ) out_gt[s0+s] = make_int2(gt0, gt1);
                    if ( out_gq ) out_gq[s0+s] = gq;
                    if ( out_pl )               /* mcall.c:1158-1194; `row` holds the filled PLs */
                    {
                        int32_t *dst = out_pl + (size_t)(s0+s)*ngt_new;
                        if ( pld==2 )
                        {
                            #pragma unroll
                            for (int k=0; k<G; k++) if ( k<ngt_new ) dst[k] = row[sh.pl_map[k]];
                        }
                        else if ( pld==1 )
                        {
                            #pragma unroll
                            for (int k=0; k<G; k++)
                                if ( k<ngt_new ) dst[k] = k<nals_new ? row[sh.pl_map[hom_idx(k)]] : I32_VEC_END;
                        }
                        else
                        {
                            #pragma unroll
                            for (int k=0; k<G; k++) if ( k<ngt_new ) dst[k] = k==0 ? I32_MISSING : I32_VEC_END;
                        }
                    }
                }
                if ( !resident )
                {
                    fence_proxy_async();
                    __syncthreads();
                    if ( tid==0 && v+nstage < total_visits ) issue(v+nstage);
                }
            }
            /* ---- AC: packed 16-bit counters per thread -> shared int counters (mcall.c:839-840) */
            #pragma unroll
            for (int j=0; j<4; j++)
            {
                int c0 = (int)((ac_lo >> (16*j)) & 0xffff), c1 = (int)((ac_hi >> (16*j)) & 0xffff);
                #pragma unroll
                for (int off=16; off; off>>=1) { c0 += __shfl_xor_sync(0xffffffffu, c0, off); c1 += __shfl_xor_sync(0xffffffffu, c1, off); }
                if ( lane==0 ) { if ( c0 ) atomicAdd(&sh.ac[j], c0); if ( c1 ) atomicAdd(&sh.ac[4+j], c1); }
            }
            if ( tflags2 ) atomicOr(&sh.flags, tflags2);
        }
        __syncthreads();

        /* ---- site record: QUAL (mcall.c:1631-1645), AC/AN (1648-1650) ---------------------------------- */
        if ( tid==0 )
        {
            int nAC = 0;
            if ( !sh.ref_gt ) for (int j=1; j<sh.nals_new && j<8; j++) nAC += sh.ac[j];
            int ret = sh.nals_new;
            if ( !sh.ref_gt && !nAC && (a.flag & MCB_CALL_VARONLY) ) ret = 0;      /* mcall.c:1618 */
            float qual;
            if ( nAC ) qual = (float)sh.max_qual;
            else if ( sh.lk_sum != -CUDART_INF ) qual = (float)(-4.343*(sh.lk_sum - logsumexp2_dev(sh.lk_sum, sh.ref_lk)));
            else if ( sh.ac[0] ) qual = a.theta ? (float)(-4.343*a.theta) : 0.f;
            else qual = __uint_as_float(MCB_FLOAT_MISSING_BITS);
            a.ret[site] = ret;
            if ( a.als_new ) a.als_new[site] = sh.als_new;
            if ( a.als_map ) for (int j=0; j<a.max_nals; j++) a.als_map[(size_t)site*a.max_nals + j] = j<NALS ? (int8_t)sh.als_map[j] : (int8_t)-1;
            if ( a.qual ) a.qual[site] = qual;
            if ( a.ac ) for (int j=0; j<a.max_nals; j++) a.ac[(size_t)site*a.max_nals + j] = (j<sh.nals_new && j<8) ? sh.ac[j] : 0;
            if ( a.an ) a.an[site] = nAC + sh.ac[0];
            if ( a.site_flags ) a.site_flags[site] = sh.flags;
            if ( a.diag ) { double *d = a.diag + (size_t)site*4; d[0] = sh.max_qual; d[1] = sh.lk_sum; d[2] = sh.ref_lk; d[3] = sh.gap; }
        }
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------------------------
 *  site classification: one list of site indices per allele count (1..5), one for everything else
 * ---------------------------------------------------------------------------------------------- */
__global__ void classify_sites_kernel(const uint8_t *nals, int nsites, int32_t *lists, int32_t *counts, int list_stride)
{
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if ( i >= nsites ) return;
    int n = nals[i];
    int cls = (n>=1 && n<=5) ? n : 0;
    int pos = atomicAdd(&counts[cls], 1);
    lists[(size_t)cls*list_stride + pos] = i;
}

/*  sites the templated kernels do not cover (n_allele 0 or >5): reported as skipped for now  */
__global__ void unsupported_sites_kernel(const int32_t *list, const int32_t *count, int32_t *ret, uint32_t *site_flags, const uint8_t *nals)
{
    int n = *count;
    for (int i = blockIdx.x*blockDim.x + threadIdx.x; i<n; i += gridDim.x*blockDim.x)
    {
        int site = list[i];
        ret[site] = 0;
        if ( site_flags ) site_flags[site] = nals[site] > 32 ? MCB_SITE_TOO_MANY_ALS : MCB_SITE_UNSUPPORTED;
    }
}

template<int NALS, bool PLOIDY>
static cudaError_t launch_one(const KArgs &a, int grid, size_t ring_bytes, cudaStream_t st)
{
    auto kern = mcall_site_kernel<NALS,PLOIDY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes);
    if ( e!=cudaSuccess ) return e;
    kern<<<grid, BLOCK, ring_bytes, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_site_kernel(int nals, bool ploidy, const KArgs &a, int grid, size_t ring_bytes, cudaStream_t st)
{
    switch ( nals*2 + (ploidy?1:0) )
    {
        case 2:  return launch_one<1,false>(a,grid,ring_bytes,st);
        case 3:  return launch_one<1,true >(a,grid,ring_bytes,st);
        case 4:  return launch_one<2,false>(a,grid,ring_bytes,st);
        case 5:  return launch_one<2,true >(a,grid,ring_bytes,st);
        case 6:  return launch_one<3,false>(a,grid,ring_bytes,st);
        case 7:  return launch_one<3,true >(a,grid,ring_bytes,st);
        case 8:  return launch_one<4,false>(a,grid,ring_bytes,st);
        case 9:  return launch_one<4,true >(a,grid,ring_bytes,st);
        case 10: return launch_one<5,false>(a,grid,ring_bytes,st);
        case 11: return launch_one<5,true >(a,grid,ring_bytes,st);
    }
    return cudaErrorInvalidValue;
}

template<int NALS, bool PLOIDY>
static cudaError_t occ_one(size_t ring_bytes, int *nb)
{
    auto kern = mcall_site_kernel<NALS,PLOIDY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes);
    if ( e!=cudaSuccess ) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, kern, BLOCK, ring_bytes);
}
cudaError_t site_kernel_occupancy(int nals, bool ploidy, size_t ring_bytes, int *nb)
{
    switch ( nals*2 + (ploidy?1:0) )
    {
        case 2:  return occ_one<1,false>(ring_bytes,nb);
        case 3:  return occ_one<1,true >(ring_bytes,nb);
        case 4:  return occ_one<2,false>(ring_bytes,nb);
        case 5:  return occ_one<2,true >(ring_bytes,nb);
        case 6:  return occ_one<3,false>(ring_bytes,nb);
        case 7:  return occ_one<3,true >(ring_bytes,nb);
        case 8:  return occ_one<4,false>(ring_bytes,nb);
        case 9:  return occ_one<4,true >(ring_bytes,nb);
        case 10: return occ_one<5,false>(ring_bytes,nb);
        case 11: return occ_one<5,true >(ring_bytes,nb);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_classify(const uint8_t *nals, int nsites, int32_t *lists, int32_t *counts, int list_stride, cudaStream_t st)
{
    if ( nsites<=0 ) return cudaSuccess;
    classify_sites_kernel<<<(nsites+255)/256, 256, 0, st>>>(nals, nsites, lists, counts, list_stride);
    return cudaGetLastError();
}
cudaError_t launch_unsupported(const int32_t *list, const int32_t *count, int32_t *ret, uint32_t *site_flags, const uint8_t *nals, cudaStream_t st)
{
    unsupported_sites_kernel<<<64, 256, 0, st>>>(list, count, ret, site_flags, nals);
    return cudaGetLastError();
}

}   // namespace mcb
