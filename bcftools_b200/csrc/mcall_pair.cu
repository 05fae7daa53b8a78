/*  mcall_pair.cu -- phase 2 of the "pair sites" of the 3-5 allele classes, one CTA per site.
 *
 *  A pair site is a multi-allelic site whose selected allele set (mcall_find_best_alleles, mcall.c:591-710) is a pair
 *  s0<s1 and whose kept alleles are exactly those two: 61 % of the 3-allele, 46 % of the 4-allele and 38 % of the
 *  5-allele sites of the C3 mix.  The fused kernel (mcall_kernels.cu) runs phase 1 for them, writes a PairRec and
 *  appends the site to the class's pair list; this kernel then does mcall_call_genotypes + GQ (mcall.c:745-886),
 *  mcall_trim_and_update_PLs (mcall.c:1158-1194) and the site epilogue (mcall.c:1631-1650) for the listed sites.
 *
 *  Why a second kernel: the per-sample work of a pair site is the same straight-line code as the two-allele kernel's
 *  (fast2_call), but inside the fused kernel it competes with the general phase-2 code for instruction fetch and made
 *  every class slower.  Here every warp runs the same ~300-instruction loop (one CTA of 8 warps per site: with one
 *  WARP per site the few thousand pair sites of a batch leave most of the GPU idle behind 120-us warps).  The PL block is read a second time --
 *  straight from global memory with 128-bit loads, two adjacent samples per lane, no shared-memory staging; it was read
 *  by the fused kernel moments earlier, so most of it is still in the 126 MB L2 -- which these classes can afford: their
 *  kernels sit at 14-32 % of the DRAM bandwidth.
 *
 *  Samples with a missing / vector_end value or a PL >= 256 take the literal general path per lane (set_pdg's fill of
 *  mcall.c:495-527 on a private copy, the host-built big-PL table, IEEE division, bw_call_sample<false>).
 */

#include "mcall_device.cuh"

namespace mcb {

#define PW_WARPS   8
#define PW_MINCTA  3

struct PWTables
{
    double pl2p[256];
    double gq_thr[130];
    int4   slot[4];                 /* diploid call of slot k for the general path: {gt0, gt1, AC[0] inc, AC[1] inc} */
    int    next_site, alt, called;
    uint32_t tflags;
};

/*  set_pdg's missing-value fill (mcall.c:495-527) on a private copy of one sample's PLs; returns 0 for "no data"  */
template<int NALS>
__device__ __noinline__ int pw_fix_missing(int *pl, int unseen)
{
    constexpr int G = NALS*(NALS+1)/2;
    int j;
    for (j=0; j<G; j++)
    {
        if ( pl[j]==I32_VEC_END ) return 0;         /* not diploid-shaped: all missing, mcall.c:465-470 */
        if ( pl[j]==I32_MISSING ) break;
    }
    if ( j==0 || j==G ) return 0;                   /* first value missing (mcall.c:476-481) / negative garbage that is no sentinel */
    j = 0;
    for (int ia=0; ia<NALS; ia++)
        for (int ib=0; ib<=ia; ib++)
        {
            if ( pl[j]==I32_MISSING )
            {
                int k = gt_idx(ia,unseen);
                if ( pl[k]==I32_MISSING ) k = gt_idx(ib,unseen);
                if ( pl[k]==I32_MISSING ) k = gt_idx(unseen,unseen);
                pl[j] = pl[k]==I32_MISSING ? 255 : pl[k];
            }
            else if ( pl[j] < 0 ) return 0;
            j++;
        }
    return 1;
}

/*  one sample through the general path: fills out[] = {gt0, gt1, gq, PL'0, PL'1, PL'2, alt copies, called}  */
template<int NALS>
__device__ __noinline__ void pw_slow_sample(const int32_t *row, const PairRec *rec, int unseen, uint32_t pl2p_s, uint32_t thr_s, uint32_t slot_s,
                                            const DevTables *tab, uint32_t *flags, int *out)
{
    constexpr int G = NALS*(NALS+1)/2;
    int pl[G]; double p[G];
    int orv = 0;
    for (int j=0; j<G; j++) { pl[j] = row[j]; orv |= pl[j]; }
    bool data = orv != 0;                           /* PL=0,..,0: no data (mcall.c:529-537) */
    if ( orv < 0 )
    {
        data = pw_fix_missing<NALS>(pl, unseen);
        if ( data ) { orv = 0; for (int j=0; j<G; j++) orv |= pl[j]; data = orv > 0; }
    }
    for (int j=0; j<G; j++)
    {
        double v = 1.0;
        if ( data )
        {
            if ( (unsigned)pl[j] < 256u ) v = lds64(pl2p_s + 8u*(uint32_t)pl[j]);
            else { if ( pl[j] > 2500 ) *flags |= MCB_SITE_PL_RANGE; v = (unsigned)pl[j] < (unsigned)MCB_PL2P_BIG ? tab->pl2p_big[pl[j]] : 0.0; }
        }
        p[j] = v;
    }
    double sum = p[0];
    for (int j=1; j<G; j++) sum = __dadd_rn(sum, p[j]);
    out[3] = pl[rec->g00]; out[4] = pl[rec->g10]; out[5] = pl[rec->g11];        /* the filled values are what gets output */
    out[0] = MCB_GT_MISSING; out[1] = MCB_GT_MISSING; out[2] = 0; out[6] = 0; out[7] = 0;
    if ( !data ) return;
    BWConsts c;
    c.q0 = rec->q0; c.q1 = rec->q1; c.slot_s = slot_s; c.thr_s = thr_s; c.nsel = 2; c.jgt0 = 0; c.inc_dip = 7; c.want_gq = true;
    int gq = 0;
    const int4 o = bw_call_sample<false>(p[rec->g00], p[rec->g10], p[rec->g11], sum, c, gq);
    out[0] = o.x; out[1] = o.y; out[2] = gq; out[6] = o.w; out[7] = 1;
}

template<int NALS>
__global__ void __launch_bounds__(PW_WARPS*32, PW_MINCTA) mcall_pair_phase2_kernel(const KArgs a)
{
    constexpr int G = NALS*(NALS+1)/2;
    PWTables &tb = *reinterpret_cast<PWTables*>(mcb_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sbase = smem_base();
    const uint32_t pl2p_s = sbase + (uint32_t)offsetof(PWTables, pl2p), thr_s = sbase + (uint32_t)offsetof(PWTables, gq_thr);
    const uint32_t slot_s = sbase + (uint32_t)offsetof(PWTables, slot);
    for (int i=tid; i<256; i+=blockDim.x) tb.pl2p[i] = a.tab->pl2p[i];
    for (int i=tid; i<130; i+=blockDim.x) tb.gq_thr[i] = i<128 ? a.tab->gq_thr[i] : -1.0;
    if ( tid < 4 )          /* new alleles 0 and 1 (mcall.c:830-831): slot k has k copies of allele 1; [3] unused */
        tb.slot[tid] = make_int4(MCB_GT_UNPHASED(tid==2 ? 1 : 0), MCB_GT_UNPHASED(tid ? 1 : 0), 2 - tid, tid);

    const int S = a.nsmpl, npair = S >> 1;
    const int nsites = *a.pair_count;
    for (;;)
    {
        if ( tid==0 ) { tb.next_site = atomicAdd(a.pair_work, 1); tb.alt = 0; tb.called = 0; tb.tflags = 0; }
        __syncthreads();
        const int isite = tb.next_site;
        if ( isite >= nsites ) break;
        const int site = a.pair_list[isite];
        const PairRec *rec = a.pair_rec + site;
        const int32_t *site_pl = reinterpret_cast<const int32_t*>(a.pl) + a.pl_off[site];
        const int unseen = a.unseen ? a.unseen[site] : 0;
        const double q0 = rec->q0, q1 = rec->q1, q1x2 = __dmul_rn(2.0, q1);
        const int g00 = rec->g00, g10 = rec->g10, g11 = rec->g11;
        int32_t *out_pl = a.out_pl + rec->out_off;
        int32_t *out_gt = a.gt + 2*(size_t)site*S;
        int32_t *out_gq = a.gq + (size_t)site*S;
        int alt = 0, called = 0;
        uint32_t tflags = 0;

        #pragma unroll 1
        for (int pr=tid; pr<npair; pr+=PW_WARPS*32)
        {
            const int32_t *row = site_pl + (size_t)(2*G)*pr;
            int pl2[2*G];
            if constexpr ( (G & 1)==0 )
            {
                #pragma unroll
                for (int j=0; j<2*G; j+=4) { const int4 v = __ldg(reinterpret_cast<const int4*>(row + j)); pl2[j] = v.x; pl2[j+1] = v.y; pl2[j+2] = v.z; pl2[j+3] = v.w; }
            }
            else
            {
                #pragma unroll
                for (int j=0; j<2*G; j+=2) { const int2 v = __ldg(reinterpret_cast<const int2*>(row + j)); pl2[j] = v.x; pl2[j+1] = v.y; }
            }
            int orv0 = 0, orv1 = 0;
            #pragma unroll
            for (int j=0; j<G; j++) { orv0 |= pl2[j]; orv1 |= pl2[G+j]; }
            int x0, y0, x1, y1, gq0, gq1, a0, b0, c0, a1, b1, c1;
            if ( (unsigned)(orv0 | orv1) <= 255u )
            {
                /* the three genotypes of the pair: site-uniform indices, re-read (the lines are in L1) */
                a0 = __ldg(row + g00); b0 = __ldg(row + g10); c0 = __ldg(row + g11);
                a1 = __ldg(row + G + g00); b1 = __ldg(row + G + g10); c1 = __ldg(row + G + g11);
                int k0, k1;
                {
                    double sum = lds64c(pl2p_s + 8u*(uint32_t)pl2[0]);
                    #pragma unroll
                    for (int j=1; j<G; j++) sum = __dadd_rn(sum, lds64c(pl2p_s + 8u*(uint32_t)pl2[j]));
                    fast2_call(lds64c(pl2p_s + 8u*(uint32_t)a0), lds64c(pl2p_s + 8u*(uint32_t)b0), lds64c(pl2p_s + 8u*(uint32_t)c0),
                               sum, q0, q1, q1x2, thr_s, k0, gq0);
                }
                {
                    double sum = lds64c(pl2p_s + 8u*(uint32_t)pl2[G]);
                    #pragma unroll
                    for (int j=1; j<G; j++) sum = __dadd_rn(sum, lds64c(pl2p_s + 8u*(uint32_t)pl2[G+j]));
                    fast2_call(lds64c(pl2p_s + 8u*(uint32_t)a1), lds64c(pl2p_s + 8u*(uint32_t)b1), lds64c(pl2p_s + 8u*(uint32_t)c1),
                               sum, q0, q1, q1x2, thr_s, k1, gq1);
                }
                const bool has0 = orv0 != 0, has1 = orv1 != 0;      /* PL=0,..,0: no data (mcall.c:529-537) */
                /* new alleles 0 and 1: GT codes 2 and 4, slot index = copies of allele 1; no data: ./. and GQ 0 */
                x0 = has0 ? (k0==2 ? 4 : 2) : 0; y0 = has0 ? (k0 ? 4 : 2) : 0; gq0 = has0 ? gq0 : 0;
                x1 = has1 ? (k1==2 ? 4 : 2) : 0; y1 = has1 ? (k1 ? 4 : 2) : 0; gq1 = has1 ? gq1 : 0;
                alt += (has0 ? k0 : 0) + (has1 ? k1 : 0); called += (int)has0 + (int)has1;
            }
            else
            {
                int o[8];
                pw_slow_sample<NALS>(row, rec, unseen, pl2p_s, thr_s, slot_s, a.tab, &tflags, o);
                x0 = o[0]; y0 = o[1]; gq0 = o[2]; a0 = o[3]; b0 = o[4]; c0 = o[5]; alt += o[6]; called += o[7];
                pw_slow_sample<NALS>(row + G, rec, unseen, pl2p_s, thr_s, slot_s, a.tab, &tflags, o);
                x1 = o[0]; y1 = o[1]; gq1 = o[2]; a1 = o[3]; b1 = o[4]; c1 = o[5]; alt += o[6]; called += o[7];
            }
            asm volatile("st.global.v4.s32 [%0], {%1,%2,%3,%4};" :: "l"(out_gt + 4*(size_t)pr), "r"(x0), "r"(y0), "r"(x1), "r"(y1) : "memory");
            asm volatile("st.global.v2.s32 [%0], {%1,%2};" :: "l"(out_gq + 2*(size_t)pr), "r"(gq0), "r"(gq1) : "memory");
            const int32_t *dst = out_pl + 6*(size_t)pr;             /* mcall.c:1158-1194: the kept genotypes are the three of the pair */
            asm volatile("st.global.v2.s32 [%0], {%1,%2};" :: "l"(dst), "r"(a0), "r"(b0) : "memory");
            asm volatile("st.global.v2.s32 [%0+8], {%1,%2};" :: "l"(dst), "r"(c0), "r"(a1) : "memory");
            asm volatile("st.global.v2.s32 [%0+16], {%1,%2};" :: "l"(dst), "r"(b1), "r"(c1) : "memory");
        }
        #pragma unroll
        for (int off=16; off; off>>=1)
        {
            alt += __shfl_xor_sync(0xffffffffu, alt, off); called += __shfl_xor_sync(0xffffffffu, called, off);
            tflags |= __shfl_xor_sync(0xffffffffu, tflags, off);
        }
        if ( lane==0 ) { atomicAdd(&tb.alt, alt); atomicAdd(&tb.called, called); if ( tflags ) atomicOr(&tb.tflags, tflags); }
        __syncthreads();
        /* ---- site record: QUAL (mcall.c:1631-1645), AC/AN (1648-1650) */
        if ( tid==0 )
        {
            alt = tb.alt; called = tb.called; tflags = tb.tflags;
            const int ac0 = 2*called - alt, nAC = alt;
            int ret = rec->nals_new;
            if ( !nAC && (a.flag & MCB_CALL_VARONLY) ) ret = 0;      /* mcall.c:1618 */
            float qual;
            if ( nAC ) qual = (float)rec->max_qual;
            else if ( rec->lk_sum != -CUDART_INF ) qual = (float)(-4.343*(rec->lk_sum - logsumexp2_dev(rec->lk_sum, rec->ref_lk)));
            else if ( ac0 ) qual = a.theta ? (float)(-4.343*a.theta) : 0.f;
            else qual = __uint_as_float(MCB_FLOAT_MISSING_BITS);
            a.ret[site] = ret;
            if ( a.als_new ) a.als_new[site] = rec->als_new;
            if ( a.als_map ) for (int j=0; j<a.max_nals; j++) a.als_map[(size_t)site*a.max_nals + j] = j<8 ? rec->als_map[j] : (int8_t)-1;
            if ( a.qual ) a.qual[site] = qual;
            if ( a.ac ) for (int j=0; j<a.max_nals; j++) a.ac[(size_t)site*a.max_nals + j] = j==0 ? ac0 : (j==1 ? alt : 0);
            if ( a.an ) a.an[site] = nAC + ac0;
            if ( a.site_flags ) a.site_flags[site] = rec->flags | tflags;
            if ( a.diag ) { double *d = a.diag + (size_t)site*4; d[0] = rec->max_qual; d[1] = rec->lk_sum; d[2] = rec->ref_lk; d[3] = rec->gap; }
        }
        __syncthreads();
    }
}

cudaError_t launch_pair_kernel(int nals, const KArgs &a, int nsm, cudaStream_t st)
{
    const size_t smem = align128(sizeof(PWTables));
    const int grid = nsm*PW_MINCTA;
    switch ( nals )
    {
        case 3: mcall_pair_phase2_kernel<3><<<grid, PW_WARPS*32, smem, st>>>(a); break;
        case 4: mcall_pair_phase2_kernel<4><<<grid, PW_WARPS*32, smem, st>>>(a); break;
        case 5: mcall_pair_phase2_kernel<5><<<grid, PW_WARPS*32, smem, st>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}   // namespace mcb
