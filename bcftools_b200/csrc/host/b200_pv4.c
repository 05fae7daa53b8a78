/*  b200_pv4.c -- INFO/PV4 (`call -a PV4`): the four bias p-values test16() derives from INFO/I16 (ccall.c:103-138, called at
 *  mcall.c:1668-1677).  test16 needs two special functions of htslib's kfunc.c, which is not in the reference tree:
 *  kt_fisher_exact (Fisher's exact test on the 2x2 strand table, hypergeometric terms accumulated outwards from both tails)
 *  and kf_betai (regularised incomplete beta by the modified Lentz continued fraction, Lanczos log-gamma), restated here
 *  from those published algorithms.  Pinned by the PV4 values in the reference's own expected outputs (test/mpileup.c.*.out hold
 *  the PV4 of records whose I16 is in test/mpileup.c.vcf; tests/test_vcfcall_host.py).  */
#include <math.h>
#include <stdlib.h>
#include "b200_vcfcall.h"

#define KF_GAMMA_EPS 1e-14
#define KF_TINY 1e-290

static double kf_lgamma(double z)
{
    double x = 0;
    x += 0.1659470187408462e-06 / (z+7);
    x += 0.9934937113930748e-05 / (z+6);
    x -= 0.1385710331296526     / (z+5);
    x += 12.50734324009056      / (z+4);
    x -= 176.6150291498386      / (z+3);
    x += 771.3234287757674      / (z+2);
    x -= 1259.139216722289      / (z+1);
    x += 676.5203681218835      / z;
    x += 0.9999999999995183;
    return log(x) - 5.58106146679532777 - z + (z-0.5) * log(z+6.5);
}
static double kf_betai_aux(double a, double b, double x)
{
    double C, D, f;
    int j;
    if ( x==0. ) return 0.;
    if ( x==1. ) return 1.;
    f = 1.; C = f; D = 0.;
    for (j = 1; j < 200; ++j)
    {
        double aa, d;
        int m = j>>1;
        aa = (j&1) ? -(a + m) * (a + b + m) * x / ((a + 2*m) * (a + 2*m + 1))
                   : m * (b - m) * x / ((a + 2*m - 1) * (a + 2*m));
        D = 1. + aa * D;
        if ( D < KF_TINY ) D = KF_TINY;
        C = 1. + aa / C;
        if ( C < KF_TINY ) C = KF_TINY;
        D = 1. / D;
        d = C * D;
        f *= d;
        if ( fabs(d - 1.) < KF_GAMMA_EPS ) break;
    }
    return exp(kf_lgamma(a+b) - kf_lgamma(a) - kf_lgamma(b) + a * log(x) + b * log(1.-x)) / a / f;
}
static double kf_betai(double a, double b, double x)
{
    return x < (a + 1.) / (a + b + 2.) ? kf_betai_aux(a, b, x) : 1. - kf_betai_aux(b, a, 1. - x);
}

static double lbinom(int n, int k)
{
    if ( k==0 || n==k ) return 0;
    return lgamma(n+1) - lgamma(k+1) - lgamma(n-k+1);
}
static double hypergeo(int n11, int n1_, int n_1, int n)
{
    return exp(lbinom(n1_, n11) + lbinom(n-n1_, n_1-n11) - lbinom(n, n_1));
}
typedef struct { int n11, n1_, n_1, n; double p; } hgacc_t;
static double hypergeo_acc(int n11, int n1_, int n_1, int n, hgacc_t *aux)
{
    if ( n1_ || n_1 || n ) { aux->n11 = n11; aux->n1_ = n1_; aux->n_1 = n_1; aux->n = n; }
    else
    {
        if ( n11%11 && n11 + aux->n - aux->n1_ - aux->n_1 )
        {
            if ( n11==aux->n11 + 1 )
            {
                aux->p *= (double)(aux->n1_ - aux->n11) / n11 * (aux->n_1 - aux->n11) / (n11 + aux->n - aux->n1_ - aux->n_1);
                aux->n11 = n11;
                return aux->p;
            }
            if ( n11==aux->n11 - 1 )
            {
                aux->p *= (double)aux->n11 / (aux->n1_ - n11) * (aux->n11 + aux->n - aux->n1_ - aux->n_1) / (aux->n_1 - n11);
                aux->n11 = n11;
                return aux->p;
            }
        }
        aux->n11 = n11;
    }
    aux->p = hypergeo(aux->n11, aux->n1_, aux->n_1, aux->n);
    return aux->p;
}
static double fisher_exact(int n11, int n12, int n21, int n22, double *_left, double *_right, double *two)
{
    int i, j, max, min;
    double p, q, left, right;
    hgacc_t aux;
    int n1_ = n11 + n12, n_1 = n11 + n21, n = n11 + n12 + n21 + n22;
    max = (n_1 < n1_) ? n_1 : n1_;
    min = n1_ + n_1 - n;
    if ( min < 0 ) min = 0;
    *two = *_left = *_right = 1.;
    if ( min==max ) return 1.;
    q = hypergeo_acc(n11, n1_, n_1, n, &aux);
    p = hypergeo_acc(min, 0, 0, 0, &aux);
    for (left = 0., i = min + 1; p < 0.99999999 * q && i<=max; ++i) { left += p; p = hypergeo_acc(i, 0, 0, 0, &aux); }
    --i;
    if ( p < 1.00000001 * q ) left += p; else --i;
    p = hypergeo_acc(max, 0, 0, 0, &aux);
    for (right = 0., j = max - 1; p < 0.99999999 * q && j>=0; --j) { right += p; p = hypergeo_acc(j, 0, 0, 0, &aux); }
    ++j;
    if ( p < 1.00000001 * q ) right += p; else ++j;
    *two = left + right;
    if ( *two > 1. ) *two = 1.;
    if ( abs(i - n11) < abs(j - n11) ) right = 1. - left + q;
    else left = 1.0 - right + q;
    *_left = left; *_right = right;
    return q;
}

static double ttest(int n1, int n2, const float a[4])      /* ccall.c:103-113 */
{
    double t, v, u1, u2;
    if ( n1==0 || n2==0 || n1 + n2 < 3 ) return 1.0;
    u1 = (double)a[0] / n1; u2 = (double)a[2] / n2;
    if ( u1 <= u2 ) return 1.;
    t = (u1 - u2) / sqrt(((a[1] - n1 * u1 * u1) + (a[3] - n2 * u2 * u2)) / (n1 + n2 - 2) * (1./n1 + 1./n2));
    v = n1 + n2 - 2;
    return t < 0. ? 1. : .5 * kf_betai(.5*v, .5, v/(v+t*t));
}

int b200_pv4(const float *anno, float pv[4])       /* test16_core, ccall.c:115-138 */
{
    double p[4] = {1., 1., 1., 1.}, left, right;
    int depth = anno[0] + anno[1] + anno[2] + anno[3];
    int is_tested = (anno[0] + anno[1] > 0 && anno[2] + anno[3] > 0);
    for (int i=0; i<4; i++) pv[i] = 1.f;
    if ( depth==0 ) return 0;
    fisher_exact(anno[0], anno[1], anno[2], anno[3], &left, &right, &p[0]);
    for (int i=1; i<4; i++) p[i] = ttest(anno[0] + anno[1], anno[2] + anno[3], anno+4*i);
    for (int i=0; i<4; i++) pv[i] = (float)p[i];
    return is_tested;
}
