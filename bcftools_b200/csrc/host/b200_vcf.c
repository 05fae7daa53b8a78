/*  b200_vcf.c -- text VCF header / record model without htslib; see include/b200_vcf.h.
 *  Reference call sites this serves: vcfcall.c:471-499 (read), mcall.c:1444-1510 (typed getters), mcall.c:1583-1681
 *  (record edits), vcfcall.c:1147 (write).  The typed-vector and formatting conventions are htslib's ([htslib] vcf.c:
 *  vcf_parse_format, bcf_fmt_array, bcf_format_gt, vcf_format; kstring.c: kputd), restated from their documented
 *  behaviour and pinned by byte-for-byte round trips of the reference's own test VCFs (tests/test_vcf_text.py).  */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <errno.h>
#include "b200_vcf.h"

/* ---- growable string ------------------------------------------------------------------------------- */
static int str_reserve(b200_str_t *s, size_t extra)
{
    if ( s->l + extra + 1 <= s->m ) return 0;
    size_t m = s->m ? s->m : 64;
    while ( m < s->l + extra + 1 ) m *= 2;
    char *p = (char*) realloc(s->s, m);
    if ( !p ) return -1;
    s->s = p; s->m = m;
    return 0;
}
int b200_str_putsn(b200_str_t *s, const char *p, size_t n)
{
    if ( str_reserve(s, n) ) return -1;
    memcpy(s->s + s->l, p, n); s->l += n; s->s[s->l] = 0;
    return 0;
}
int b200_str_puts(b200_str_t *s, const char *p) { return b200_str_putsn(s, p, strlen(p)); }
int b200_str_putc(b200_str_t *s, int c)
{
    if ( s->l + 2 > s->m && str_reserve(s, 1) ) return -1;
    s->s[s->l++] = (char)c; s->s[s->l] = 0;
    return 0;
}
/*  decimal digits by hand ([htslib] kputw / kputl do the same): a 2,504-sample record prints ~30,000 integers  */
int b200_str_putw(b200_str_t *s, long long v)
{
    char buf[24]; int n = 0;
    unsigned long long x = v<0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { buf[n++] = (char)('0' + x % 10); x /= 10; } while ( x );
    if ( v<0 ) buf[n++] = '-';
    if ( s->l + (size_t)n + 1 > s->m && str_reserve(s, (size_t)n) ) return -1;
    while ( n ) s->s[s->l++] = buf[--n];
    s->s[s->l] = 0;
    return 0;
}
/*  [htslib] kputd: "%g" for |d| outside [1e-4, 999999]; otherwise the first six significant digits of trunc(d*1e10),
 *  rounded half up at the seventh, trailing zeros (and a bare '.') culled.  */
int b200_str_putd(b200_str_t *s, double d)
{
    char buf[24], *cp = buf+20, *ep;
    if ( d==0 ) return signbit(d) ? b200_str_putsn(s, "-0", 2) : b200_str_putc(s, '0');
    if ( d<0 ) { if ( b200_str_putc(s, '-') ) return -1; d = -d; }
    if ( !(d >= 0.0001 && d <= 999999) )
    {
        char tmp[64]; snprintf(tmp, sizeof tmp, "%g", d);
        return b200_str_puts(s, tmp);
    }
    uint64_t i = (uint64_t)(d*10000000000LL);
    if ( d<.0001 ) i += 0;
    else if ( d<0.001 ) i += 5;
    else if ( d<0.01 ) i += 50;
    else if ( d<0.1 ) i += 500;
    else if ( d<1 ) i += 5000;
    else if ( d<10 ) i += 50000;
    else if ( d<100 ) i += 500000;
    else if ( d<1000 ) i += 5000000;
    else if ( d<10000 ) i += 50000000;
    else if ( d<100000 ) i += 500000000;
    else i += 5000000000LL;
    do { *--cp = (char)('0' + i%10); i /= 10; } while ( i >= 1 );
    buf[20] = 0;
    int p = (int)(buf+20-cp);
    if ( p <= 10 )      /* d < 1: pad to ten decimals, "0." in front */
    {
        cp[6] = 0; ep = cp+5;
        while ( p < 10 ) { *--cp = '0'; p++; }
        *--cp = '.'; *--cp = '0';
    }
    else
    {
        char *xp = --cp;
        while ( p > 10 ) { xp[0] = xp[1]; p--; xp++; }
        xp[0] = '.';
        cp[7] = 0; ep = cp+6;
        if ( cp[6]=='.' ) cp[6] = 0;
    }
    while ( *ep=='0' && ep > cp ) ep--;
    char *z = ep+1;
    while ( ep > cp )
    {
        if ( *ep=='.' ) { if ( z[-1]=='.' ) z[-1] = 0; else z[0] = 0; break; }
        ep--;
    }
    return b200_str_puts(s, cp);
}

/* ---- header ---------------------------------------------------------------------------------------- */
static char *dupn(const char *p, size_t n)
{
    char *d = (char*) malloc(n+1);
    if ( !d ) return NULL;
    memcpy(d, p, n); d[n] = 0;
    return d;
}
/*  value of key= inside a "##INFO=<...>" line (unquoted values only); returns length, 0 if absent  */
static size_t hdr_field(const char *line, const char *key, const char **val)
{
    const char *p = strchr(line, '<');
    size_t kl = strlen(key);
    if ( !p ) return 0;
    p++;
    while ( *p && *p!='>' )
    {
        const char *e = p;
        int inq = 0;
        while ( *e && (inq || (*e!=',' && *e!='>')) ) { if ( *e=='"' ) inq = !inq; e++; }
        if ( (size_t)(e-p) > kl && !strncmp(p, key, kl) && p[kl]=='=' ) { *val = p+kl+1; return (size_t)(e-(p+kl+1)); }
        if ( *e!=',' ) break;
        p = e+1;
    }
    return 0;
}
static int hdr_line_class(const char *line)     /* 0 INFO, 1 FORMAT, -1 other */
{
    if ( !strncmp(line, "##INFO=<", 8) ) return 0;
    if ( !strncmp(line, "##FORMAT=<", 10) ) return 1;
    return -1;
}
static int hdr_add_def(b200_vhdr_t *h, const char *line)
{
    int cls = hdr_line_class(line);
    if ( cls<0 ) return 0;
    const char *v; size_t n = hdr_field(line, "ID", &v);
    if ( !n ) return 0;
    if ( h->ndefs==h->mdefs )
    {
        int m = h->mdefs ? 2*h->mdefs : 32;
        b200_vdef_t *d = (b200_vdef_t*) realloc(h->defs, sizeof(*d)*m);
        if ( !d ) return -1;
        h->defs = d; h->mdefs = m;
    }
    b200_vdef_t *d = &h->defs[h->ndefs];
    memset(d, 0, sizeof *d);
    d->id = dupn(v, n); d->is_fmt = cls;
    if ( !d->id ) return -1;
    d->vl = B200_VL_VAR; d->number = 0; d->type = B200_HT_STR;
    n = hdr_field(line, "Number", &v);
    if ( n )
    {
        if ( n==1 && *v=='A' ) d->vl = B200_VL_A;
        else if ( n==1 && *v=='G' ) d->vl = B200_VL_G;
        else if ( n==1 && *v=='R' ) d->vl = B200_VL_R;
        else if ( n==1 && *v=='.' ) d->vl = B200_VL_VAR;
        else { d->vl = B200_VL_FIXED; d->number = atoi(v); }
    }
    n = hdr_field(line, "Type", &v);
    if ( n )
    {
        if ( !strncmp(v, "Integer", n) ) d->type = B200_HT_INT;
        else if ( !strncmp(v, "Float", n) ) d->type = B200_HT_REAL;
        else if ( !strncmp(v, "Flag", n) ) d->type = B200_HT_FLAG;
        else d->type = B200_HT_STR;
    }
    h->ndefs++;
    return 0;
}
const b200_vdef_t *b200_vhdr_def(const b200_vhdr_t *h, int is_fmt, const char *id)
{
    for (int i=0; i<h->ndefs; i++)
        if ( h->defs[i].is_fmt==is_fmt && !strcmp(h->defs[i].id, id) ) return &h->defs[i];
    return NULL;
}
static int hdr_push_line(b200_vhdr_t *h, char *line)
{
    if ( h->nlines==h->mlines )
    {
        int m = h->mlines ? 2*h->mlines : 64;
        char **p = (char**) realloc(h->lines, sizeof(char*)*m);
        if ( !p ) return -1;
        h->lines = p; h->mlines = m;
    }
    h->lines[h->nlines++] = line;
    return 0;
}
int b200_vhdr_append(b200_vhdr_t *h, const char *line)
{
    size_t n = strlen(line);
    while ( n && (line[n-1]=='\n' || line[n-1]=='\r') ) n--;
    char *copy = dupn(line, n);
    if ( !copy ) return -1;
    int cls = hdr_line_class(copy);
    if ( cls>=0 )
    {
        const char *v; size_t l = hdr_field(copy, "ID", &v);
        if ( l )
        {
            char *id = dupn(v, l);
            const b200_vdef_t *d = id ? b200_vhdr_def(h, cls, id) : NULL;
            free(id);
            if ( d ) { free(copy); return 0; }      /* already defined: the first definition stays */
        }
    }
    if ( hdr_push_line(h, copy) || hdr_add_def(h, copy) ) return -1;
    return 0;
}
void b200_vhdr_remove(b200_vhdr_t *h, int is_fmt, const char *id)
{
    int j = 0;
    for (int i=0; i<h->nlines; i++)
    {
        int drop = 0;
        if ( hdr_line_class(h->lines[i])==is_fmt )
        {
            const char *v; size_t l = hdr_field(h->lines[i], "ID", &v);
            drop = l==strlen(id) && !strncmp(v, id, l);
        }
        if ( drop ) free(h->lines[i]); else h->lines[j++] = h->lines[i];
    }
    h->nlines = j;
    j = 0;
    for (int i=0; i<h->ndefs; i++)
    {
        if ( h->defs[i].is_fmt==is_fmt && !strcmp(h->defs[i].id, id) ) free(h->defs[i].id);
        else h->defs[j++] = h->defs[i];
    }
    h->ndefs = j;
}
b200_vhdr_t *b200_vhdr_parse(const char *text, size_t len, size_t *consumed)
{
    b200_vhdr_t *h = (b200_vhdr_t*) calloc(1, sizeof *h);
    if ( !h ) return NULL;
    size_t off = 0;
    int have_chrom = 0;
    while ( off < len && text[off]=='#' )
    {
        const char *e = (const char*) memchr(text+off, '\n', len-off);
        size_t ll = e ? (size_t)(e-(text+off)) : len-off;
        size_t n = ll;
        while ( n && text[off+n-1]=='\r' ) n--;
        if ( n>1 && text[off+1]=='#' )
        {
            char *copy = dupn(text+off, n);
            if ( !copy || hdr_push_line(h, copy) || hdr_add_def(h, copy) ) { b200_vhdr_destroy(h); return NULL; }
        }
        else        /* #CHROM POS ID REF ALT QUAL FILTER INFO [FORMAT samples...] */
        {
            int col = 0; size_t p = 0;
            while ( p <= n )
            {
                size_t q = p;
                while ( q<n && text[off+q]!='\t' ) q++;
                if ( col>=9 )
                {
                    char **s = (char**) realloc(h->samples, sizeof(char*)*(h->nsamples+1));
                    if ( !s ) { b200_vhdr_destroy(h); return NULL; }
                    h->samples = s;
                    h->samples[h->nsamples++] = dupn(text+off+p, q-p);
                }
                col++; p = q+1;
            }
            h->n_in_samples = h->nsamples;
            have_chrom = 1;
            off += ll + (e ? 1 : 0);
            break;
        }
        off += ll + (e ? 1 : 0);
    }
    if ( !have_chrom ) { b200_vhdr_destroy(h); return NULL; }
    if ( consumed ) *consumed = off;
    return h;
}
void b200_vhdr_destroy(b200_vhdr_t *h)
{
    if ( !h ) return;
    for (int i=0; i<h->nlines; i++) free(h->lines[i]);
    for (int i=0; i<h->nsamples; i++) free(h->samples[i]);
    for (int i=0; i<h->ndefs; i++) free(h->defs[i].id);
    free(h->lines); free(h->samples); free(h->defs); free(h->smpl_map); free(h);
}
int b200_vhdr_subset(b200_vhdr_t *h, int n, const int *map)
{
    char **s = (char**) calloc(n ? n : 1, sizeof(char*));
    int *m = (int*) malloc(sizeof(int)*(n ? n : 1));
    if ( !s || !m ) { free(s); free(m); return -1; }
    for (int i=0; i<n; i++)
    {
        if ( map[i]<0 || map[i]>=h->nsamples ) { for (int j=0; j<i; j++) free(s[j]); free(s); free(m); return -1; }
        s[i] = dupn(h->samples[map[i]], strlen(h->samples[map[i]]));
        m[i] = h->smpl_map ? h->smpl_map[map[i]] : map[i];
    }
    for (int i=0; i<h->nsamples; i++) free(h->samples[i]);
    free(h->samples); free(h->smpl_map);
    h->samples = s; h->nsamples = n; h->smpl_map = m;
    return 0;
}
int b200_vhdr_format(const b200_vhdr_t *h, b200_str_t *out)
{
    for (int i=0; i<h->nlines; i++)
        if ( b200_str_puts(out, h->lines[i]) || b200_str_putc(out, '\n') ) return -1;
    if ( b200_str_puts(out, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO") ) return -1;
    if ( h->nsamples )
    {
        if ( b200_str_puts(out, "\tFORMAT") ) return -1;
        for (int i=0; i<h->nsamples; i++)
            if ( b200_str_putc(out, '\t') || b200_str_puts(out, h->samples[i]) ) return -1;
    }
    return b200_str_putc(out, '\n');
}

/* ---- record ---------------------------------------------------------------------------------------- */
static void *rec_own(b200_vrec_t *r, void *p)
{
    if ( !p ) return NULL;
    if ( r->nowned==r->mowned )
    {
        int m = r->mowned ? 2*r->mowned : 16;
        void **o = (void**) realloc(r->owned, sizeof(void*)*m);
        if ( !o ) { free(p); return NULL; }
        r->owned = o; r->mowned = m;
    }
    r->owned[r->nowned++] = p;
    return p;
}
static char *rec_strdup(b200_vrec_t *r, const char *s) { return (char*) rec_own(r, dupn(s, strlen(s))); }
static float f32_from_bits(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static uint32_t f32_bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

b200_vrec_t *b200_vrec_new(int nsmpl)
{
    b200_vrec_t *r = (b200_vrec_t*) calloc(1, sizeof *r);
    if ( !r ) return NULL;
    r->nsmpl = nsmpl;
    r->qual = f32_from_bits(B200_F32_MISSING_BITS);
    r->chrom = r->id = r->filter = (char*)".";
    return r;
}
void b200_vrec_destroy(b200_vrec_t *r)
{
    if ( !r ) return;
    for (int i=0; i<r->nowned; i++) free(r->owned[i]);
    free(r->owned); free(r->line); free(r->allele); free(r->info); free(r->fmt); free(r);
}
static int split_count(const char *s, char sep)
{
    int n = 1;
    for (; *s; s++) if ( *s==sep ) n++;
    return n;
}
b200_vrec_t *b200_vrec_parse(const b200_vhdr_t *h, const char *line, size_t len)
{
    while ( len && (line[len-1]=='\n' || line[len-1]=='\r') ) len--;
    b200_vrec_t *r = b200_vrec_new(h->nsamples);
    if ( !r ) return NULL;
    r->line = dupn(line, len);
    if ( !r->line ) { b200_vrec_destroy(r); return NULL; }
    /* split the columns in place */
    int ncol = split_count(r->line, '\t');
    char **col = (char**) rec_own(r, malloc(sizeof(char*)*ncol));
    if ( !col ) { b200_vrec_destroy(r); return NULL; }
    {
        char *p = r->line; int i = 0;
        col[i++] = p;
        for (; *p; p++) if ( *p=='\t' ) { *p = 0; col[i++] = p+1; }
    }
    if ( ncol < 8 ) { b200_vrec_destroy(r); return NULL; }
    r->chrom = col[0];
    { char *e; errno = 0; long long v = strtoll(col[1], &e, 10); if ( e==col[1] || *e ) { b200_vrec_destroy(r); return NULL; } r->pos = v-1; }
    r->id = col[2];
    /* alleles: REF, then ALT split at commas ("." = none) */
    {
        int nalt = strcmp(col[4], ".") ? split_count(col[4], ',') : 0;
        r->allele = (char**) malloc(sizeof(char*)*(1+nalt));
        if ( !r->allele ) { b200_vrec_destroy(r); return NULL; }
        r->allele[0] = col[3]; r->n_allele = 1;
        if ( nalt )
        {
            char *p = col[4];
            r->allele[r->n_allele++] = p;
            for (; *p; p++) if ( *p==',' ) { *p = 0; r->allele[r->n_allele++] = p+1; }
        }
    }
    if ( !strcmp(col[5], ".") ) r->qual = f32_from_bits(B200_F32_MISSING_BITS);
    else { char *e; r->qual = (float) strtod(col[5], &e); if ( e==col[5] ) { b200_vrec_destroy(r); return NULL; } }
    r->filter = col[6];
    if ( strcmp(col[7], ".") )
    {
        r->m_info = split_count(col[7], ';');
        r->info = (b200_vinfo_t*) calloc(r->m_info, sizeof(b200_vinfo_t));
        if ( !r->info ) { b200_vrec_destroy(r); return NULL; }
        char *p = col[7];
        while ( p )
        {
            char *e = strchr(p, ';');
            if ( e ) *e = 0;
            if ( *p )
            {
                char *eq = strchr(p, '=');
                if ( eq ) *eq = 0;
                r->info[r->n_info].key = p; r->info[r->n_info].val = eq ? eq+1 : NULL;
                r->n_info++;
            }
            p = e ? e+1 : NULL;
        }
    }
    if ( ncol > 8 && h->nsamples && strcmp(col[8], ".") )
    {
        if ( ncol < 9 + h->n_in_samples ) { b200_vrec_destroy(r); return NULL; }
        r->m_fmt = split_count(col[8], ':');
        r->fmt = (b200_vfmt_t*) calloc(r->m_fmt, sizeof(b200_vfmt_t));
        if ( !r->fmt ) { b200_vrec_destroy(r); return NULL; }
        char *p = col[8];
        while ( p )
        {
            char *e = strchr(p, ':');
            if ( e ) *e = 0;
            b200_vfmt_t *f = &r->fmt[r->n_fmt++];
            f->key = p; f->kind = B200_FMT_TEXT;
            f->txt = (char**) rec_own(r, malloc(sizeof(char*)*(r->nsmpl ? r->nsmpl : 1)));
            if ( !f->txt ) { b200_vrec_destroy(r); return NULL; }
            p = e ? e+1 : NULL;
        }
        if ( h->smpl_map )      /* bcf_subset works on the parsed record: a vector keeps the length of the longest INPUT sample */
            for (int i=0; i<h->n_in_samples; i++)
            {
                const char *q = col[9+i];
                for (int j=0; j<r->n_fmt && q; j++)
                {
                    int c = 1;
                    for (; *q && *q!=':'; q++) if ( *q==',' ) c++;
                    if ( c > r->fmt[j].n ) r->fmt[j].n = c;
                    q = *q ? q+1 : NULL;
                }
            }
        for (int i=0; i<r->nsmpl; i++)
        {
            char *q = col[9 + (h->smpl_map ? h->smpl_map[i] : i)];
            for (int j=0; j<r->n_fmt; j++)
            {
                if ( !q ) { r->fmt[j].txt[i] = (char*)"."; continue; }     /* trailing fields dropped: missing */
                char *e = strchr(q, ':');
                if ( e ) *e = 0;
                r->fmt[j].txt[i] = q;
                q = e ? e+1 : NULL;
            }
        }
    }
    return r;
}

static void fmt_ints_one(b200_str_t *out, const int32_t *v, int n)
{
    int j;
    for (j=0; j<n && v[j]!=B200_I32_VECTOR_END; j++)
    {
        if ( j ) b200_str_putc(out, ',');
        if ( v[j]==B200_I32_MISSING ) b200_str_putc(out, '.'); else b200_str_putw(out, v[j]);
    }
    if ( n && j==0 ) b200_str_putc(out, '.');
}
static void fmt_floats_one(b200_str_t *out, const float *v, int n)
{
    int j;
    for (j=0; j<n && f32_bits(v[j])!=B200_F32_VECTOR_END_BITS; j++)
    {
        if ( j ) b200_str_putc(out, ',');
        if ( f32_bits(v[j])==B200_F32_MISSING_BITS ) b200_str_putc(out, '.'); else b200_str_putd(out, v[j]);
    }
    if ( n && j==0 ) b200_str_putc(out, '.');
}
static void fmt_gt_one(b200_str_t *out, const int32_t *v, int n)      /* [htslib] bcf_format_gt */
{
    int j;
    for (j=0; j<n && v[j]!=B200_I32_VECTOR_END; j++)
    {
        if ( j ) b200_str_putc(out, (v[j]&1) ? '|' : '/');
        if ( !(v[j]>>1) ) b200_str_putc(out, '.'); else b200_str_putw(out, (v[j]>>1) - 1);
    }
    if ( j==0 ) b200_str_putc(out, '.');
}
int b200_vrec_format(const b200_vrec_t *r, b200_str_t *out)
{
    b200_str_puts(out, r->chrom); b200_str_putc(out, '\t');
    b200_str_putw(out, r->pos+1); b200_str_putc(out, '\t');
    b200_str_puts(out, r->id); b200_str_putc(out, '\t');
    b200_str_puts(out, r->n_allele ? r->allele[0] : "."); b200_str_putc(out, '\t');
    if ( r->n_allele > 1 )
        for (int i=1; i<r->n_allele; i++) { if ( i>1 ) b200_str_putc(out, ','); b200_str_puts(out, r->allele[i]); }
    else b200_str_putc(out, '.');
    b200_str_putc(out, '\t');
    if ( f32_bits(r->qual)==B200_F32_MISSING_BITS ) b200_str_putc(out, '.'); else b200_str_putd(out, r->qual);
    b200_str_putc(out, '\t');
    b200_str_puts(out, r->filter); b200_str_putc(out, '\t');
    if ( r->n_info )
        for (int i=0; i<r->n_info; i++)
        {
            if ( i ) b200_str_putc(out, ';');
            b200_str_puts(out, r->info[i].key);
            if ( r->info[i].val ) { b200_str_putc(out, '='); b200_str_puts(out, r->info[i].val); }
        }
    else b200_str_putc(out, '.');
    if ( r->nsmpl )
    {
        b200_str_putc(out, '\t');
        if ( r->n_fmt )
            for (int j=0; j<r->n_fmt; j++) { if ( j ) b200_str_putc(out, ':'); b200_str_puts(out, r->fmt[j].key); }
        else b200_str_putc(out, '.');
        for (int i=0; i<r->nsmpl; i++)
        {
            b200_str_putc(out, '\t');
            if ( !r->n_fmt ) { b200_str_putc(out, '.'); continue; }
            for (int j=0; j<r->n_fmt; j++)
            {
                const b200_vfmt_t *f = &r->fmt[j];
                if ( j ) b200_str_putc(out, ':');
                if ( f->kind==B200_FMT_TEXT ) b200_str_puts(out, f->txt[i]);
                else if ( f->kind==B200_FMT_INT ) fmt_ints_one(out, f->iv + (size_t)i*f->n, f->n);
                else if ( f->kind==B200_FMT_REAL ) fmt_floats_one(out, f->fv + (size_t)i*f->n, f->n);
                else fmt_gt_one(out, f->iv + (size_t)i*f->n, f->n);
            }
        }
    }
    return b200_str_putc(out, '\n');
}

/* ---- getters --------------------------------------------------------------------------------------- */
const char *b200_vrec_info(const b200_vrec_t *r, const char *key, int *found)
{
    for (int i=0; i<r->n_info; i++)
        if ( !strcmp(r->info[i].key, key) ) { if ( found ) *found = 1; return r->info[i].val; }
    if ( found ) *found = 0;
    return NULL;
}
static int grow(void **dst, int *m, int need, size_t es)
{
    if ( need <= *m ) return 0;
    int n = *m ? *m : 16;
    while ( n < need ) n *= 2;
    void *p = realloc(*dst, es*(size_t)n);
    if ( !p ) return -1;
    *dst = p; *m = n;
    return 0;
}
int b200_vrec_info_floats(const b200_vrec_t *r, const char *key, float **dst, int *mdst)
{
    int found; const char *v = b200_vrec_info(r, key, &found);
    if ( !found || !v ) return -1;
    int n = split_count(v, ',');
    if ( grow((void**)dst, mdst, n, sizeof(float)) ) return -1;
    for (int i=0; i<n; i++)
    {
        if ( *v=='.' && (v[1]==',' || !v[1]) ) { (*dst)[i] = f32_from_bits(B200_F32_MISSING_BITS); v++; }
        else { char *e; (*dst)[i] = (float) strtod(v, &e); if ( e==v ) return -1; v = e; }
        if ( *v==',' ) v++;
    }
    return n;
}
int b200_vrec_info_ints(const b200_vrec_t *r, const char *key, int32_t **dst, int *mdst)
{
    int found; const char *v = b200_vrec_info(r, key, &found);
    if ( !found || !v ) return -1;
    int n = split_count(v, ',');
    if ( grow((void**)dst, mdst, n, sizeof(int32_t)) ) return -1;
    for (int i=0; i<n; i++)
    {
        if ( *v=='.' && (v[1]==',' || !v[1]) ) { (*dst)[i] = B200_I32_MISSING; v++; }
        else { char *e; (*dst)[i] = (int32_t) strtol(v, &e, 10); if ( e==v ) return -1; v = e; }
        if ( *v==',' ) v++;
    }
    return n;
}
b200_vfmt_t *b200_vrec_fmt(const b200_vrec_t *r, const char *key)
{
    for (int j=0; j<r->n_fmt; j++)
        if ( !strcmp(r->fmt[j].key, key) ) return &r->fmt[j];
    return NULL;
}
int b200_vrec_fmt_ints(const b200_vrec_t *r, const char *key, int32_t **dst, int *mdst)
{
    const b200_vfmt_t *f = b200_vrec_fmt(r, key);
    if ( !f ) return -1;
    if ( f->kind==B200_FMT_INT || f->kind==B200_FMT_GT )
    {
        if ( grow((void**)dst, mdst, r->nsmpl*f->n, sizeof(int32_t)) ) return -1;
        memcpy(*dst, f->iv, sizeof(int32_t)*(size_t)r->nsmpl*f->n);
        return r->nsmpl*f->n;
    }
    if ( f->kind!=B200_FMT_TEXT ) return -2;
    int n = f->n > 1 ? f->n : 1;
    for (int i=0; i<r->nsmpl; i++) { int c = split_count(f->txt[i], ','); if ( c>n ) n = c; }
    if ( grow((void**)dst, mdst, r->nsmpl*n, sizeof(int32_t)) ) return -1;
    for (int i=0; i<r->nsmpl; i++)
    {
        const char *v = f->txt[i];
        int32_t *d = *dst + (size_t)i*n;
        int k = 0;
        while ( *v && k<n )
        {
            if ( *v=='.' && (v[1]==',' || !v[1]) ) { d[k++] = B200_I32_MISSING; v++; }
            else if ( (unsigned)(*v - '0') < 10u && (unsigned)(v[1] - '0') >= 10u ) { d[k++] = *v - '0'; v++; }        /* one digit: most PL values of a confident call */
            else if ( (unsigned)(*v - '0') < 10u )                  /* plain digits (what every PL / AD is): no locale, no errno */
            {
                long long x = 0; int nd = 0;
                while ( (unsigned)(*v - '0') < 10u && nd < 18 ) { x = x*10 + (*v - '0'); v++; nd++; }
                if ( (unsigned)(*v - '0') < 10u ) return -2;
                d[k++] = x > INT32_MAX ? INT32_MAX : (int32_t)x;
            }
            else { char *e; long x = strtol(v, &e, 10); if ( e==v ) return -2; d[k++] = (int32_t)x; v = e; }
            if ( *v==',' ) v++; else break;
        }
        if ( !k ) d[k++] = B200_I32_MISSING;
        for (; k<n; k++) d[k] = B200_I32_VECTOR_END;
    }
    return r->nsmpl*n;
}

/* ---- setters --------------------------------------------------------------------------------------- */
static int info_slot(b200_vrec_t *r, const char *key, int create)
{
    for (int i=0; i<r->n_info; i++) if ( !strcmp(r->info[i].key, key) ) return i;
    if ( !create ) return -1;
    if ( r->n_info==r->m_info )
    {
        int m = r->m_info ? 2*r->m_info : 8;
        b200_vinfo_t *p = (b200_vinfo_t*) realloc(r->info, sizeof(*p)*m);
        if ( !p ) return -1;
        r->info = p; r->m_info = m;
    }
    r->info[r->n_info].key = rec_strdup(r, key);
    r->info[r->n_info].val = NULL;
    if ( !r->info[r->n_info].key ) return -1;
    return r->n_info++;
}
static void info_remove(b200_vrec_t *r, const char *key)
{
    int i = info_slot(r, key, 0);
    if ( i<0 ) return;
    memmove(&r->info[i], &r->info[i+1], sizeof(b200_vinfo_t)*(r->n_info-i-1));
    r->n_info--;
}
int b200_vrec_set_info_text(b200_vrec_t *r, const char *key, const char *val)
{
    int i = info_slot(r, key, 1);
    if ( i<0 ) return -1;
    r->info[i].val = val ? rec_strdup(r, val) : NULL;
    return (val && !r->info[i].val) ? -1 : 0;
}
int b200_vrec_set_info_ints(b200_vrec_t *r, const char *key, const int32_t *v, int n)
{
    if ( !n ) { info_remove(r, key); return 0; }
    b200_str_t s = {0,0,0};
    fmt_ints_one(&s, v, n);
    int ret = b200_vrec_set_info_text(r, key, s.s ? s.s : ".");
    free(s.s);
    return ret;
}
int b200_vrec_set_info_floats(b200_vrec_t *r, const char *key, const float *v, int n)
{
    if ( !n ) { info_remove(r, key); return 0; }
    b200_str_t s = {0,0,0};
    fmt_floats_one(&s, v, n);
    int ret = b200_vrec_set_info_text(r, key, s.s ? s.s : ".");
    free(s.s);
    return ret;
}
static b200_vfmt_t *fmt_slot(b200_vrec_t *r, const char *key, int first)
{
    b200_vfmt_t *f = b200_vrec_fmt(r, key);
    if ( f ) return f;
    if ( r->n_fmt==r->m_fmt )
    {
        int m = r->m_fmt ? 2*r->m_fmt : 8;
        b200_vfmt_t *p = (b200_vfmt_t*) realloc(r->fmt, sizeof(*p)*m);
        if ( !p ) return NULL;
        r->fmt = p; r->m_fmt = m;
    }
    int at = r->n_fmt;
    if ( first ) { memmove(&r->fmt[1], &r->fmt[0], sizeof(b200_vfmt_t)*r->n_fmt); at = 0; }
    r->n_fmt++;
    f = &r->fmt[at];
    memset(f, 0, sizeof *f);
    f->key = rec_strdup(r, key);
    return f->key ? f : NULL;
}
static void fmt_remove(b200_vrec_t *r, const char *key)
{
    b200_vfmt_t *f = b200_vrec_fmt(r, key);
    if ( !f ) return;
    int i = (int)(f - r->fmt);
    memmove(&r->fmt[i], &r->fmt[i+1], sizeof(b200_vfmt_t)*(r->n_fmt-i-1));
    r->n_fmt--;
}
static int fmt_set(b200_vrec_t *r, const char *key, int kind, const void *v, int nvals)
{
    if ( !nvals ) { fmt_remove(r, key); return 0; }
    if ( !r->nsmpl || nvals % r->nsmpl ) return -1;
    b200_vfmt_t *f = fmt_slot(r, key, kind==B200_FMT_GT);
    if ( !f ) return -1;
    void *copy = rec_own(r, malloc(4*(size_t)nvals));
    if ( !copy ) return -1;
    memcpy(copy, v, 4*(size_t)nvals);
    f->kind = kind; f->n = nvals / r->nsmpl; f->txt = NULL;
    f->iv = kind==B200_FMT_REAL ? NULL : (int32_t*)copy;
    f->fv = kind==B200_FMT_REAL ? (float*)copy : NULL;
    return 0;
}
int b200_vrec_set_fmt_ints(b200_vrec_t *r, const char *key, const int32_t *v, int nvals) { return fmt_set(r, key, B200_FMT_INT, v, nvals); }
int b200_vrec_set_fmt_floats(b200_vrec_t *r, const char *key, const float *v, int nvals) { return fmt_set(r, key, B200_FMT_REAL, v, nvals); }
int b200_vrec_set_genotypes(b200_vrec_t *r, const int32_t *gts, int nvals) { return fmt_set(r, "GT", B200_FMT_GT, gts, nvals); }
int b200_vrec_set_alleles(b200_vrec_t *r, const char *const *als, int n)
{
    char **a = (char**) malloc(sizeof(char*)*(n ? n : 1));
    if ( !a ) return -1;
    for (int i=0; i<n; i++) { a[i] = rec_strdup(r, als[i]); if ( !a[i] ) { free(a); return -1; } }
    free(r->allele);
    r->allele = a; r->n_allele = n;
    return 0;
}
