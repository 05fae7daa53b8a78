/*  b200_bcfio.c -- BCF2.2 container without htslib; see include/b200_bcfio.h (layout: hts-specs VCFv4.2 §6, SAMv1 §4.1).
 *  Serves the write side of vcfcall.c:1147 (`bcf_write1` with -Ob / -Ou) and the read side of vcfcall.c:471-499.  */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include "b200_bcfio.h"

#define BT_NULL 0
#define BT_INT8 1
#define BT_INT16 2
#define BT_INT32 3
#define BT_FLOAT 5
#define BT_CHAR 7

static char *dupn(const char *p, size_t n) { char *d = (char*) malloc(n+1); if ( d ) { memcpy(d, p, n); d[n] = 0; } return d; }
static void put_u16(b200_str_t *s, uint32_t v) { char b[2] = { (char)(v & 0xff), (char)(v >> 8) }; b200_str_putsn(s, b, 2); }
static void put_u32(b200_str_t *s, uint32_t v) { char b[4] = { (char)(v & 0xff), (char)((v>>8) & 0xff), (char)((v>>16) & 0xff), (char)(v>>24) }; b200_str_putsn(s, b, 4); }
static uint32_t get_u32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1]<<8 | (uint32_t)p[2]<<16 | (uint32_t)p[3]<<24; }
static uint32_t get_u16(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1]<<8; }

/* ---- BGZF (SAMv1 §4.1) ------------------------------------------------------------------------------- */
static int bgzf_block(const uint8_t *raw, size_t n, int level, b200_str_t *out)
{
    uint8_t cbuf[0x10000 + 64];
    z_stream zs; memset(&zs, 0, sizeof zs);
    if ( deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK ) return -1;
    zs.next_in = (Bytef*) raw; zs.avail_in = (uInt) n;
    zs.next_out = cbuf; zs.avail_out = sizeof cbuf;
    const int zr = deflate(&zs, Z_FINISH);
    const size_t clen = sizeof cbuf - zs.avail_out;
    deflateEnd(&zs);
    if ( zr != Z_STREAM_END || clen + 26 > 0x10000 ) return -1;
    static const char hd[12] = { 31, (char)139, 8, 4, 0, 0, 0, 0, 0, (char)255, 6, 0 };
    b200_str_putsn(out, hd, 12);
    b200_str_putsn(out, "BC", 2); put_u16(out, 2); put_u16(out, (uint32_t)(clen + 25));     /* BSIZE = block size - 1 */
    b200_str_putsn(out, (const char*)cbuf, clen);
    put_u32(out, (uint32_t) crc32(crc32(0L, Z_NULL, 0), raw, (uInt) n)); put_u32(out, (uint32_t) n);
    return 0;
}
int b200_bgzf_compress(const uint8_t *raw, size_t n, int level, b200_str_t *out)
{
    for (size_t off = 0; off < n; off += 0xff00)
        if ( bgzf_block(raw + off, n - off < 0xff00 ? n - off : 0xff00, level, out) ) return -1;
    return 0;
}
int b200_bgzf_finish(b200_str_t *out) { return bgzf_block((const uint8_t*)"", 0, 6, out); }
int b200_bgzf_decompress(const uint8_t *in, size_t n, b200_str_t *raw)
{
    size_t off = 0;
    if ( n < 18 || in[0]!=31 || in[1]!=139 ) return -1;
    while ( off + 18 <= n )
    {
        const uint8_t *b = in + off;
        if ( b[0]!=31 || b[1]!=139 || b[2]!=8 || !(b[3] & 4) ) return -1;
        const uint32_t xlen = get_u16(b + 10);
        uint32_t bsize = 0, x = 0;
        while ( x + 4 <= xlen )         /* the BC subfield carries the block size */
        {
            const uint8_t *f = b + 12 + x;
            const uint32_t slen = get_u16(f + 2);
            if ( f[0]=='B' && f[1]=='C' && slen==2 ) bsize = get_u16(f + 4) + 1;
            x += 4 + slen;
        }
        if ( !bsize || off + bsize > n || bsize < 12 + xlen + 8 ) return -1;
        const uint32_t isize = get_u32(b + bsize - 4), clen = bsize - 12 - xlen - 8;
        if ( isize )
        {
            uint8_t *dst = (uint8_t*) malloc(isize);
            if ( !dst ) return -1;
            z_stream zs; memset(&zs, 0, sizeof zs);
            if ( inflateInit2(&zs, -15) != Z_OK ) { free(dst); return -1; }
            zs.next_in = (Bytef*)(b + 12 + xlen); zs.avail_in = clen;
            zs.next_out = dst; zs.avail_out = isize;
            const int zr = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if ( zr != Z_STREAM_END || zs.avail_out || (uint32_t) crc32(crc32(0L, Z_NULL, 0), dst, isize) != get_u32(b + bsize - 8) ) { free(dst); return -1; }
            b200_str_putsn(raw, (const char*)dst, isize);
            free(dst);
        }
        off += bsize;
    }
    return off==n ? 0 : -1;
}

/* ---- dictionaries ------------------------------------------------------------------------------------- */
static size_t hfield(const char *line, const char *key, const char **val)
{
    const char *p = strchr(line, '<');
    const size_t kl = strlen(key);
    if ( !p ) return 0;
    p++;
    while ( *p && *p!='>' )
    {
        const char *e = p; int inq = 0;
        while ( *e && (inq || (*e!=',' && *e!='>')) ) { if ( *e=='"' ) inq = !inq; e++; }
        if ( (size_t)(e-p) > kl && !strncmp(p, key, kl) && p[kl]=='=' ) { *val = p+kl+1; return (size_t)(e-(p+kl+1)); }
        if ( *e!=',' ) break;
        p = e+1;
    }
    return 0;
}
static int dict_find(char **a, int n, const char *s, size_t l)
{
    for (int i=0; i<n; i++) if ( a[i] && strlen(a[i])==l && !strncmp(a[i], s, l) ) return i;
    return -1;
}
static int dict_set(char ***a, int *n, int at, const char *s, size_t l)
{
    if ( at >= *n )
    {
        char **p = (char**) realloc(*a, sizeof(char*)*(at+1));
        if ( !p ) return -1;
        for (int i=*n; i<=at; i++) p[i] = NULL;
        *a = p; *n = at+1;
    }
    if ( !(*a)[at] ) (*a)[at] = dupn(s, l);
    return (*a)[at] ? 0 : -1;
}
b200_bcfdict_t *b200_bcfdict_build(const b200_vhdr_t *h)
{
    b200_bcfdict_t *d = (b200_bcfdict_t*) calloc(1, sizeof *d);
    if ( !d || dict_set(&d->str, &d->nstr, 0, "PASS", 4) ) { b200_bcfdict_destroy(d); return NULL; }
    for (int i=0; i<h->nlines; i++)
    {
        const char *l = h->lines[i], *v, *x;
        const int is_ctg = !strncmp(l, "##contig=<", 10);
        if ( !is_ctg && strncmp(l, "##INFO=<", 8) && strncmp(l, "##FORMAT=<", 10) && strncmp(l, "##FILTER=<", 10) ) continue;
        const size_t n = hfield(l, "ID", &v);
        if ( !n ) continue;
        char ***arr = is_ctg ? &d->ctg : &d->str; int *cnt = is_ctg ? &d->nctg : &d->nstr;
        if ( dict_find(*arr, *cnt, v, n) >= 0 ) continue;
        int at = *cnt;
        if ( hfield(l, "IDX", &x) ) at = atoi(x);
        if ( dict_set(arr, cnt, at, v, n) ) { b200_bcfdict_destroy(d); return NULL; }
    }
    return d;
}
void b200_bcfdict_destroy(b200_bcfdict_t *d)
{
    if ( !d ) return;
    for (int i=0; i<d->nstr; i++) free(d->str[i]);
    for (int i=0; i<d->nctg; i++) free(d->ctg[i]);
    free(d->str); free(d->ctg); free(d);
}

/* ---- header block ------------------------------------------------------------------------------------- */
int b200_bcf_write_header(const b200_vhdr_t *h, b200_str_t *raw)
{
    b200_bcfdict_t *d = b200_bcfdict_build(h);
    if ( !d ) return -1;
    b200_str_t t = {0,0,0};
    for (int i=0; i<h->nlines; i++)
    {
        const char *l = h->lines[i], *v, *x;
        const int is_ctg = !strncmp(l, "##contig=<", 10);
        const int keyed = is_ctg || !strncmp(l, "##INFO=<", 8) || !strncmp(l, "##FORMAT=<", 10) || !strncmp(l, "##FILTER=<", 10);
        size_t n;
        if ( keyed && (n = hfield(l, "ID", &v)) && !hfield(l, "IDX", &x) && l[strlen(l)-1]=='>' )
        {
            b200_str_putsn(&t, l, strlen(l)-1);
            b200_str_puts(&t, ",IDX=");
            b200_str_putw(&t, dict_find(is_ctg ? d->ctg : d->str, is_ctg ? d->nctg : d->nstr, v, n));
            b200_str_puts(&t, ">\n");
        }
        else { b200_str_puts(&t, l); b200_str_putc(&t, '\n'); }
    }
    b200_bcfdict_destroy(d);
    b200_str_puts(&t, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO");
    if ( h->nsamples )
    {
        b200_str_puts(&t, "\tFORMAT");
        for (int i=0; i<h->nsamples; i++) { b200_str_putc(&t, '\t'); b200_str_puts(&t, h->samples[i]); }
    }
    b200_str_putc(&t, '\n');
    b200_str_putsn(raw, "BCF\2\2", 5);
    put_u32(raw, (uint32_t)(t.l + 1));
    b200_str_putsn(raw, t.s, t.l + 1);      /* with the terminating NUL */
    free(t.s);
    return 0;
}
b200_vhdr_t *b200_bcf_read_header(const uint8_t *raw, size_t n, size_t *used)
{
    if ( n < 9 || memcmp(raw, "BCF\2\2", 5) ) return NULL;
    const uint32_t l_text = get_u32(raw + 5);
    if ( 9 + (size_t)l_text > n ) return NULL;
    size_t len = l_text;
    while ( len && raw[9+len-1]==0 ) len--;
    size_t cons = 0;
    b200_vhdr_t *h = b200_vhdr_parse((const char*)raw + 9, len, &cons);
    if ( !h ) return NULL;
    /* the IDX= fields belong to the binary form only.  The dictionary must be built BEFORE they go: callers use
       b200_bcfdict_build on the returned header, so the order of first appearance has to reproduce them -- which it does
       for headers written by b200_bcf_write_header; foreign headers with sparse IDX values keep their IDX fields */
    int dense = 1;
    {
        b200_bcfdict_t *d = b200_bcfdict_build(h);
        if ( d ) { for (int i=0; i<d->nstr; i++) if ( !d->str[i] ) dense = 0; for (int i=0; i<d->nctg; i++) if ( !d->ctg[i] ) dense = 0; }
        b200_bcfdict_t *d2 = NULL;
        if ( d && dense )
        {
            /* would the dictionary come out the same without IDX? */
            b200_vhdr_t tmp = *h;
            char **lines = (char**) malloc(sizeof(char*)*(h->nlines ? h->nlines : 1));
            for (int i=0; i<h->nlines; i++)
            {
                lines[i] = dupn(h->lines[i], strlen(h->lines[i]));
                char *x = strstr(lines[i], ",IDX=");
                if ( x ) { char *e = x+5; while ( *e>='0' && *e<='9' ) e++; memmove(x, e, strlen(e)+1); }
            }
            tmp.lines = lines;
            d2 = b200_bcfdict_build(&tmp);
            if ( !d2 || d2->nstr!=d->nstr || d2->nctg!=d->nctg ) dense = 0;
            for (int i=0; dense && i<d->nstr; i++) if ( strcmp(d->str[i], d2->str[i]) ) dense = 0;
            for (int i=0; dense && i<d->nctg; i++) if ( strcmp(d->ctg[i], d2->ctg[i]) ) dense = 0;
            if ( dense ) { for (int i=0; i<h->nlines; i++) { free(h->lines[i]); h->lines[i] = lines[i]; } }
            else for (int i=0; i<h->nlines; i++) free(lines[i]);
            free(lines);
        }
        b200_bcfdict_destroy(d); b200_bcfdict_destroy(d2);
    }
    if ( used ) *used = 9 + (size_t)l_text;
    return h;
}

/* ---- typed values ------------------------------------------------------------------------------------- */
static void enc_size(b200_str_t *s, int size, int type);
static void enc_int1(b200_str_t *s, int32_t x)
{
    if ( x <= INT8_MAX && x > INT8_MIN+7 ) { enc_size(s, 1, BT_INT8); b200_str_putc(s, x); }
    else if ( x <= INT16_MAX && x > INT16_MIN+7 ) { enc_size(s, 1, BT_INT16); put_u16(s, (uint32_t)(uint16_t)(int16_t)x); }
    else { enc_size(s, 1, BT_INT32); put_u32(s, (uint32_t)x); }
}
static void enc_size(b200_str_t *s, int size, int type)
{
    if ( size >= 15 ) { b200_str_putc(s, 15<<4 | type); enc_int1(s, size); }
    else b200_str_putc(s, size<<4 | type);
}
static int int_type(const int32_t *a, int n)
{
    int32_t mn = INT32_MAX, mx = INT32_MIN + 2;
    for (int i=0; i<n; i++)
    {
        if ( a[i]==B200_I32_MISSING || a[i]==B200_I32_VECTOR_END ) continue;
        if ( a[i] < mn ) mn = a[i];
        if ( a[i] > mx ) mx = a[i];
    }
    if ( mx <= INT8_MAX && mn > INT8_MIN+7 ) return BT_INT8;
    if ( mx <= INT16_MAX && mn > INT16_MIN+7 ) return BT_INT16;
    return BT_INT32;
}
static void put_ints(b200_str_t *s, const int32_t *a, int n, int type)
{
    for (int i=0; i<n; i++)
    {
        const int32_t v = a[i];
        if ( type==BT_INT8 ) b200_str_putc(s, v==B200_I32_MISSING ? 0x80 : (v==B200_I32_VECTOR_END ? 0x81 : (v & 0xff)));
        else if ( type==BT_INT16 ) put_u16(s, v==B200_I32_MISSING ? 0x8000u : (v==B200_I32_VECTOR_END ? 0x8001u : (uint32_t)(uint16_t)(int16_t)v));
        else put_u32(s, (uint32_t)v);
    }
}
static void enc_vint(b200_str_t *s, const int32_t *a, int n)          /* one vector (bcf_enc_vint with wsize = n) */
{
    if ( n<=0 ) { enc_size(s, 0, BT_NULL); return; }
    const int t = int_type(a, n);
    enc_size(s, n, t);
    put_ints(s, a, n, t);
}
static void enc_vchar(b200_str_t *s, const char *p, size_t l) { enc_size(s, (int)l, BT_CHAR); b200_str_putsn(s, p, l); }
static uint32_t f32_bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }
static float f32_from(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }

static int parse_ints(const char *v, int32_t **out)
{
    int n = 1;
    for (const char *p = v; *p; p++) if ( *p==',' ) n++;
    int32_t *a = (int32_t*) malloc(sizeof(int32_t)*n);
    if ( !a ) return -1;
    for (int i=0; i<n; i++)
    {
        if ( *v=='.' && (v[1]==',' || !v[1]) ) { a[i] = B200_I32_MISSING; v++; }
        else { char *e; a[i] = (int32_t) strtol(v, &e, 10); if ( e==v ) { free(a); return -1; } v = e; }
        if ( *v==',' ) v++;
    }
    *out = a;
    return n;
}
static int parse_floats(const char *v, float **out)
{
    int n = 1;
    for (const char *p = v; *p; p++) if ( *p==',' ) n++;
    float *a = (float*) malloc(sizeof(float)*n);
    if ( !a ) return -1;
    for (int i=0; i<n; i++)
    {
        if ( *v=='.' && (v[1]==',' || !v[1]) ) { a[i] = f32_from(B200_F32_MISSING_BITS); v++; }
        else { char *e; a[i] = (float) strtod(v, &e); if ( e==v ) { free(a); return -1; } v = e; }
        if ( *v==',' ) v++;
    }
    *out = a;
    return n;
}
/* GT text of one sample -> htslib allele codes; returns the number of alleles */
static int parse_gt(const char *t, int32_t *out, int cap)
{
    int n = 0, phased = 0;
    while ( *t && n<cap )
    {
        if ( *t=='.' ) { out[n++] = 0 | phased; t++; }
        else { char *e; long a = strtol(t, &e, 10); if ( e==t ) return -1; out[n++] = (int32_t)((a+1)<<1 | phased); t = e; }
        if ( *t=='/' ) { phased = 0; t++; } else if ( *t=='|' ) { phased = 1; t++; } else break;
    }
    return n;
}

/* ---- record encode ------------------------------------------------------------------------------------ */
int b200_bcf_encode_rec(const b200_vhdr_t *h, const b200_bcfdict_t *d, const b200_vrec_t *r, b200_str_t *raw)
{
    b200_str_t sh = {0,0,0}, in = {0,0,0};
    int rc = -1;
    const int rid = dict_find(d->ctg, d->nctg, r->chrom, strlen(r->chrom));
    if ( rid<0 ) goto done;
    int64_t rlen = r->n_allele ? (int64_t)strlen(r->allele[0]) : 0;
    { int found; const char *e = b200_vrec_info(r, "END", &found); if ( found && e ) { long v = strtol(e, NULL, 10); if ( v > r->pos ) rlen = v - r->pos; } }
    put_u32(&sh, (uint32_t)rid); put_u32(&sh, (uint32_t)r->pos); put_u32(&sh, (uint32_t)rlen); put_u32(&sh, f32_bits(r->qual));
    put_u32(&sh, (uint32_t)r->n_allele<<16 | (uint32_t)r->n_info);
    put_u32(&sh, (uint32_t)r->n_fmt<<24 | (uint32_t)r->nsmpl);
    if ( strcmp(r->id, ".") ) enc_vchar(&sh, r->id, strlen(r->id)); else enc_size(&sh, 0, BT_CHAR);
    for (int i=0; i<r->n_allele; i++) enc_vchar(&sh, r->allele[i], strlen(r->allele[i]));
    if ( !strcmp(r->filter, ".") ) enc_size(&sh, 0, BT_NULL);
    else
    {
        int32_t fl[64]; int nf = 0;
        const char *p = r->filter;
        while ( *p && nf<64 )
        {
            const char *e = strchr(p, ';'); const size_t l = e ? (size_t)(e-p) : strlen(p);
            const int k = dict_find(d->str, d->nstr, p, l);
            if ( k<0 ) goto done;
            fl[nf++] = k;
            p += l; if ( *p==';' ) p++;
        }
        enc_vint(&sh, fl, nf);
    }
    for (int i=0; i<r->n_info; i++)
    {
        const b200_vdef_t *def = b200_vhdr_def(h, 0, r->info[i].key);
        const int k = dict_find(d->str, d->nstr, r->info[i].key, strlen(r->info[i].key));
        if ( !def || k<0 ) goto done;
        enc_int1(&sh, k);
        const char *v = r->info[i].val;
        if ( !v || def->type==B200_HT_FLAG ) enc_size(&sh, 0, BT_NULL);
        else if ( def->type==B200_HT_INT )
        {
            int32_t *a; const int n = parse_ints(v, &a);
            if ( n<0 ) goto done;
            enc_vint(&sh, a, n); free(a);
        }
        else if ( def->type==B200_HT_REAL )
        {
            float *a; const int n = parse_floats(v, &a);
            if ( n<0 ) goto done;
            enc_size(&sh, n, BT_FLOAT);
            for (int j=0; j<n; j++) put_u32(&sh, f32_bits(a[j]));
            free(a);
        }
        else enc_vchar(&sh, v, strlen(v));
    }
    for (int j=0; j<r->n_fmt; j++)
    {
        const b200_vfmt_t *f = &r->fmt[j];
        const b200_vdef_t *def = b200_vhdr_def(h, 1, f->key);
        const int k = dict_find(d->str, d->nstr, f->key, strlen(f->key));
        if ( !def || k<0 ) goto done;
        enc_int1(&in, k);
        const int is_gt = !strcmp(f->key, "GT");
        if ( f->kind==B200_FMT_REAL || (f->kind==B200_FMT_TEXT && def->type==B200_HT_REAL) )
        {
            int n = f->n; float *vals = NULL;
            if ( f->kind==B200_FMT_TEXT )
            {
                n = 1;
                for (int i=0; i<r->nsmpl; i++) { int c = 1; for (const char *p = f->txt[i]; *p; p++) if ( *p==',' ) c++; if ( c>n ) n = c; }
                vals = (float*) malloc(sizeof(float)*(size_t)n*r->nsmpl);
                for (int i=0; i<r->nsmpl; i++)
                {
                    float *a; const int m = parse_floats(f->txt[i], &a);
                    if ( m<0 ) { free(vals); goto done; }
                    for (int q=0; q<n; q++) vals[(size_t)i*n+q] = q<m ? a[q] : f32_from(B200_F32_VECTOR_END_BITS);
                    free(a);
                }
            }
            const float *src = vals ? vals : f->fv;
            enc_size(&in, n, BT_FLOAT);
            for (size_t q=0; q<(size_t)n*r->nsmpl; q++) put_u32(&in, f32_bits(src[q]));
            free(vals);
        }
        else if ( f->kind==B200_FMT_INT || f->kind==B200_FMT_GT || is_gt || def->type==B200_HT_INT )
        {
            int32_t *vals = NULL; int n = f->n;
            if ( f->kind==B200_FMT_TEXT && is_gt )
            {
                n = 1;
                int32_t tmp[16];
                for (int i=0; i<r->nsmpl; i++) { const int m = parse_gt(f->txt[i], tmp, 16); if ( m<0 ) goto done; if ( m>n ) n = m; }
                vals = (int32_t*) malloc(sizeof(int32_t)*(size_t)n*r->nsmpl);
                for (int i=0; i<r->nsmpl; i++)
                {
                    const int m = parse_gt(f->txt[i], tmp, 16);
                    for (int q=0; q<n; q++) vals[(size_t)i*n+q] = q<m ? tmp[q] : B200_I32_VECTOR_END;
                }
            }
            else if ( f->kind==B200_FMT_TEXT )
            {
                int mv = 0;
                const int tot = b200_vrec_fmt_ints(r, f->key, &vals, &mv);
                if ( tot<0 ) { free(vals); goto done; }
                n = r->nsmpl ? tot / r->nsmpl : 0;
            }
            const int32_t *src = vals ? vals : f->iv;
            const int t = int_type(src, n*r->nsmpl);
            enc_size(&in, n, t);
            put_ints(&in, src, n*r->nsmpl, t);
            free(vals);
        }
        else        /* strings: char vectors padded with NUL to the longest sample */
        {
            if ( f->kind!=B200_FMT_TEXT ) goto done;
            size_t n = 1;
            for (int i=0; i<r->nsmpl; i++) if ( strlen(f->txt[i]) > n ) n = strlen(f->txt[i]);
            enc_size(&in, (int)n, BT_CHAR);
            for (int i=0; i<r->nsmpl; i++)
            {
                const size_t l = strlen(f->txt[i]);
                b200_str_putsn(&in, f->txt[i], l);
                for (size_t q=l; q<n; q++) b200_str_putc(&in, 0);
            }
        }
    }
    put_u32(raw, (uint32_t)sh.l); put_u32(raw, (uint32_t)in.l);
    b200_str_putsn(raw, sh.s ? sh.s : "", sh.l);
    b200_str_putsn(raw, in.s ? in.s : "", in.l);
    rc = 0;
done:
    free(sh.s); free(in.s);
    return rc;
}

/* ---- record decode ------------------------------------------------------------------------------------ */
typedef struct { const uint8_t *p, *end; int bad; } rd_t;
static int rd_size(rd_t *r, int *type);
static int32_t rd_int(rd_t *r, int type)
{
    if ( type==BT_INT8 ) { if ( r->p+1 > r->end ) { r->bad = 1; return 0; } const int8_t v = (int8_t)*r->p++; return v==INT8_MIN ? B200_I32_MISSING : (v==INT8_MIN+1 ? B200_I32_VECTOR_END : v); }
    if ( type==BT_INT16 ) { if ( r->p+2 > r->end ) { r->bad = 1; return 0; } const int16_t v = (int16_t)get_u16(r->p); r->p += 2; return v==INT16_MIN ? B200_I32_MISSING : (v==INT16_MIN+1 ? B200_I32_VECTOR_END : v); }
    if ( type==BT_INT32 ) { if ( r->p+4 > r->end ) { r->bad = 1; return 0; } const int32_t v = (int32_t)get_u32(r->p); r->p += 4; return v; }
    r->bad = 1; return 0;
}
static int rd_size(rd_t *r, int *type)
{
    if ( r->p >= r->end ) { r->bad = 1; return 0; }
    const uint8_t b = *r->p++;
    *type = b & 0xf;
    int n = b >> 4;
    if ( n==15 ) { int t2; const int n2 = rd_size(r, &t2); if ( n2!=1 ) { r->bad = 1; return 0; } n = rd_int(r, t2); }
    /* a length from a damaged file: every element takes at least one byte, so it cannot exceed what is left of the block */
    if ( r->bad || n<0 || (size_t)n > (size_t)(r->end - r->p) ) { r->bad = 1; return 0; }
    return n;
}
static size_t bt_size(int t) { return t==BT_INT8 || t==BT_CHAR ? 1 : (t==BT_INT16 ? 2 : ((t==BT_INT32 || t==BT_FLOAT) ? 4 : 0)); }
static char *rd_str(rd_t *r, b200_vrec_t *rec, void *(*own)(b200_vrec_t*, void*))
{
    int t; const int n = rd_size(r, &t);
    if ( r->bad || (n && t!=BT_CHAR) || (size_t)n > (size_t)(r->end - r->p) ) { r->bad = 1; return NULL; }
    char *s = dupn((const char*)r->p, (size_t)n);
    r->p += n;
    return (char*) own(rec, s);
}
static void *own_ptr(b200_vrec_t *r, void *p)
{
    if ( !p ) return NULL;
    if ( r->nowned==r->mowned )
    {
        const int m = r->mowned ? 2*r->mowned : 16;
        void **o = (void**) realloc(r->owned, sizeof(void*)*m);
        if ( !o ) { free(p); return NULL; }
        r->owned = o; r->mowned = m;
    }
    r->owned[r->nowned++] = p;
    return p;
}
static void fmt_typed(b200_str_t *s, rd_t *r, int n, int t)      /* one typed vector as VCF text */
{
    if ( n<0 || !bt_size(t) || (size_t)n*bt_size(t) > (size_t)(r->end - r->p) ) { r->bad = 1; return; }        /* the whole vector lies inside the block */
    if ( t==BT_CHAR ) { size_t l = 0; while ( l<(size_t)n && r->p[l] ) l++; b200_str_putsn(s, (const char*)r->p, l); if ( !l ) b200_str_putc(s, '.'); r->p += n; return; }
    int k = 0;
    for (int i=0; i<n; i++)
    {
        if ( t==BT_FLOAT )
        {
            if ( r->p+4 > r->end ) { r->bad = 1; return; }
            const uint32_t b = get_u32(r->p); r->p += 4;
            if ( b==B200_F32_VECTOR_END_BITS ) { r->p += 4*(size_t)(n-i-1); break; }
            if ( k++ ) b200_str_putc(s, ',');
            if ( b==B200_F32_MISSING_BITS ) b200_str_putc(s, '.'); else b200_str_putd(s, f32_from(b));
        }
        else
        {
            const int32_t v = rd_int(r, t);
            if ( r->bad ) return;
            if ( v==B200_I32_VECTOR_END ) { r->p += (size_t)(t==BT_INT8 ? 1 : (t==BT_INT16 ? 2 : 4))*(size_t)(n-i-1); break; }
            if ( k++ ) b200_str_putc(s, ',');
            if ( v==B200_I32_MISSING ) b200_str_putc(s, '.'); else b200_str_putw(s, v);
        }
    }
    if ( !k ) b200_str_putc(s, '.');
}
b200_vrec_t *b200_bcf_decode_rec(const b200_vhdr_t *h, const b200_bcfdict_t *d, const uint8_t *p, size_t n, size_t *used)
{
    if ( n < 8 ) return NULL;
    const uint32_t l_shared = get_u32(p), l_indiv = get_u32(p + 4);
    if ( 8 + (size_t)l_shared + l_indiv > n || l_shared < 24 ) return NULL;
    rd_t r = { p + 8, p + 8 + l_shared, 0 };
    const int32_t rid = (int32_t)get_u32(r.p), pos = (int32_t)get_u32(r.p+4);
    const uint32_t qbits = get_u32(r.p+12), nai = get_u32(r.p+16), nfs = get_u32(r.p+20);
    r.p += 24;
    const int n_allele = nai >> 16, n_info = nai & 0xffff, n_fmt = nfs >> 24, nsmpl = nfs & 0xffffff;
    if ( rid<0 || rid>=d->nctg || !d->ctg[rid] || nsmpl!=h->nsamples ) return NULL;
    b200_vrec_t *rec = b200_vrec_new(nsmpl);
    if ( !rec ) return NULL;
    rec->chrom = (char*) own_ptr(rec, dupn(d->ctg[rid], strlen(d->ctg[rid])));
    rec->pos = pos;
    rec->qual = f32_from(qbits);
    {
        char *id = rd_str(&r, rec, own_ptr);
        rec->id = (id && *id) ? id : (char*)".";
    }
    rec->allele = (char**) malloc(sizeof(char*)*(n_allele ? n_allele : 1));
    for (int i=0; i<n_allele && !r.bad; i++) { rec->allele[i] = rd_str(&r, rec, own_ptr); if ( !rec->allele[i] ) r.bad = 1; }
    rec->n_allele = r.bad ? 0 : n_allele;
    if ( !r.bad )
    {
        int t; const int nf = rd_size(&r, &t);
        if ( nf<=0 ) rec->filter = (char*)".";
        else
        {
            b200_str_t s = {0,0,0};
            for (int i=0; i<nf && !r.bad; i++)
            {
                const int k = rd_int(&r, t);
                if ( r.bad || k<0 || k>=d->nstr || !d->str[k] ) { r.bad = 1; break; }
                if ( i ) b200_str_putc(&s, ';');
                b200_str_puts(&s, d->str[k]);
            }
            rec->filter = (char*) own_ptr(rec, s.s);
            if ( !rec->filter ) r.bad = 1;
        }
    }
    if ( !r.bad && n_info )
    {
        rec->info = (b200_vinfo_t*) calloc(n_info, sizeof(b200_vinfo_t)); rec->m_info = n_info;
        for (int i=0; i<n_info && !r.bad; i++)
        {
            int t; if ( rd_size(&r, &t)!=1 ) { r.bad = 1; break; }
            const int k = rd_int(&r, t);
            if ( r.bad || k<0 || k>=d->nstr || !d->str[k] ) { r.bad = 1; break; }
            rec->info[i].key = (char*) own_ptr(rec, dupn(d->str[k], strlen(d->str[k])));
            const int nv = rd_size(&r, &t);
            if ( r.bad ) break;
            if ( nv==0 ) rec->info[i].val = NULL;      /* a flag */
            else
            {
                b200_str_t s = {0,0,0};
                fmt_typed(&s, &r, nv, t);
                rec->info[i].val = (char*) own_ptr(rec, s.s);
            }
            rec->n_info++;
        }
    }
    if ( !r.bad && n_fmt )
    {
        rd_t q = { p + 8 + l_shared, p + 8 + l_shared + l_indiv, 0 };
        rec->fmt = (b200_vfmt_t*) calloc(n_fmt, sizeof(b200_vfmt_t)); rec->m_fmt = n_fmt;
        for (int j=0; j<n_fmt && !q.bad; j++)
        {
            int t; if ( rd_size(&q, &t)!=1 ) { q.bad = 1; break; }
            const int k = rd_int(&q, t);
            if ( q.bad || k<0 || k>=d->nstr || !d->str[k] ) { q.bad = 1; break; }
            b200_vfmt_t *f = &rec->fmt[rec->n_fmt];
            f->key = (char*) own_ptr(rec, dupn(d->str[k], strlen(d->str[k])));
            const int nv = rd_size(&q, &t);
            if ( q.bad ) break;
            f->n = nv;
            if ( !bt_size(t) ? nv!=0 : (size_t)nv*bt_size(t) > (size_t)(q.end - q.p)/(nsmpl ? (size_t)nsmpl : 1) ) { q.bad = 1; break; }     /* nsmpl vectors of nv values must fit */
            if ( t==BT_CHAR )
            {
                f->kind = B200_FMT_TEXT;
                f->txt = (char**) own_ptr(rec, calloc(nsmpl ? nsmpl : 1, sizeof(char*)));
                if ( !f->txt ) { q.bad = 1; break; }
                for (int i=0; i<nsmpl; i++)
                {
                    if ( q.p + nv > q.end ) { q.bad = 1; break; }
                    size_t l = 0; while ( l<(size_t)nv && q.p[l] ) l++;
                    f->txt[i] = (char*) own_ptr(rec, l ? dupn((const char*)q.p, l) : dupn(".", 1));
                    q.p += nv;
                }
            }
            else if ( t==BT_FLOAT )
            {
                f->kind = B200_FMT_REAL;
                f->fv = (float*) own_ptr(rec, malloc(sizeof(float)*((size_t)nv*nsmpl + 1)));
                if ( !f->fv ) { q.bad = 1; break; }
                for (size_t i=0; i<(size_t)nv*nsmpl; i++) { if ( q.p+4 > q.end ) { q.bad = 1; break; } f->fv[i] = f32_from(get_u32(q.p)); q.p += 4; }
            }
            else
            {
                f->kind = !strcmp(f->key, "GT") ? B200_FMT_GT : B200_FMT_INT;
                f->iv = (int32_t*) own_ptr(rec, malloc(sizeof(int32_t)*((size_t)nv*nsmpl + 1)));
                if ( !f->iv ) { q.bad = 1; break; }
                for (size_t i=0; i<(size_t)nv*nsmpl && !q.bad; i++) f->iv[i] = rd_int(&q, t);
            }
            rec->n_fmt++;
        }
        if ( q.bad ) r.bad = 1;
    }
    if ( r.bad ) { rec->chrom = (char*)"."; b200_vrec_destroy(rec); return NULL; }
    if ( used ) *used = 8 + (size_t)l_shared + l_indiv;
    return rec;
}

/* ---- whole files -------------------------------------------------------------------------------------- */
int b200_vcf_text_to_bcf(const char *text, size_t len, int level, b200_str_t *bcf)
{
    size_t off = 0;
    b200_vhdr_t *h = b200_vhdr_parse(text, len, &off);
    if ( !h ) return -1;
    b200_bcfdict_t *d = b200_bcfdict_build(h);
    b200_str_t raw = {0,0,0};
    int rc = d ? b200_bcf_write_header(h, &raw) : -1;
    while ( !rc && off < len )
    {
        const char *p = text + off, *e = (const char*) memchr(p, '\n', len - off);
        const size_t ll = e ? (size_t)(e-p) : len - off;
        off += ll + (e ? 1 : 0);
        if ( !ll ) continue;
        b200_vrec_t *r = b200_vrec_parse(h, p, ll);
        if ( !r ) { rc = -1; break; }
        rc = b200_bcf_encode_rec(h, d, r, &raw);
        b200_vrec_destroy(r);
    }
    if ( !rc ) rc = b200_bgzf_compress((const uint8_t*)raw.s, raw.l, level, bcf);
    if ( !rc ) rc = b200_bgzf_finish(bcf);
    free(raw.s); b200_bcfdict_destroy(d); b200_vhdr_destroy(h);
    return rc;
}
int b200_bcf_to_vcf_text(const uint8_t *bcf, size_t len, b200_str_t *text)
{
    b200_str_t raw = {0,0,0};
    if ( b200_bgzf_decompress(bcf, len, &raw) ) { free(raw.s); return -1; }
    size_t off = 0;
    b200_vhdr_t *h = b200_bcf_read_header((const uint8_t*)raw.s, raw.l, &off);
    if ( !h ) { free(raw.s); return -1; }
    b200_bcfdict_t *d = b200_bcfdict_build(h);
    int rc = d ? b200_vhdr_format(h, text) : -1;
    while ( !rc && off < raw.l )
    {
        size_t used = 0;
        b200_vrec_t *r = b200_bcf_decode_rec(h, d, (const uint8_t*)raw.s + off, raw.l - off, &used);
        if ( !r ) { rc = -1; break; }
        rc = b200_vrec_format(r, text);
        b200_vrec_destroy(r);
        off += used;
    }
    free(raw.s); b200_bcfdict_destroy(d); b200_vhdr_destroy(h);
    return rc;
}
