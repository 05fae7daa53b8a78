/*  b200_bcf.c -- BCF2 typed vectors of FORMAT fields without htslib (include/b200_bcf.h).  Host code, plain C.  */
#include <string.h>
#include <limits.h>
#include "b200_bcf.h"

static const int type_bytes[8] = { 0, 1, 2, 4, 0, 4, 0, 1 };

static int64_t rd_int(const uint8_t *p, int bytes)          /* little endian, sign extended */
{
    if ( bytes==1 ) return (int8_t)p[0];
    if ( bytes==2 ) return (int16_t)((uint16_t)p[0] | (uint16_t)p[1]<<8);
    return (int32_t)((uint32_t)p[0] | (uint32_t)p[1]<<8 | (uint32_t)p[2]<<16 | (uint32_t)p[3]<<24);
}
static void wr_int(uint8_t *p, int bytes, int64_t v)
{
    for (int i=0; i<bytes; i++) p[i] = (uint8_t)((uint64_t)v >> (8*i));
}
static int64_t miss_of(int bytes) { return bytes==1 ? INT8_MIN : (bytes==2 ? INT16_MIN : INT32_MIN); }

/*  typed scalar integer: descriptor byte (1<<4 | type) and the value  */
static int dec_typed_int(const uint8_t **pp, const uint8_t *end, int32_t *val)
{
    const uint8_t *p = *pp;
    if ( p >= end ) return B200_BCF_ETRUNC;
    const int t = p[0] & 0xf;
    if ( t!=B200_BT_INT8 && t!=B200_BT_INT16 && t!=B200_BT_INT32 ) return B200_BCF_ETYPE;
    const int b = type_bytes[t];
    if ( p + 1 + b > end ) return B200_BCF_ETRUNC;
    *val = (int32_t) rd_int(p+1, b);
    *pp = p + 1 + b;
    return 0;
}
/*  type descriptor: low nibble type, high nibble length, 15 = the length follows as a typed integer  */
static int dec_size(const uint8_t **pp, const uint8_t *end, int *type, int32_t *n)
{
    const uint8_t *p = *pp;
    if ( p >= end ) return B200_BCF_ETRUNC;
    *type = p[0] & 0xf;
    *n = p[0] >> 4;
    *pp = p + 1;
    if ( *n==15 ) { int rc = dec_typed_int(pp, end, n); if ( rc ) return rc; if ( *n < 0 ) return B200_BCF_ETYPE; }
    return 0;
}

int b200_bcf_unpack_fmt(const uint8_t *indiv, size_t len, int n_fmt, int n_sample, b200_bcf_fmt_t *fmt)
{
    const uint8_t *p = indiv, *end = indiv + len;
    for (int i=0; i<n_fmt; i++)
    {
        b200_bcf_fmt_t *f = fmt + i;
        int rc = dec_typed_int(&p, end, &f->key);
        if ( rc ) return rc;
        int type; int32_t n;
        rc = dec_size(&p, end, &type, &n);
        if ( rc ) return rc;
        if ( type > 7 || (type && !type_bytes[type]) ) return B200_BCF_ETYPE;
        f->type = type; f->n = n; f->size = n*type_bytes[type]; f->p = p;
        if ( (size_t)(end - p) < (size_t)f->size*(size_t)n_sample ) return B200_BCF_ETRUNC;
        p += (size_t)f->size*(size_t)n_sample;
    }
    return 0;
}

int b200_bcf_get_int(const b200_bcf_fmt_t *f, int n_sample, int dst_bytes, void *dst)
{
    if ( f->type!=B200_BT_INT8 && f->type!=B200_BT_INT16 && f->type!=B200_BT_INT32 ) return B200_BCF_ETYPE;
    if ( dst_bytes!=2 && dst_bytes!=4 ) return B200_BCF_ETYPE;
    const int sb = type_bytes[f->type];
    const size_t tot = (size_t)n_sample*(size_t)f->n;
    if ( sb==dst_bytes ) { memcpy(dst, f->p, tot*sb); return f->n; }        /* the typed vector travels as it is */
    const int64_t smiss = miss_of(sb), dmiss = miss_of(dst_bytes);
    uint8_t *d = (uint8_t*) dst;
    for (size_t i=0; i<tot; i++)
    {
        int64_t v = rd_int(f->p + i*sb, sb);
        if ( v==smiss ) v = dmiss;
        else if ( v==smiss+1 ) v = dmiss+1;
        else if ( dst_bytes==2 && (v > INT16_MAX || v < INT16_MIN+8) ) return B200_BCF_ERANGE;
        wr_int(d + i*dst_bytes, dst_bytes, v);
    }
    return f->n;
}

static int enc_typed_int(uint8_t *p, const uint8_t *end, int32_t v)     /* bcf_enc_int1 */
{
    int t = B200_BT_INT32;
    if ( v <= INT8_MAX && v >= INT8_MIN+8 ) t = B200_BT_INT8;
    else if ( v <= INT16_MAX && v >= INT16_MIN+8 ) t = B200_BT_INT16;
    const int b = type_bytes[t];
    if ( p + 1 + b > end ) return B200_BCF_ESPACE;
    p[0] = (uint8_t)(1<<4 | t);
    wr_int(p+1, b, v);
    return 1 + b;
}
static int enc_size(uint8_t *p, const uint8_t *end, int n, int type)    /* bcf_enc_size */
{
    if ( p >= end ) return B200_BCF_ESPACE;
    if ( n < 15 ) { p[0] = (uint8_t)(n<<4 | type); return 1; }
    p[0] = (uint8_t)(15<<4 | type);
    int k = enc_typed_int(p+1, end, n);
    return k<0 ? k : 1 + k;
}

int64_t b200_bcf_enc_int(uint8_t *dst, size_t cap, int key, const void *vals, int src_bytes, int n, int n_sample)
{
    if ( src_bytes!=1 && src_bytes!=2 && src_bytes!=4 ) return B200_BCF_ETYPE;
    const uint8_t *s = (const uint8_t*) vals;
    const size_t tot = (size_t)n*(size_t)n_sample;
    const int64_t smiss = miss_of(src_bytes);
    int64_t mx = INT64_MIN, mn = INT64_MAX;
    for (size_t i=0; i<tot; i++)                    /* bcf_enc_vint: sentinels do not take part in the type choice */
    {
        const int64_t v = rd_int(s + i*src_bytes, src_bytes);
        if ( v==smiss || v==smiss+1 ) continue;
        if ( v > mx ) mx = v;
        if ( v < mn ) mn = v;
    }
    int type = B200_BT_INT32;
    if ( mx <= INT8_MAX && mn >= INT8_MIN+8 ) type = B200_BT_INT8;                  /* BCF_MAX_BT_INT8 / BCF_MIN_BT_INT8 */
    else if ( mx <= INT16_MAX && mn >= INT16_MIN+8 ) type = B200_BT_INT16;
    const int db = type_bytes[type];
    uint8_t *p = dst; const uint8_t *end = dst + cap;
    int k = enc_typed_int(p, end, key);
    if ( k<0 ) return k;
    p += k;
    k = enc_size(p, end, n, type);
    if ( k<0 ) return k;
    p += k;
    if ( (size_t)(end - p) < tot*db ) return B200_BCF_ESPACE;
    if ( db==src_bytes ) { memcpy(p, s, tot*db); return (p - dst) + (int64_t)(tot*db); }
    const int64_t dmiss = miss_of(db);
    for (size_t i=0; i<tot; i++)
    {
        int64_t v = rd_int(s + i*src_bytes, src_bytes);
        if ( v==smiss ) v = dmiss;
        else if ( v==smiss+1 ) v = dmiss+1;
        wr_int(p + i*db, db, v);
    }
    return (p - dst) + (int64_t)(tot*db);
}
