/*  b200_bcf.h -- BCF2 typed vectors of FORMAT fields without htslib (first piece).
 *
 *  What crosses the device boundary can be the narrow typed vectors a BCF record stores (mcb_batch.pl_type = 2,
 *  mcb_result.gt8 / gq8 / pl16).  These helpers are the host-side counterpart: they locate and decode the FORMAT
 *  vectors inside a record's `indiv` block and encode result vectors the way htslib would, so that a batcher can feed
 *  slabs from raw records and emit records without the int32 round trip of
 *      bcf_get_format_int32   (mcall.c:1444, 1475)            -> b200_bcf_unpack_fmt + b200_bcf_get_int
 *      bcf_update_format_int32 / bcf_update_genotypes (mcall.c:1193, 1623, 1657) -> b200_bcf_enc_int
 *  The byte layout is the published one (hts-specs, BCFv2.2 "typed values" and "genotype fields"; in htslib:
 *  bcf_dec_size / bcf_unpack_fmt_core1 / bcf_enc_vint in vcf.c and vcf.h) -- htslib itself is not in the reference tree.
 *
 *  Plain C, no allocation, no exit(): every function returns a negative B200_BCF_E* code on malformed input.
 */
#ifndef B200_BCF_H
#define B200_BCF_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* typed-value type codes [htslib vcf.h BCF_BT_*] */
#define B200_BT_NULL   0
#define B200_BT_INT8   1
#define B200_BT_INT16  2
#define B200_BT_INT32  3
#define B200_BT_FLOAT  5
#define B200_BT_CHAR   7

#define B200_BCF_ETRUNC  -1     /* the block ends inside a value */
#define B200_BCF_ETYPE   -2     /* unknown type code / not an integer vector */
#define B200_BCF_ERANGE  -3     /* a value does not fit the requested destination type */
#define B200_BCF_ESPACE  -4     /* destination buffer too small */

/*  one FORMAT field inside `indiv` = bcf_fmt_t of htslib (vcf.h): n values of `type` per sample, samples contiguous  */
typedef struct b200_bcf_fmt
{
    int32_t key;            /* dictionary index of the tag (bcf_fmt_t.id) */
    int32_t type;           /* B200_BT_* */
    int32_t n;              /* values per sample */
    int32_t size;           /* bytes per sample = n * sizeof(type) */
    const uint8_t *p;       /* n_sample * size bytes */
}
b200_bcf_fmt_t;

/*  Decode the n_fmt FORMAT field headers of a record's indiv block (what bcf_unpack(rec, BCF_UN_FMT) does).  */
int b200_bcf_unpack_fmt(const uint8_t *indiv, size_t len, int n_fmt, int n_sample, b200_bcf_fmt_t *fmt);

/*  Copy an integer FORMAT vector into dst as int16 (dst_bytes = 2: the device's pl_type = 2 layout) or int32
 *  (dst_bytes = 4: exactly what bcf_get_format_int32 returns), widening narrower types and mapping the missing /
 *  vector_end sentinels to those of the destination type.  Returns the number of values per sample.  */
int b200_bcf_get_int(const b200_bcf_fmt_t *f, int n_sample, int dst_bytes, void *dst);

/*  Encode one integer FORMAT field (key, type descriptor, data) from n_sample*n values of src_bytes (1, 2 or 4) bytes
 *  each, in the smallest integer type that holds every value -- the choice bcf_update_format_int32 makes (int8 for
 *  -120..127, int16 for -32760..32767, else int32; sentinels excluded) -- so the bytes equal what htslib would write.
 *  Returns the number of bytes written.  */
int64_t b200_bcf_enc_int(uint8_t *dst, size_t cap, int key, const void *vals, int src_bytes, int n, int n_sample);

#ifdef __cplusplus
}
#endif
#endif
