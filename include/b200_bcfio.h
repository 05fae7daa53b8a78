/*  b200_bcfio.h -- the BCF2.2 binary container without htslib: BGZF blocks, header, record framing.
 *
 *  `bcftools call -Ob|-Ou` writes, and every bcftools command reads, records as
 *      l_shared l_indiv | CHROM POS rlen QUAL n_info|n_allele<<16 n_sample|n_fmt<<24 ID alleles FILTER INFO... | FORMAT...
 *  with "typed values" (one descriptor byte: length << 4 | type; int8 / int16 / int32 / float / char vectors with
 *  missing and end-of-vector sentinels), behind a header block "BCF\2\2" + l_text + text, inside BGZF (concatenated
 *  gzip members of <= 64 KiB with a BC extra field, closed by the 28-byte empty block).  The layout is the published
 *  one (hts-specs VCFv4.2 §6 "BCF specification", SAMv1 §4.1 "The BGZF compression format"); in htslib it is
 *  vcf.c: bcf_hdr_write / bcf_write / vcf_parse + bcf_enc_vint / bcf_enc_vfloat / bcf_enc_vchar and bgzf.c.
 *
 *  PARITY STATUS: self-consistent, not pinned against htslib output -- the reference tree holds no BCF file
 *  (test/reheader.1.out.bcf is text) and htslib is not in this image.  What the tests can check they check: every VCF of
 *  the reference's `call` tests goes text -> BCF -> text unchanged, the container is valid gzip (Python's gzip module
 *  inflates it), and the typed-value encodings equal the worked examples of the specification.
 */
#ifndef B200_BCFIO_H
#define B200_BCFIO_H

#include <stdint.h>
#include <stddef.h>
#include "b200_vcf.h"

#ifdef __cplusplus
extern "C" {
#endif

/* dictionaries of a header: strings (PASS = 0, then FILTER / INFO / FORMAT IDs in order of first appearance or their IDX=) and contigs */
typedef struct
{
    char **str; int nstr;
    char **ctg; int nctg;
}
b200_bcfdict_t;

b200_bcfdict_t *b200_bcfdict_build(const b200_vhdr_t *h);
void b200_bcfdict_destroy(b200_bcfdict_t *d);

/*  BGZF: `level` 0..9 (-Ou is level 0: stored blocks).  The EOF block is appended by b200_bgzf_finish.  */
int b200_bgzf_compress(const uint8_t *raw, size_t n, int level, b200_str_t *out);
int b200_bgzf_finish(b200_str_t *out);
int b200_bgzf_decompress(const uint8_t *in, size_t n, b200_str_t *raw);        /* returns -1 if `in` is not a BGZF / gzip stream */

/*  header block: "BCF\2\2", l_text, the header text with IDX= fields, NUL  */
int b200_bcf_write_header(const b200_vhdr_t *h, b200_str_t *raw);
/*  parses a header block at raw[0..n): returns the header (IDX= fields removed from the lines), *used = bytes consumed  */
b200_vhdr_t *b200_bcf_read_header(const uint8_t *raw, size_t n, size_t *used);

/*  one record: text-model record -> BCF bytes (appended to raw) and back.  Values are typed by the header definitions.
 *  Returns 0 / the record, negative / NULL on malformed input or a tag that the header does not define.  */
int b200_bcf_encode_rec(const b200_vhdr_t *h, const b200_bcfdict_t *d, const b200_vrec_t *r, b200_str_t *raw);
b200_vrec_t *b200_bcf_decode_rec(const b200_vhdr_t *h, const b200_bcfdict_t *d, const uint8_t *p, size_t n, size_t *used);

/*  whole files in memory: VCF text <-> BCF (BGZF-compressed at `level`)  */
int b200_vcf_text_to_bcf(const char *text, size_t len, int level, b200_str_t *bcf);
int b200_bcf_to_vcf_text(const uint8_t *bcf, size_t len, b200_str_t *text);

#ifdef __cplusplus
}
#endif
#endif
