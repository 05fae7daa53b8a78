/*  b200_driver.h -- the per-record state the `call` driver hands to mcall(), htslib-free host C:
 *
 *    ploidy definitions and the per-record ploidy vector   ploidy.c:40-260 (ploidy_init_string, ploidy_query,
 *                                                           ploidy_add_sex), vcfcall.c:807-825 (set_ploidy)
 *    -S sample / PED files -> sample subset and sexes        vcfcall.c:200-344 (parse_ped_samples, set_samples)
 *    -G sample-group files -> group member lists            mcall.c:250-349 (init_sample_groups)
 *    the unseen allele <*> of a record                      vcfcall.c:1101-1111
 *
 *  These produce exactly the inputs of the C-ABI: mcb_set_ploidy() vectors, mcb_params.grp_off / grp_smpl and
 *  mcb_batch.unseen.  No function exits; errors are negative codes plus a message in the caller's buffer.
 */
#ifndef B200_DRIVER_H
#define B200_DRIVER_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_DRV_EPARSE   -1    /* malformed line */
#define B200_DRV_EDUP     -2    /* "the sample is listed twice" (mcall.c:318) */
#define B200_DRV_EMISSING -3    /* "The sample is not listed" (mcall.c:335) / no matching samples (mcall.c:330) */
#define B200_DRV_ENOMEM   -4

/* ---- ploidy (ploidy.c) -------------------------------------------------------------------------- */
typedef struct b200_ploidy b200_ploidy_t;

/*  ploidy_init_string (ploidy.c:154-181): lines "CHROM FROM TO SEX PLOIDY", 1-based inclusive coordinates; CHROM "*" sets the
 *  default of a sex, SEX "*" the default of everything else; `dflt` applies where nothing is given.  NULL on a parse error.  */
b200_ploidy_t *b200_ploidy_init_string(const char *str, int dflt);
/*  --ploidy <alias> (init_ploidy, vcfcall.c:138-198, 827-861; case-insensitive): GRCh37, GRCh38 (sex chromosomes and MT of the
 *  two assemblies, with and without the "chr" prefix), X (males haploid), Y (males haploid, females absent), 1 (all haploid).
 *  NULL for an unknown alias.  */
b200_ploidy_t *b200_ploidy_init_alias(const char *alias);
void b200_ploidy_destroy(b200_ploidy_t *p);
int  b200_ploidy_add_sex(b200_ploidy_t *p, const char *sex);               /* ploidy.c:247-257: id of the sex, added with the default ploidy */
int  b200_ploidy_nsex(const b200_ploidy_t *p);
int  b200_ploidy_sex2id(const b200_ploidy_t *p, const char *sex);          /* -1 if unknown */
const char *b200_ploidy_id2sex(const b200_ploidy_t *p, int id);
int  b200_ploidy_min(const b200_ploidy_t *p);
int  b200_ploidy_max(const b200_ploidy_t *p);
/*  ploidy_query (ploidy.c:192-230): pos is 0-based (rec->pos).  Returns 1 if a region overlaps; sex2ploidy[nsex], min, max may be NULL.  */
int  b200_ploidy_query(const b200_ploidy_t *p, const char *seq, int64_t pos, int *sex2ploidy, int *min, int *max);
/*  set_ploidy (vcfcall.c:807-825): refreshes ploidy[nsmpl] when the per-sex ploidy changed since the previous record.
 *  sample2sex[i] >= 0: sex id; < 0: a fixed ploidy -sample2sex[i] (vcfcall.c:317-320).  sex2ploidy_prev[nsex] is caller-kept
 *  state, initialised to the maximum ploidy (vcfcall.c:652-655).  Returns 1 if ploidy[] was rewritten.  */
int  b200_set_ploidy(const b200_ploidy_t *p, const char *seq, int64_t pos, const int *sample2sex, int nsmpl,
                     int *sex2ploidy_prev, uint8_t *ploidy);

/* ---- -S samples files (vcfcall.c:114-130, 200-344) ---------------------------------------------- */
/*  text: the content of the -S file -- "name [sex-or-ploidy]" lines, '#' comments, or a PED file (>= 6 columns on every line:
 *  family, sample, father, mother, sex 1=M / other=F; parents named on a line are added as M / F when not listed themselves).
 *  For the i-th selected sample (file order): samples_map[i] = its index in hdr_samples, sample2sex[i] = the sex id
 *  (b200_ploidy_add_sex) or -ploidy for a literal 0 / 1 / 2; a missing second column means ploidy 2.  Samples that are not in
 *  the header or listed twice are skipped and counted in *nwarn (the reference prints a warning).  Capacity of both arrays:
 *  nhdr.  Without -S every sample gets the sex id nsex-1 (vcfcall.c:645-649): b200_samples_default.  */
int  b200_samples_parse(const char *text, const char *const *hdr_samples, int nhdr, b200_ploidy_t *ploidy,
                        int *samples_map, int *sample2sex, int *nsel, int *nwarn, char *err, size_t errlen);
void b200_samples_default(const b200_ploidy_t *ploidy, int nhdr, int *samples_map, int *sample2sex);

/* ---- -G groups (mcall.c:250-349) ------------------------------------------------------------------ */
/*  text: the content of the group file ("sample<whitespace>group" lines) or "-" for one group per sample.  Groups are
 *  numbered in the order their first listed sample appears in the file, members are in header order; samples of the file
 *  that are not in `samples` are skipped.  Like the reference (mcall.c:321-325 hashes ptr+1) populations are told apart by
 *  their name WITHOUT its first character.  Fills grp_off[ngroups+1] (capacity nsmpl+1) and grp_smpl[nsmpl].  */
int  b200_groups_parse(const char *text, const char *const *samples, int nsmpl, uint32_t *grp_off, uint32_t *grp_smpl,
                       int *ngroups, char *err, size_t errlen);

/* ---- record finaliser pieces ---------------------------------------------------- */
/*  mcall_trim_and_update_numberR (mcall.c:1196-1265) on one Number=R tag of 4-byte values (int32 or float): nvec vectors
 *  (1 for INFO, nsmpl for FORMAT) of nals_ori values each; dst[v][als_map[k]] = src[v][k] for every kept allele k
 *  (als_map from mcb_result.als_map).  Returns the number of values per output vector (nals_new).  */
int  b200_trim_numberR(const void *src, void *dst, int nvec, int nals_ori, int nals_new, const int8_t *als_map);
/*  INFO/I16 -> DP4 and MQ (mcall.c:1660-1666): float arithmetic as the reference does it, truncated to int32  */
void b200_i16_to_dp4_mq(const float *i16, int32_t *dp4, int32_t *mq);

/* ---- unseen allele (vcfcall.c:1101-1111) ---------------------------------------------------------- */
int  b200_unseen_allele(const char *const *alleles, int n_allele);         /* index of X, <X> or <*> among the ALTs, 0 = none */

#ifdef __cplusplus
}
#endif
#endif
