/*  b200_vcfcall.h -- the `bcftools call -m` driver around the B200 path, htslib-free host C.
 *
 *  main_vcfcall (vcfcall.c:906-1156) reads a record, filters it, sets the per-record state (unseen allele, ploidy),
 *  calls mcall() (mcall.c:1430-1684) -- which both computes AND edits the record -- and writes it, optionally through
 *  the gVCF block writer (gvcf.c:88-227) or after constraining the alleles to a targets file (mcall.c:1271-1421,
 *  vcfcall.c:359-606).  Here the computation runs on the GPU in batches, so the same steps are split in three:
 *
 *    b200_vc_next()    everything main_vcfcall and mcall() do BEFORE the likelihood code: next_line incl. -T targets and
 *                      the `-C alleles` duplicate-position buffer, -V / -M / -v pre-filters, the unseen allele,
 *                      set_ploidy, -i missed lines, mcall_constrain_alleles, and the unpacking of FORMAT/PL, INFO/QS or
 *                      FORMAT/AD, -F prior tags into a b200_rec_t;
 *    (the call)        b200_mcall() of b200_call.h -> mcb_call_host() on the device;
 *    b200_vc_finish()  everything mcall() and main_vcfcall do AFTER it, in input order: GP / GQ / trimmed PL, Number=R
 *                      tags, QUAL, AC / AN, ALT trimming, GT, DP4 / MQ / PV4, QS / I16 removal, gVCF blocks, output.
 *
 *  b200_vcfcall_run() is the whole command with the CUDA batcher in between (records are retained while their batch
 *  is on the device: the record ring of vcfbuf.c:146-175); the two halves are also exported so that the host logic can
 *  be replayed against the reference's expected outputs without a GPU (tests/test_vcfcall_host.py feeds them results
 *  computed by the CPU oracle -- the product never does).  Input may be text VCF, BGZF-compressed VCF or BCF; output is
 *  -O v / z / b / u (include/b200_bcfio.h), without the ##bcftools_ version lines (--no-version); -c, -C trio, -r/-R, -n, -p
 *  and --threads are not part of this path.
 *  Errors: functions return NULL / negative and leave a message in b200_vc_error(); nothing exits.
 */
#ifndef B200_VCFCALL_H
#define B200_VCFCALL_H

#include <stdint.h>
#include <stddef.h>
#include "b200_call.h"
#include "b200_vcf.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CALL_CONSTR_TRIO    (1<<2)      /* call.h:34-35 */
#define CALL_CONSTR_ALLELES (1<<3)
#define CALL_FMT_PV4        (1<<8)      /* call.h:39 */

typedef struct b200_vc b200_vc_t;
typedef struct b200_vcrec b200_vcrec_t;

/*  argv: the options of `bcftools call` (vcfcall.c:945-1062) without the program name and without the input file, e.g.
 *  {"-mv", "-S", "samples.txt"}; file arguments (-S, -G, -T, --ploidy-file) are paths.  vcf_text: the whole input VCF.  */
b200_vc_t *b200_vc_open(int argc, const char *const *argv, const char *vcf_text, size_t len, char *err, size_t errlen);
void b200_vc_close(b200_vc_t *vc);
const char *b200_vc_error(const b200_vc_t *vc);

/*  what b200_mcall_init needs (filled from the options and the header; the caller sets device / batch sizes)  */
void b200_vc_call_params(const b200_vc_t *vc, b200_call_t *call);
const uint8_t *b200_vc_ploidy(const b200_vc_t *vc);         /* the current per-sample ploidy vector (call->ploidy) */
int  b200_vc_unseen(const b200_vc_t *vc);                   /* the current record's unseen allele (call->unseen) */
int  b200_vc_output_type(const b200_vc_t *vc);              /* -O: 'v', 'z', 'b' or 'u' */

/*  Next record that goes to the caller: 1 and *rec / *in filled (in's pointers stay valid until the record is finished),
 *  0 at the end of the input, -1 on error.  Records the driver writes without calling (too many alleles) or drops are
 *  handled inside.  */
int  b200_vc_next(b200_vc_t *vc, b200_vcrec_t **rec, b200_rec_t *in);
/*  Apply the caller's result to the record and emit it (records must be finished in the order they were returned).  */
int  b200_vc_finish(b200_vc_t *vc, b200_vcrec_t *rec, const b200_out_t *out);
/*  after the last record: flush the gVCF block and the remaining -i lines  */
int  b200_vc_flush(b200_vc_t *vc);
/*  output text accumulated so far (header first); the caller may consume it with b200_vc_output_clear  */
const char *b200_vc_output(const b200_vc_t *vc, size_t *len);
void b200_vc_output_clear(b200_vc_t *vc);

/*  `bcftools call --no-version -O v <argv> in_path > out_path` with the device doing the calling; 0 on success  */
int  b200_vcfcall_run(int argc, const char *const *argv, const char *in_path, const char *out_path, int device, char *err, size_t errlen);

/*  test16 (ccall.c:103-138): the four PV4 p-values from INFO/I16; returns is_tested  */
int  b200_pv4(const float *anno16, float p[4]);
/*  vcmp.c:55-131, used by the constrained-alleles matching  */
int  b200_vcmp_set_ref(const char *ref1, const char *ref2, char *dref, size_t mdref, int *ndref);
int  b200_vcmp_find_allele(const char *dref, int ndref, const char *const *als1, int nals1, const char *al2);

#ifdef __cplusplus
}
#endif
#endif
