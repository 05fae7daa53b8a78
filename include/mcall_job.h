/*  mcall_job.h -- one `call -m` job over the N GPUs of a box (C-ABI, plain C).
 *
 *  The reference processes records strictly one after the other in one thread (vcfcall.c:1089-1148); sites are independent
 *  on this path (no state is carried between records: SURVEY.md 8e), so a job shards by CONTIGUOUS SITE RANGES, one per
 *  GPU, each range balanced by its FORMAT/PL bytes.  Every device has its own mcb_ctx (mcall_b200.h), its own host thread
 *  and its own slab pipeline; there is no collective and no NVLink traffic on the data path.  Results come back in input
 *  order: the per-site and per-sample arrays are indexed by site, and compacted trimmed-PL blocks (mcb_result.pl_off_out)
 *  stand range after range, in site order at ascending offsets, when the call returns -- the "ordered concatenation before
 *  BCF re-encoding" of the north star, equivalent to `bcftools concat` of region shards.  Every range compacts into its own
 *  stretch of the output, so a few unused elements may separate two ranges' blocks (pl_off_out[] says where each site's
 *  block is); mcb_job_set_option(job, "pack", 1) closes these gaps with a host-side move of the later ranges.
 *
 *  The same device may be listed more than once (two contexts on one GPU): that is how the single-GPU tests exercise the
 *  sharding logic.
 */
#ifndef MCALL_JOB_H
#define MCALL_JOB_H
#include "mcall_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcb_job mcb_job;

/*  One context per entry of devices[] (params->device is ignored).  Returns MCB_OK or the first context's error.  */
int  mcb_job_init(mcb_job **job, const mcb_params *params, const int *devices, int ndevices);
void mcb_job_destroy(mcb_job *job);
int  mcb_job_ndevices(const mcb_job *job);

/*  mcb_set_ploidy / mcb_set_option on every context of the job; the job's own option: "pack" (see above, default 0)  */
int  mcb_job_set_ploidy(mcb_job *job, int id, const uint8_t *ploidy);
int  mcb_job_set_option(mcb_job *job, const char *key, int64_t value);

/*  mcb_call_host over the whole batch: HOST pointers, same layouts as mcb_call_host.  The batch is cut into ndevices contiguous
 *  site ranges of about equal PL volume; range k runs on device k from its own host thread.  Returns when every result is
 *  in host memory.  first_site[0..ndevices] (optional, may be NULL) receives the range boundaries that were used.  */
int  mcb_job_call_host(mcb_job *job, const mcb_batch *batch, const mcb_result *result, int32_t *first_site);

/*  The range boundaries mcb_job_call_host uses (host arithmetic only, needs no GPU): nparts contiguous site ranges of about
 *  equal FORMAT/PL volume; first_site[0] = 0 <= ... <= first_site[nparts] = nsites.  bench.py shards its ranks with the same call.  */
int  mcb_job_partition(int32_t nsmpl, const uint8_t *nals, int32_t nsites, int32_t nparts, int32_t *first_site);

/*  the error text of the device that failed last (mcb_last_cuda_error of its context)  */
const char *mcb_job_last_error(const mcb_job *job);

#ifdef __cplusplus
}
#endif
#endif
