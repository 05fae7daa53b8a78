/* htslib declarations the glue uses that the oracle's stub headers (oracle/ref_shim/htslib) do not carry: type-check only */
typedef struct htsFile htsFile;
struct bcf1_t_fwd;
#include <htslib/vcf.h>
bcf1_t *bcf_dup(bcf1_t *src);
void bcf_destroy(bcf1_t *v);
int bcf_write1(htsFile *fp, const bcf_hdr_t *h, bcf1_t *v);
