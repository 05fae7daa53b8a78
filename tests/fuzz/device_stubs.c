/* the device library is not part of this host-only fuzz build */
#include <stdlib.h>
#include "mcall_b200.h"
int mcb_init(mcb_ctx **c, const mcb_params *p) { (void)c; (void)p; return MCB_ENODEV; }
void mcb_destroy(mcb_ctx *c) { (void)c; }
int mcb_set_option(mcb_ctx *c, const char *k, int64_t v) { (void)c; (void)k; (void)v; return MCB_ENODEV; }
int mcb_set_ploidy(mcb_ctx *c, int id, const uint8_t *p) { (void)c; (void)id; (void)p; return MCB_ENODEV; }
int mcb_call_host(mcb_ctx *c, const mcb_batch *b, const mcb_result *r) { (void)c; (void)b; (void)r; return MCB_ENODEV; }
const char *mcb_last_cuda_error(const mcb_ctx *c) { (void)c; return ""; }
const char *mcb_strerror(int rc) { (void)rc; return "stub"; }
void *mcb_host_alloc(size_t n) { return malloc(n); }
void mcb_host_free(void *p) { free(p); }
