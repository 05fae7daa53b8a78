/*  ORACLE-ONLY stub of <htslib/synced_bcf_reader.h>: call.h only stores a pointer.  */
#ifndef ORACLE_STUB_SYNCED_H
#define ORACLE_STUB_SYNCED_H
typedef struct bcf_srs_t_ bcf_srs_t;
#endif
