/*  ORACLE-ONLY stub of <htslib/hts.h>: just what mcall.c uses [htslib].  */
#ifndef ORACLE_STUB_HTS_H
#define ORACLE_STUB_HTS_H
#include <stdint.h>
#include <stdlib.h>
typedef int64_t hts_pos_t;
#ifndef kroundup32
#define kroundup32(x) (--(x), (x)|=(x)>>1, (x)|=(x)>>2, (x)|=(x)>>4, (x)|=(x)>>8, (x)|=(x)>>16, ++(x))
#endif
#define hts_expand(type_t, n, m, ptr) do { \
        if ((n) > (m)) { (m) = (n); kroundup32(m); (ptr) = (type_t*)realloc((ptr), (size_t)(m) * sizeof(type_t)); } \
    } while (0)
#define hts_expand0(type_t, n, m, ptr) do { \
        if ((n) > (m)) { int t_ = (m); (m) = (n); kroundup32(m); (ptr) = (type_t*)realloc((ptr), (size_t)(m) * sizeof(type_t)); \
            memset(((type_t*)ptr)+t_, 0, sizeof(type_t)*((m)-t_)); } \
    } while (0)
char **hts_readlist(const char *fn, int is_file, int *_n);
#endif
