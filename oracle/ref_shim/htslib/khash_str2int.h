/*  ORACLE-ONLY stub of <htslib/khash_str2int.h>: a string->int map with the same call
 *  signatures [htslib]; implemented as a small linear table in ../shim.c (only `-G <file>` uses it). */
#ifndef ORACLE_STUB_KHASH_STR2INT_H
#define ORACLE_STUB_KHASH_STR2INT_H
void *khash_str2int_init(void);
void  khash_str2int_destroy(void *hash);
void  khash_str2int_destroy_free(void *hash);
int   khash_str2int_has_key(void *hash, const char *str);
int   khash_str2int_get(void *hash, const char *str, int *value);
int   khash_str2int_set(void *hash, const char *str, int value);
#endif
