/* ASan fuzz driver for the htslib-free VCF / BCF readers: mutates a BCF (below the BGZF layer) and a text VCF */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "b200_vcf.h"
#include "b200_bcfio.h"
static uint64_t rs = 88172645463325252ull;
static uint32_t rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }
int main(int argc, char **argv)
{
    FILE *f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    char *text = malloc(n + 1); if ( fread(text, 1, n, f)!=(size_t)n ) return 1; text[n] = 0; fclose(f);
    int iters = atoi(argv[2]); rs ^= (uint64_t)atoi(argv[3]) * 0x9e3779b97f4a7c15ull;
    b200_str_t bcf = {0,0,0}, raw = {0,0,0};
    if ( b200_vcf_text_to_bcf(text, n, 1, &bcf) ) { fprintf(stderr, "to_bcf failed\n"); return 1; }
    if ( b200_bgzf_decompress((uint8_t*)bcf.s, bcf.l, &raw) ) { fprintf(stderr, "inflate failed\n"); return 1; }
    int ok = 0, rej = 0, tok = 0, trej = 0;
    for (int it=0; it<iters; it++)
    {
        /* ---- BCF */
        size_t m = raw.l; uint8_t *b = malloc(m + 16); memcpy(b, raw.s, m);
        int mode = rnd() % 3;
        if ( mode==0 ) m = rnd() % m;
        else if ( mode==1 ) { int k = 1 + rnd() % 4; while ( k-- ) { static const uint8_t sp[] = {0,1,0x7f,0x80,0xff,0x11,0xf7,0x15}; b[rnd() % m] = (rnd() & 1) ? sp[rnd() % 8] : (uint8_t)rnd(); } }
        else { size_t p = rnd() % m; int k = 1 + rnd() % 8; memmove(b + p + k, b + p, m - p); for (int j=0; j<k; j++) b[p+j] = (uint8_t)rnd(); m += k; }
        b200_str_t z = {0,0,0}, out = {0,0,0};
        b200_bgzf_compress(b, m, 1, &z); b200_bgzf_finish(&z);
        if ( b200_bcf_to_vcf_text((uint8_t*)z.s, z.l, &out)==0 ) ok++; else rej++;
        free(z.s); free(out.s); free(b);
        /* ---- text VCF: mutate, then text -> BCF (parses header and records) */
        size_t tn = n; char *t = malloc(tn + 16); memcpy(t, text, tn);
        mode = rnd() % 3;
        if ( mode==0 ) tn = rnd() % tn;
        else if ( mode==1 ) { int k = 1 + rnd() % 4; while ( k-- ) { static const char sp[] = "\t\n,;:=./|<>#0-"; t[rnd() % tn] = sp[rnd() % 15]; } }
        else { size_t p = rnd() % tn; int k = 1 + rnd() % 8; memmove(t + p + k, t + p, tn - p); for (int j=0; j<k; j++) t[p+j] = "\t\n,;:=./|<>#09AZ"[rnd() % 16]; tn += k; }
        t[tn] = 0;
        b200_str_t o2 = {0,0,0};
        if ( b200_vcf_text_to_bcf(t, tn, 1, &o2)==0 ) tok++; else trej++;
        free(o2.s); free(t);
    }
    printf("bcf ok %d rejected %d; text ok %d rejected %d\n", ok, rej, tok, trej);
    return 0;
}
