/* ASan fuzz driver for the front half of the call -m driver: b200_vc_open / b200_vc_next over mutated VCF text */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "b200_vcfcall.h"
static uint64_t rs = 88172645463325252ull;
static uint32_t rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }
int main(int argc, char **argv)
{
    FILE *f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    char *text = malloc(n + 1); if ( fread(text, 1, n, f)!=(size_t)n ) return 1; text[n] = 0; fclose(f);
    int iters = atoi(argv[2]); rs ^= (uint64_t)atoi(argv[3]) * 0x9e3779b97f4a7c15ull;
    const char *args[8]; int na = 0;
    for (int i=4; i<argc && na<8; i++) args[na++] = argv[i];
    int opened = 0, recs = 0, errs = 0;
    for (int it=0; it<iters; it++)
    {
        size_t tn = n; char *t = malloc(tn + 16); memcpy(t, text, tn);
        int mode = it ? rnd() % 3 : 3;
        if ( mode==0 ) tn = rnd() % tn;
        else if ( mode==1 ) { int k = 1 + rnd() % 4; while ( k-- ) t[rnd() % tn] = "\t\n,;:=./|<>#0-9"[rnd() % 16]; }
        else if ( mode==2 ) { size_t p = rnd() % tn; int k = 1 + rnd() % 8; memmove(t + p + k, t + p, tn - p); for (int j=0; j<k; j++) t[p+j] = "\t\n,;:=./|<>#09AZ"[rnd() % 16]; tn += k; }
        t[tn] = 0;
        char err[256];
        b200_vc_t *vc = b200_vc_open(na, args, t, tn, err, sizeof err);
        if ( vc )
        {
            opened++;
            b200_vcrec_t *rec; b200_rec_t in; int rc;
            while ( (rc = b200_vc_next(vc, &rec, &in)) > 0 ) recs++;      /* records are never finished: the device half is not here */
            if ( rc<0 ) errs++;
            b200_vc_close(vc);
        }
        free(t);
    }
    printf("opened %d records %d errors %d\n", opened, recs, errs);
    return 0;
}
